"""CUDA-event timings of the kernels next to the tick (lidar scan K4, NPC controllers K5, reset K2, drift / dynamics
models inside K1) at the C2 batch size.  Run on a GPU box: python profiles/tools/aux_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from tactics2d_b200 import BatchedWorld, TypeParams, TypeTable, synthetic
from tactics2d_b200.controller import AccelerationController, IDMController, PurePursuitController

dev = torch.device("cuda", 0)
n, m = 4096, 64


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


sc = synthetic.config2(n, m, seed=1)
w = BatchedWorld(n, m, sc.table, device=dev)
w.set_map(sc.segments, sc.bounds)
w.set_state(sc.x, sc.y, sc.heading, sc.speed, vx=sc.vx, vy=sc.vy, type_id=sc.type_id)
act = torch.from_numpy(synthetic.random_actions(3, (n, m))).to(dev)
print("K1 step, kinematics, 4096 x 64            %8.1f us" % timed(lambda: w.step(act)))
for beams, rng in ((360, 20.0), (500, 12.0)):
    print("K4 lidar %3d beams, range %4.1f m, 4096 egos  %8.1f us" % (beams, rng, timed(lambda: w.lidar_scan(beams, rng))))
rs = np.random.default_rng(0)
ctrls = [IDMController(), AccelerationController(8.0), PurePursuitController()]
w.set_paths([np.array([[0, 0], [50, 10], [120, 10], [200, 60]], np.float32)])
cid = rs.integers(0, 3, size=(n, m)).astype(np.uint8)
cid[:, 0] = 255
w.set_controllers(ctrls, cid, rs.integers(-1, m, size=(n, m)).astype(np.int16), np.zeros((n, m), np.int16))
t = timed(lambda: w.control(act))
print("K5 controllers (IDM / cruise / pure pursuit mix)  %8.1f us   (%.0f GB/s of ~30 B / participant)" % (t, n * m * 30 / t / 1e3))
ones = torch.ones(n, dtype=torch.uint8, device=dev)
pool = {k: getattr(w, k).clone() for k in ("x", "y", "heading", "speed", "vx", "vy")}
t = timed(lambda: w.reset(ones, pool))
print("K2 reset of all scenarios                  %8.1f us   (%.0f GB/s of 48 B / participant)" % (t, n * m * 48 / t / 1e3))
for model in ("dynamics", "drift"):
    table = TypeTable([TypeParams.vehicle("medium_car", model=model)])
    w2 = BatchedWorld(n, m, table, device=dev)
    w2.set_map(sc.segments, sc.bounds)
    sp = np.clip(sc.speed, 3, None)
    w2.set_state(sc.x, sc.y, sc.heading, sp, type_id=np.zeros((n, m), np.uint8))
    if model == "drift":
        w2.set_wheel_state(sp / 0.344, sp / 0.344)
    print("K1 step, %-9s (fp64), 4096 x 64        %8.1f us" % (model, timed(lambda: w2.step(act), reps=10, warm=2)))
