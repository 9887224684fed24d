"""A small tour of every kernel of the library for compute-sanitizer (memcheck / racecheck / synccheck):
K1 in its variants (kinematics only / mixed models, ragged M, inactive slots, map in shared and in global memory, goal,
ego binding), the drift pre-pass, K2 reset, K3 flat physics, K4 lidar, K5 controllers, the env epilogue and the done
exchange on a world of one.   compute-sanitizer --tool racecheck python profiles/tools/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tactics2d_b200 import BatchedWorld, TypeParams, TypeTable, synthetic
from tactics2d_b200.controller import AccelerationController, IDMController, PurePursuitController

dev = torch.device("cuda", 0)


def tour(scene, steps=2, **kw):
    n, m = scene.shape
    w = BatchedWorld(n, m, scene.table, device=dev, max_step=3, **kw)
    w.set_map(scene.segments, scene.bounds)
    w.set_state(scene.x, scene.y, scene.heading, scene.speed, vx=scene.vx, vy=scene.vy, type_id=scene.type_id)
    for t in range(steps):
        w.step(torch.from_numpy(synthetic.random_actions(t, (n, m))).to(dev))
    w.check_events()
    torch.cuda.synchronize()
    return w


w = tour(synthetic.config2(12, 64, seed=1, size=60.0))                       # dense: candidate queue, narrowphase, static phase
w.lidar_scan(72, 20.0)
tgt = np.stack([w.x[:, 0].cpu().numpy(), w.y[:, 0].cpu().numpy(), w.heading[:, 0].cpu().numpy(), np.full(12, 2.4), np.full(12, 1.0)], 1).astype(np.float32)
w.set_goal(tgt, 0.95, 2)
ego = torch.zeros((12, 2), device=dev)
w.set_ego_action(ego)
cid = np.zeros((12, 64), np.uint8); cid[:, 0] = 255; cid[:, 1::3] = 1; cid[:, 2::3] = 2
w.set_paths([np.array([[0, 0], [50, 10], [120, 10]], np.float32)])
w.set_controllers([IDMController(), AccelerationController(8.0), PurePursuitController()], cid,
                  lead_index=np.tile(np.arange(64, dtype=np.int16) - 1, (12, 1)), path_id=np.zeros((12, 64), np.int16))
act = torch.zeros((12, 64, 2), device=dev)
for _ in range(3):
    w.control(act); w.step(act); w.env_epilogue()
w.step_host_ego(np.zeros((12, 2), np.float32), act)
pool = {k: getattr(w, k).clone() for k in ("x", "y", "heading", "speed", "vx", "vy")}
w.reset(torch.ones(12, dtype=torch.uint8, device=dev), pool)
torch.cuda.synchronize(); w.close()
tour(synthetic.with_inactive(synthetic.config2(7, 13, seed=2, size=40.0), 0.2, seed=1)).close()      # ragged M, inactive slots
tour(synthetic.config4(9, 32, seed=3, size=60.0), any_participant=True).close()                      # mixed models + discs
tour(synthetic.config3(5, 16, seed=4)).close()                                                         # fp64 dynamics
tour(synthetic.config5(3, 128, seed=5, size=120.0)).close()                                            # one scenario per warp
big = synthetic.config2(4, 64, seed=6)
rng = np.random.default_rng(0)
segs = rng.uniform(0, 200, (9000, 4)).astype(np.float32); segs[:, 2:] = segs[:, :2] + rng.uniform(-3, 3, (9000, 2)).astype(np.float32)
big = synthetic.Scene(big.table, big.x, big.y, big.heading, big.speed, big.vx, big.vy, big.type_id, segs, big.bounds, "global-map", {})
tour(big, steps=1).close()                                                                              # map blob in global memory
table = TypeTable([TypeParams.vehicle("medium_car", model="drift")])
w = BatchedWorld(4, 8, table, device=dev)
w.set_state(np.zeros((4, 8)), np.zeros((4, 8)), np.zeros((4, 8)), np.full((4, 8), 5.0), type_id=np.zeros((4, 8), np.uint8))
w.set_wheel_state(np.full((4, 8), 14.5), np.full((4, 8), 14.5))
w.step(torch.zeros((4, 8, 2), device=dev)); torch.cuda.synchronize(); w.close()
from tactics2d_b200.physics import SingleTrackKinematics
import torch.distributed as dist
dist.init_process_group("gloo", init_method="file:///tmp/t2d_sanitize_rdv", rank=0, world_size=1)
from tactics2d_b200.distributed import PeerDoneExchange
for lag in (0, 2):
    ex = PeerDoneExchange(100, dev, lag=lag)
    for t in range(7):
        ex(torch.ones(100, dtype=torch.uint8, device=dev))
    torch.cuda.synchronize(); ex.close()
print("SANITIZE_TOUR_DONE")
