"""Where the tick's time goes, measured by leaving work out (CUDA graph of 64 ticks over 24 world replicas, C2 shapes):
full tick / 1 sub-step / no map / no bounds / events only (no physics) / physics only on shapeless participants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tactics2d_b200 import BatchedWorld, synthetic
from tactics2d_b200.types import TypeTable

dev = torch.device("cuda", 0)
n, m, K, R = int(os.environ.get("ABL_N", 4096)), int(os.environ.get("ABL_M", 64)), 64, 24


def run(name, interval=100, use_map=True, use_bounds=True, physics=True, shapeless=False):
    worlds, acts, pools = [], [], []
    ones = torch.ones(n, dtype=torch.uint8, device=dev)
    for r in range(R):
        sc = synthetic.config2(n, m, seed=1 + r)
        table = sc.table
        if shapeless:
            rows = [type(row)(**{**row.__dict__, "shape": 2}) for row in table.rows]
            table = TypeTable(rows)
        w = BatchedWorld(n, m, table, device=dev, interval=interval, max_step=0)
        w.set_map(sc.segments if use_map else None, sc.bounds if use_bounds else None)
        w.set_state(sc.x, sc.y, sc.heading, sc.speed, type_id=sc.type_id)
        worlds.append(w)
        acts.append(torch.from_numpy(synthetic.random_actions(9000 + r, (n, m))).to(dev))
        pools.append({k: getattr(w, k).clone() for k in ("x", "y", "heading", "speed", "vx", "vy")})

    def body():
        for i in range(K):
            if physics:
                worlds[i % R].step(acts[i % R])
            else:
                worlds[i % R].check_events()
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        for w, p in zip(worlds, pools):   # fresh states every repetition, as bench.py does
            w.reset(ones, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K * 1e3)
    print("%-44s %7.2f us / tick" % (name, float(np.median(ts))), flush=True)
    for w in worlds:
        w.close()


run("full tick")
run("1 sub-step per tick (interval 5 ms)", interval=5)
run("no map (no static phase)", use_map=False)
run("no bounds", use_bounds=False)
run("no map, no bounds", use_map=False, use_bounds=False)
run("events only (check_events: no physics)", physics=False)
try:
    run("shapeless participants (physics + empty pair loop)", shapeless=True)
except Exception as e:
    print("shapeless: failed", type(e).__name__, e)
