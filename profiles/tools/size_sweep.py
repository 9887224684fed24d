"""Tick time against the batch size (M = 64 kinematic participants, grid map): the latency floor of one warp tile's serial
chain and the slope once the GPU is full.  CUDA graph of 48 ticks over R world replicas, states restored before every rep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tactics2d_b200 import BatchedWorld, synthetic

dev = torch.device("cuda", 0)
m, K = int(os.environ.get("SW_M", 64)), 48
for n in [int(v) for v in os.environ.get("SW_N", "37,296,1184,2368,4096,4736,8192,16384,32768").split(",")]:
    R = max(4, min(24, int(350e6 // (n * m * 54))))
    worlds, acts, pools = [], [], []
    for r in range(R):
        sc = synthetic.config2(n, m, seed=1 + r)
        w = BatchedWorld(n, m, sc.table, device=dev, max_step=0)
        w.set_map(sc.segments, sc.bounds)
        w.set_state(sc.x, sc.y, sc.heading, sc.speed, type_id=sc.type_id)
        worlds.append(w)
        acts.append(torch.from_numpy(synthetic.random_actions(9000 + r, (n, m))).to(dev))
        pools.append({k: getattr(w, k).clone() for k in ("x", "y", "heading", "speed", "vx", "vy")})
    ones = torch.ones(n, dtype=torch.uint8, device=dev)

    def body():
        for i in range(K):
            worlds[i % R].step(acts[i % R])
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            body()
    ts = []
    for rep in range(12):
        for w, p in zip(worlds, pools):
            w.reset(ones, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K * 1e3)
    t = float(np.median(ts[2:]))
    print("N %6d x M %d: %7.2f us / tick  %6.2f G participant-steps/s  (%d warp tiles, %d replicas)" % (n, m, t, n * m / t / 1e3, n * m // 128, R), flush=True)
    for w in worlds:
        w.close()
