"""Per-warp phase timeline of t2d_step (clock64 stamps at the phase boundaries) on config 2.

usage (GPU box): python profiles/phase_clocks.py [n_scenarios] [m_participants]
"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tactics2d_b200 import BatchedWorld, _lib, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
scene = synthetic.config2(n, m, seed=1)
w = BatchedWorld(n, m, scene.table, device="cuda:0")
w.set_map(scene.segments, scene.bounds)
w.set_state(scene.x, scene.y, scene.heading, scene.speed, type_id=scene.type_id)
act = torch.from_numpy(synthetic.random_actions(3, (n, m))).cuda()
for _ in range(3):
    w.step(act)
tiles = n  # upper bound on warp tiles
buf = torch.zeros((tiles, 10), dtype=torch.int64, device="cuda")
_lib.check(w.lib.t2d_debug_set_clock_buffer(w._ctx, C.c_void_p(buf.data_ptr())))
torch.cuda.synchronize()
w.step(act)
torch.cuda.synchronize()
b = buf.cpu().numpy()
b = b[b[:, 0] != 0]
t0 = b[:, 0].min()
names = ["load", "physics", "pose", "pair loop", "pair drain", "static", "oob+status"]
smid, entry = b[:, 8], b[:, 9]
b = b[:, :8]
d = np.diff(b, axis=1)
# per-SM view (clock64 is per SM): when do warps enter the kernel, and how long is an SM busy?
spans, ramps, counts = [], [], []
for sm in np.unique(smid):
    sel = smid == sm
    e0 = entry[sel].min()
    spans.append(b[sel, 7].max() - e0)
    ramps.append(entry[sel].max() - e0)
    counts.append(sel.sum())
print(f"per SM: tiles {np.mean(counts):.1f} (min {np.min(counts)}, max {np.max(counts)}); busy span mean {np.mean(spans):.0f} max {np.max(spans)}; "
      f"last warp enters {np.mean(ramps):.0f} cycles after the first (max {np.max(ramps)}); prologue (entry -> first stamp) mean {np.mean(b[:, 0] - entry):.0f}")
print(f"{len(b)} warp tiles; kernel span {b[:, 7].max() - t0} cycles (SM clock, per-SM counters: spans across SMs are approximate)")
print(f"tile start (rel. first): mean {np.mean(b[:, 0] - t0):.0f}  p50 {np.percentile(b[:, 0] - t0, 50):.0f}  max {np.max(b[:, 0] - t0)}")
print(f"tile lifetime: mean {np.mean(b[:, 7] - b[:, 0]):.0f}  p95 {np.percentile(b[:, 7] - b[:, 0], 95):.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:10s} mean {d[:, i].mean():8.0f}  p50 {np.percentile(d[:, i], 50):8.0f}  p95 {np.percentile(d[:, i], 95):8.0f}  share {100 * d[:, i].mean() / (b[:, 7] - b[:, 0]).mean():5.1f}%")
