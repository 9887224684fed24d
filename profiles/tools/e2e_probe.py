"""Where the end-to-end (host-buffer) tick spends its time: the pieces timed alone and the two public paths.
Run on a GPU box: python profiles/tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from tactics2d_b200 import BatchedWorld, synthetic

dev = torch.device("cuda", 0)
sc = synthetic.config2(4096, 64, seed=1)
n, m = sc.shape

def world():
    w = BatchedWorld(n, m, sc.table, device=dev, max_step=0)
    w.set_map(sc.segments, sc.bounds)
    w.set_state(sc.x, sc.y, sc.heading, sc.speed, vx=sc.vx, vy=sc.vy, type_id=sc.type_id)
    return w

def wall(fn, reps=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e6

act = torch.from_numpy(synthetic.random_actions(5, (n, m))).pin_memory()
dact = torch.empty((n, m, 2), dtype=torch.float32, device=dev)
hd = torch.empty(n, dtype=torch.uint8).pin_memory(); hs = torch.empty(n, dtype=torch.uint8).pin_memory()
st = torch.cuda.current_stream(dev)
w = world()
def h2d(): dact.copy_(act, non_blocking=True); st.synchronize()
def kern(): w.step(dact); st.synchronize()
o = w.step(dact)
def d2h(): hd.copy_(o.done, non_blocking=True); st.synchronize()
def torch_path():
    dact.copy_(act, non_blocking=True); out = w.step(dact)
    hd.copy_(out.done, non_blocking=True); hs.copy_(out.status, non_blocking=True); st.synchronize()
print("h2d 2 MiB + sync      %.1f us  (%.1f GB/s incl. sync)" % (wall(h2d), act.numel() * 4 / wall(h2d) / 1e3))
for mb in (0.25, 0.5, 1, 2, 8):
    a = torch.empty(int(mb * 2**20 // 4), dtype=torch.float32).pin_memory(); d = torch.empty_like(a, device=dev)
    f = lambda: (d.copy_(a, non_blocking=True), st.synchronize())
    print("  h2d %.2f MiB %.1f us" % (mb, wall(f)))
print("kernel + sync         %.1f us" % wall(kern))
print("d2h 4 KiB + sync      %.1f us" % wall(d2h))
print("torch path            %.1f us" % wall(torch_path))
for ch in (1, 2, 3, 4):
    os.environ["T2D_HOST_CHUNKS"] = str(ch)
    w2 = world()
    print("step_host chunks=%d    %.1f us" % (ch, wall(lambda: w2.step_host(act))))
