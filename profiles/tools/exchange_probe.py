"""Where the multi-GPU step spends its time: the tick alone, the tick publishing into the peer ring, plus the gather
kernel on the same / on a side stream, and the NCCL all-gather, each as a CUDA graph of 48 steps.  torchrun, N >= 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from tactics2d_b200 import BatchedWorld, synthetic
from tactics2d_b200.distributed import PeerDoneExchange

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
n, m, K = 4096, 64, 48
sc = synthetic.config2(n, m, seed=1 + rank)
w = BatchedWorld(n, m, sc.table, device=dev)
w.set_map(sc.segments, sc.bounds)
w.set_state(sc.x, sc.y, sc.heading, sc.speed, type_id=sc.type_id)
act = torch.from_numpy(synthetic.random_actions(3, (n, m))).to(dev)
done_all = torch.zeros(world * (n + 16), dtype=torch.uint8, device=dev)
exs = {lag: PeerDoneExchange(n, dev, lag=lag) for lag in (0, 1, 2)}
ex = exs[0]
side = torch.cuda.Stream(dev)


def build(mode):
    use_side = mode.endswith("_side")
    px = exs[int(mode[4])] if mode.startswith("peer") else None   # "peerL_same" / "peerL_side": lag L

    def body(cap):
        for i in range(K):
            out = w.step(act)
            if mode == "none":
                continue
            if use_side:
                e = torch.cuda.Event(); e.record(cap); side.wait_event(e)
            with torch.cuda.stream(side if use_side else cap):
                if mode.startswith("peer"):
                    px(out.done, done_all[: world * px.pad])
                else:
                    dist.all_gather_into_tensor(done_all[: world * n], out.done)
        if use_side:
            cap.wait_stream(side)
    cap = torch.cuda.Stream(dev)
    with torch.cuda.stream(cap):
        body(cap)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            body(cap)
    return g


def timed(g, reps=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (reps * K) * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {}
for name, mode in (("tick only", "none"), ("tick + NCCL all_gather, same stream", "nccl_same"), ("tick + NCCL all_gather, side stream", "nccl_side"),
                   ("tick + peer-memory all-gather kernel, same stream", "peer0_same"), ("tick + peer-memory all-gather kernel, side stream", "peer0_side"),
                   ("tick + peer kernel, lag 1, same stream", "peer1_same"), ("tick + peer kernel, lag 1, side stream", "peer1_side"),
                   ("tick + peer kernel, lag 2, same stream", "peer2_same"), ("tick + peer kernel, lag 2, side stream", "peer2_side")):
    try:
        res[name] = timed(build(mode))
    except Exception as e:   # keep going: the other rows are still informative
        res[name] = float("nan")
        if rank == 0:
            print("FAILED", name, type(e).__name__, str(e)[:300])
if rank == 0:
    for k, v in res.items():
        print("%-52s %7.2f us / step" % (k, v))
    print("exchange status (steps, timed_out) per lag:", {lag: x.status() for lag, x in exs.items()})
sys.stdout.flush(); dist.barrier(); os._exit(0)
