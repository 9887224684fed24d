"""Two (or more) GPU check of the peer-memory done exchange, launched by torchrun (see test_gpu_exchange.py):
every rank ticks its own shard; after each tick every rank must hold every rank's done mask, equal to what NCCL's
all_gather gives, for eager launches and for a CUDA graph of several steps with the exchange on a side stream."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tactics2d_b200 import BatchedWorld, synthetic  # noqa: E402
from tactics2d_b200.distributed import PeerDoneExchange  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    n, m = 200, 24                       # 200 is not a multiple of 16: the gathered rows are padded to 208
    scene = synthetic.config2(n, m, seed=100 + rank, size=60.0)
    w = BatchedWorld(n, m, scene.table, device=dev, max_step=5)
    w.set_map(scene.segments, scene.bounds)
    w.set_state(scene.x, scene.y, scene.heading, scene.speed, type_id=scene.type_id)
    ex = PeerDoneExchange(n, dev, slots=2)
    ref = torch.zeros(world * n, dtype=torch.uint8, device=dev)
    seen = 0
    for t in range(7):                   # eager: tick, exchange on the same stream, compare with NCCL
        act = torch.from_numpy(synthetic.random_actions(1000 * rank + t, (n, m))).to(dev)
        out = w.step(act)
        got = ex(out.done).clone()
        dist.all_gather_into_tensor(ref, out.done)
        torch.cuda.synchronize()
        assert torch.equal(got.view(world, ex.pad)[:, :n].reshape(-1), ref), (rank, t)
        assert int(got.view(world, ex.pad)[:, n:].sum().item()) == 0
        seen += int(ref.sum().item())
    assert seen > 0
    # CUDA graph of 6 ticks with the exchange on a side stream (the tick never waits for it)
    side = torch.cuda.Stream(dev)
    outs = [torch.zeros(world * ex.pad, dtype=torch.uint8, device=dev) for _ in range(6)]
    dones = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(6)]
    act = torch.from_numpy(synthetic.random_actions(77 + rank, (n, m))).to(dev)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    with torch.cuda.stream(cap):
        with torch.cuda.graph(g, stream=cap):
            for i in range(6):
                out = w.step(act)
                dones[i].copy_(out.done)
                e = torch.cuda.Event()
                e.record(cap)
                side.wait_event(e)
                with torch.cuda.stream(side):
                    ex(dones[i], outs[i])
            cap.wait_stream(side)
    for rep in range(3):
        g.replay()
        torch.cuda.synchronize()
        for i in range(6):
            dist.all_gather_into_tensor(ref, dones[i])
            torch.cuda.synchronize()
            assert torch.equal(outs[i].view(world, ex.pad)[:, :n].reshape(-1), ref), (rank, rep, i)
    assert ex.status() == (7 + 18, 0), ex.status()
    # lagged delivery: call k posts step k's mask and returns the gathered masks of step k - lag (off the critical path)
    for lag in (1, 2):
        exl = PeerDoneExchange(n, dev, lag=lag)
        hist = []
        sentinel = torch.full((world * exl.pad,), 7, dtype=torch.uint8, device=dev)
        for t in range(9):
            act = torch.from_numpy(synthetic.random_actions(5000 * lag + 1000 * rank + t, (n, m))).to(dev)
            out = w.step(act)
            buf = sentinel.clone()
            got = exl(out.done, buf).clone()
            dist.all_gather_into_tensor(ref, out.done)
            torch.cuda.synchronize()
            hist.append(ref.clone())
            if t < lag:
                assert torch.equal(got, sentinel), (rank, lag, t)          # nothing delivered yet: dst untouched
            else:
                assert torch.equal(got.view(world, exl.pad)[:, :n].reshape(-1), hist[t - lag]), (rank, lag, t)
        assert exl.status() == (9, 0), exl.status()
        exl.close()
    dist.barrier()
    if rank == 0:
        print("EXCHANGE_OK", world)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
