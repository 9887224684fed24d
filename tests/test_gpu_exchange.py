"""The peer-memory done exchange (t2d_exchange_*): on one GPU with a world of one (the kernel paths: stores into the
gather ring, last-CTA publish, bounded wait, copy), and across GPUs under torchrun when the box has more than one."""

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_world_of_one(cuda_device, tmp_path):
    import torch
    import torch.distributed as dist

    from tactics2d_b200 import BatchedWorld, synthetic
    from tactics2d_b200.distributed import PeerDoneExchange

    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/rdv", rank=0, world_size=1)
    try:
        n, m = 100, 20
        scene = synthetic.config2(n, m, seed=5, size=50.0)
        a = BatchedWorld(n, m, scene.table, device=cuda_device, max_step=4)
        b = BatchedWorld(n, m, scene.table, device=cuda_device, max_step=4)
        for w in (a, b):
            w.set_map(scene.segments, scene.bounds)
            w.set_state(scene.x, scene.y, scene.heading, scene.speed, type_id=scene.type_id)
        ex = PeerDoneExchange(n, cuda_device, slots=4)
        ex.attach(a)
        total = 0
        for t in range(11):              # more steps than slots: the ring wraps
            act = torch.from_numpy(synthetic.random_actions(t, (n, m))).to(cuda_device)
            ra, rb = a.step(act), b.step(act)
            got = ex.gather()
            torch.cuda.synchronize()
            assert torch.equal(got[:n], rb.done) and torch.equal(ra.done, rb.done)      # same tick with and without the exchange
            assert torch.equal(ra.flags, rb.flags)
            total += int(rb.done.sum().item())
        assert total > 0
        assert ex.status() == (11, 11, 0)
        # a gather with nothing published runs into its bound instead of hanging the GPU
        ex.gather()
        torch.cuda.synchronize()
        assert ex.status() == (11, 11, 1)
        ex.detach(a)
        ex.close()
    finally:
        if own_group:
            dist.destroy_process_group()


def test_exchange_across_gpus():
    import torch

    k = min(torch.cuda.device_count(), 4)
    if k < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={k}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_exchange.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
