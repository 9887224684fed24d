"""The peer-memory done exchange (t2d_exchange_*): on one GPU with a world of one (the kernel's put / signal / wait /
copy on its own ring), and across GPUs under torchrun when the box has more than one."""

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

    from tactics2d_b200.distributed import PeerDoneExchange

    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/rdv", rank=0, world_size=1)
    try:
        n = 100                          # rows are padded to 112
        ex = PeerDoneExchange(n, cuda_device, slots=2)
        g = torch.Generator().manual_seed(0)
        for t in range(7):               # more steps than slots: the ring wraps
            done = (torch.rand(n, generator=g) < 0.3).to(torch.uint8).to(cuda_device)
            got = ex(done)
            torch.cuda.synchronize()
            assert got.numel() == ex.pad == 112
            assert torch.equal(got[:n], done) and int(got[n:].sum().item()) == 0
        assert ex.status() == (7, 0)
        with pytest.raises(ValueError):
            ex(torch.zeros(n + 1, dtype=torch.uint8, device=cuda_device))
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
