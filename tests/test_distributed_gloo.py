"""N > 1 host logic on CPU: scenario sharding and the done-mask exchange, world_size 2 over gloo."""

import os
import socket

import numpy as np
import pytest

from tactics2d_b200.distributed import shard_range, shard_sizes


def test_shard_ranges_cover_and_never_split():
    for n, w in [(4096, 8), (4096, 3), (7, 2), (5, 8), (16384, 8)]:
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        sizes = shard_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist

    from oracle import scenario as O
    from tactics2d_b200 import synthetic
    from tactics2d_b200.distributed import DoneExchange, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = synthetic.config2(n_total, 16, seed=5, size=40.0)      # every rank builds the same global scene
        lo, hi = shard_range(n_total, rank, world)
        table = scene.table.as_oracle_table()
        act = synthetic.random_actions(3, scene.shape)
        # this rank's shard of the tick (oracle stands in for the kernel: the host logic is what is under test)
        st = {k: v[lo:hi] for k, v in scene.state().items()}
        _, fl, _, _, _, done, _ = O.tick(st, scene.type_id[lo:hi], act[lo:hi], table, np.zeros(hi - lo, np.int32), scene.segments,
                                         scene.bounds, ego_only=False)
        ex = DoneExchange(n_total, torch.device("cpu"))
        got = ex(torch.from_numpy(done)).clone().numpy()
        got2 = ex(torch.from_numpy(done)).numpy()                    # second call reuses the buffers
        q.put((rank, got, got2, done, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [12, 13])
def test_done_mask_all_gather_world2(n_total):
    import torch.multiprocessing as mp

    from oracle import scenario as O
    from tactics2d_b200 import synthetic

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth over the whole batch
    scene = synthetic.config2(n_total, 16, seed=5, size=40.0)
    act = synthetic.random_actions(3, scene.shape)
    _, _, _, _, _, done_all, _ = O.tick(scene.state(), scene.type_id, act, scene.table.as_oracle_table(), np.zeros(n_total, np.int32),
                                        scene.segments, scene.bounds, ego_only=False)
    assert done_all.any() and not done_all.all()
    for rank, got, got2, local, (lo, hi) in results:
        assert np.array_equal(got, done_all) and np.array_equal(got2, done_all)
        assert np.array_equal(local, done_all[lo:hi])


class _FakeExchangeLib:
    """Stands in for libt2d_b200's t2d_exchange_* entry points (no GPU in this suite): records calls, can fail on demand."""

    def __init__(self, rank, fail_connect_on=None, fail_create_on=None):
        self.rank, self.fail_connect_on, self.fail_create_on = rank, fail_connect_on, fail_create_on
        self.connected_with = None
        self.destroyed = False

    def t2d_exchange_create(self, xref, device, world, rank, n_local, slots, handle):
        import ctypes as C

        if self.fail_create_on == rank:
            return -2
        xref._obj.value = 4242
        C.memmove(handle.value, bytes([rank]) * 64, 64)
        return 0

    def t2d_exchange_connect(self, x, blob):
        import ctypes as C

        if self.fail_connect_on == self.rank:
            return -2
        self.connected_with = C.string_at(blob.value, 128)
        return 0

    def t2d_exchange_destroy(self, x):
        self.destroyed = True
        return 0


def _exchange_setup_worker(rank, world, port, mode, q):
    import torch
    import torch.distributed as dist

    from tactics2d_b200.distributed import PeerDoneExchange

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = _FakeExchangeLib(rank, fail_connect_on=1 if mode == "connect" else None, fail_create_on=0 if mode == "create" else None)
        try:
            ex = PeerDoneExchange(100, torch.device("cpu"), slots=2, lib=lib)
            q.put((rank, "ok", lib.connected_with == bytes([0]) * 64 + bytes([1]) * 64, ex.pad, tuple(ex.out.shape)))
        except RuntimeError as e:
            q.put((rank, "raised", str(e), lib.destroyed, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ok", "connect", "create"])
def test_peer_exchange_setup_is_agreed_on_by_all_ranks(mode):
    """The CUDA IPC set-up of PeerDoneExchange either succeeds on every rank or raises on every rank - a rank whose
    create / connect fails never leaves the others waiting in a collective (bench.py then falls back to NCCL)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_setup_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if mode == "ok":
        assert [r[1:] for r in results] == [("ok", True, 112, (224,))] * 2      # handles in rank order, rows padded to 16
    else:
        assert all(r[1] == "raised" and "unavailable" in r[2] for r in results)
        bad = 1 if mode == "connect" else 0
        assert all(f"{bad}:" in r[2] for r in results)


def test_peer_exchange_lag_arguments_are_validated_before_any_device_call():
    """A lag of k steps needs a ring of 2 k + 2 slots; negative lags are rejected (host logic, no GPU)."""
    import torch.distributed as dist

    from tactics2d_b200.distributed import PeerDoneExchange

    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError):
            PeerDoneExchange(64, "cpu", lag=-1, lib=_FakeExchangeLib(0))
        with pytest.raises(ValueError):
            PeerDoneExchange(64, "cpu", slots=4, lag=2, lib=_FakeExchangeLib(0))
        ex = PeerDoneExchange(64, "cpu", lag=2, lib=_FakeExchangeLib(0))
        assert ex.slots >= 6 and ex.lag == 2 and ex.calls == 0
        # the kernel writes world * pad bytes: a shorter / non-uint8 destination or a wrong-sized mask never reaches it
        import torch

        ex = PeerDoneExchange(100, "cpu", lib=_FakeExchangeLib(0))
        assert ex.pad == 112
        for mask, out in ((torch.zeros(100, dtype=torch.uint8), torch.zeros(100, dtype=torch.uint8)),
                          (torch.zeros(100, dtype=torch.uint8), torch.zeros(112, dtype=torch.int8)),
                          (torch.zeros(99, dtype=torch.uint8), None),
                          (torch.zeros(100, dtype=torch.bool), None)):
            with pytest.raises(ValueError):
                ex(mask, out)
        assert ex.calls == 0
    finally:
        if own:
            dist.destroy_process_group()
