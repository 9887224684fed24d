"""Scenario-axis sharding across ranks (one process per GPU).

Scenarios never interact (the reference runs one ``ScenarioManager`` per env, scenario_manager.py:52-61), so
the N scenarios are split into contiguous blocks, one per rank, with no data-path collective inside the tick.
The single exchange per step is the all-gather of the ``uint8 done`` mask (a centralised learner / reset
scheduler needs every rank's mask).  Works on any ``torch.distributed`` backend: NCCL on the GPUs, gloo in the
CPU tests of the host logic.
"""

from __future__ import annotations

from typing import Optional, Tuple


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of scenarios owned by ``rank``; block sizes differ by at most one and the
    participants of a scenario are never split."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_total: int, world_size: int):
    return [shard_range(n_total, r, world_size)[1] - shard_range(n_total, r, world_size)[0] for r in range(world_size)]


class DoneExchange:
    """All-gather of the per-rank ``done`` masks into one [N_total] tensor on every rank.

    Equal shards use ``all_gather_into_tensor`` (one NCCL call, graph-capturable); unequal shards pad to the
    largest shard.  ``group`` is a ``torch.distributed`` process group (default: WORLD)."""

    def __init__(self, n_total: int, device, group=None):
        import torch
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_total = n_total
        self.sizes = shard_sizes(n_total, self.world_size)
        self.lo, self.hi = shard_range(n_total, self.rank, self.world_size)
        self.equal = len(set(self.sizes)) == 1
        self.pad = max(self.sizes)
        self._gathered = torch.zeros(self.world_size * self.pad, dtype=torch.uint8, device=device)
        self._send = torch.zeros(self.pad, dtype=torch.uint8, device=device)
        self.out = self._gathered if self.equal else torch.zeros(n_total, dtype=torch.uint8, device=device)

    def __call__(self, done_local):
        """``done_local``: uint8 [hi - lo] on this rank -> uint8 [N_total] (same tensor object every call)."""
        if done_local.numel() != self.hi - self.lo:
            raise ValueError("done_local does not match this rank's shard")
        if self.equal:
            self.dist.all_gather_into_tensor(self._gathered, done_local, group=self.group)
            return self._gathered
        self._send[: done_local.numel()].copy_(done_local)
        self.dist.all_gather_into_tensor(self._gathered, self._send, group=self.group)
        off = 0
        for r, sz in enumerate(self.sizes):
            self.out[off:off + sz].copy_(self._gathered[r * self.pad:r * self.pad + sz])
            off += sz
        return self.out


class PeerDoneExchange:
    """The same exchange without NCCL: one small kernel per rank and step stores the rank's mask into a ring slot on
    every rank (peer stores over NVLink / NVSwitch), signals the step on every rank's flag word, waits for all ranks'
    signals and copies the slot out (``t2d_exchange_*`` in ``include/t2d_b200.h``).  Equal shards, one node.
    ``torch.distributed`` is used once, to pass the CUDA IPC handles around.  Call it like ``DoneExchange``:
    ``done_all = exchange(out.done)`` on every rank, the same number of times, in stream order.

    ``lag`` > 0 takes the exchange off the critical path: call k posts step k's mask and returns the gathered masks of
    step k - lag (the first ``lag`` calls leave the output untouched); nothing waits for the slowest rank's current
    tick any more.  ``slots`` defaults to the 2 * lag + 2 the ring needs."""

    def __init__(self, n_local: int, device, slots: Optional[int] = None, group=None, lib=None, lag: int = 0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        self.lib = lib if lib is not None else _lib.load()
        self.device = torch.device(device)
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_local = int(n_local)
        self.pad = (self.n_local + 15) & ~15
        self.lag = int(lag)
        if self.lag < 0:
            raise ValueError("lag must be >= 0")
        self.slots = int(slots) if slots is not None else max(4, 2 * self.lag + 2)
        if self.slots < 2 * self.lag + 2:
            raise ValueError("a lag of k steps needs a ring of at least 2 k + 2 slots")
        self.calls = 0
        dev_index = self.device.index
        if dev_index is None:   # an un-indexed 'cuda' device is the CURRENT device, not GPU 0
            dev_index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._x = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        # Every step of the set-up is agreed on by all ranks before anybody proceeds: a rank whose CUDA IPC call fails
        # (no peer access, a container without shared IPC namespaces) must not leave the others waiting in a collective.
        problem = None
        try:
            _lib.check(self.lib.t2d_exchange_create(C.byref(self._x), dev_index, self.world_size, self.rank, self.n_local,
                                                    self.slots, C.cast(handle, C.c_void_p)))
        except Exception as e:   # noqa: BLE001 - reported to every rank below
            problem = f"create: {e}"
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, (problem, bytes(handle)), group=group)
        if all(p is None for p, _ in everyone):
            blob = (C.c_ubyte * (64 * self.world_size)).from_buffer_copy(b"".join(h for _, h in everyone))
            try:
                _lib.check(self.lib.t2d_exchange_connect(self._x, C.cast(blob, C.c_void_p)))
            except Exception as e:   # noqa: BLE001
                problem = f"connect: {e}"
        verdicts = [None] * self.world_size
        dist.all_gather_object(verdicts, problem if problem is not None else everyone[self.rank][0], group=group)
        failed = {r: v for r, v in enumerate(verdicts) if v is not None}
        failed.update({r: p for r, (p, _) in enumerate(everyone) if p is not None})
        if failed:
            self.close()
            raise RuntimeError(f"peer-memory done exchange unavailable (rank -> reason): {failed}")
        self.out = torch.zeros(self.world_size * self.pad, dtype=torch.uint8, device=self.device)

    def __call__(self, done_local, out=None):
        """``done_local``: uint8 [n_local] on this rank -> uint8 [world * pad] (row r = rank r's mask, zero padded) of the
        step ``lag`` calls ago (this step's with ``lag`` = 0); enqueued on the current stream.  A row of 0xFF bytes
        means a rank failed to signal within the time-out (``status()`` then reports ``timed_out``)."""
        import ctypes as C

        import torch

        from . import _lib

        if done_local.numel() != self.n_local or done_local.dtype != torch.uint8 or not done_local.is_contiguous():
            raise ValueError("done_local must be a contiguous uint8 tensor of this rank's scenarios")
        out = self.out if out is None else out
        if out.numel() < self.world_size * self.pad or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != done_local.device:
            raise ValueError(f"out must be a contiguous uint8 tensor of at least world * pad = {self.world_size * self.pad} bytes on the same device")
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.t2d_exchange_allgather_lagged(self._x, C.c_void_p(done_local.data_ptr()), C.c_void_p(out.data_ptr()),
                                                          self.lag, stream))
        self.calls += 1
        return out

    def status(self):
        """(steps exchanged, timed_out) of this rank (synchronises the device)."""
        import ctypes as C

        from . import _lib

        a, b = C.c_uint32(), C.c_uint32()
        _lib.check(self.lib.t2d_exchange_status(self._x, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "_x", None) is not None and self._x.value:
            self.lib.t2d_exchange_destroy(self._x)
            self._x.value = None
