"""``PhysicsModelBase`` - the interface of the reference's ``tactics2d/physics/physics_model_base.py:12-72``
(``step``, ``verify_state``, ``verify_states``; constants ``_DELTA_T = 5``, ``_MIN_DELTA_T = 1``, ``_G = 9.81``)
plus the batched entry point every concrete model here adds: ``step_batch`` on CUDA tensors."""

from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod

from ..participant.trajectory import State, Trajectory


class PhysicsModelBase(ABC):
    _DELTA_T: int = 5
    _MIN_DELTA_T: int = 1
    _G = 9.81

    @abstractmethod
    def step(self, state: State, action: tuple, interval: int = None) -> State:
        """Advance one participant by ``interval`` ms."""

    @abstractmethod
    def verify_state(self, state: State, last_state: State, interval: int = None) -> bool:
        """Rough reachability check of a transition."""

    def verify_states(self, trajectory: Trajectory) -> bool:
        """physics_model_base.py:53-72: every consecutive pair of the trajectory must verify."""
        frames = trajectory.frames
        if len(frames) < 2:
            return True
        fixed = 1000 / trajectory.fps if (trajectory.stable_freq is True and trajectory.fps) else None
        last = trajectory.history_states[frames[0]]
        for frame in frames[1:]:
            state = trajectory.history_states[frame]
            interval = fixed if fixed is not None else state.frame - last.frame
            if self.verify_state(state, last, interval) is False:
                return False
        return True

    # ------------------------------------------------------------------ shared plumbing
    def _effective_delta_t(self, delta_t, interval):
        """Constructor rule (single_track_kinematics.py:119-124)."""
        if delta_t is None:
            return self._DELTA_T
        d = max(delta_t, self._MIN_DELTA_T)
        if interval is not None:
            d = min(d, interval)
        return d

    def _launch(self, params, interval, n, x, y, heading, speed, vx, vy, action, applied, omega_f=None, omega_r=None):
        """One call of the C ABI's ``t2d_physics_step`` on the tensors' device and current stream."""
        import torch

        from .. import _lib

        lib = _lib.load()
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("tactics2d_b200 physics runs on a CUDA device only (there is no CPU implementation)")
        extra = tuple(t for t in (applied, omega_f, omega_r) if t is not None)
        for t in (x, y, heading, speed, vx, vy, action) + extra:
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("state/action tensors must be contiguous fp32 tensors on one CUDA device")
        c = params.to_c()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        _lib.check(lib.t2d_physics_step(dev.index if dev.index is not None else torch.cuda.current_device(), C.byref(c),
                                        int(interval), int(self.delta_t), int(n), p(x), p(y), p(heading), p(speed),
                                        p(vx), p(vy), p(omega_f), p(omega_r), p(action), p(applied), stream))
