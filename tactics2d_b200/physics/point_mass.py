"""``PointMass`` - point-mass model for pedestrians.

Constructor, range rules, backends and ``step(state, (ax, ay), interval) -> State`` follow the reference's
``tactics2d/physics/point_mass.py`` (:33-81, :209-232; ``newton`` :83-175, ``euler`` :177-207,
``verify_state`` :234-258).  Reference behaviour kept on purpose: the acceleration is *not* clipped
(``step`` computes a clipped magnitude at :222-225 and never uses it) and a tuple speed range is clamped
at zero (``(-7, 7)`` becomes ``[0, 7]``, :52-55).
"""

from __future__ import annotations

import logging
from typing import Tuple, Union

import numpy as np

from ..participant.trajectory import State
from ..types import MODEL_POINTMASS_EULER, MODEL_POINTMASS_NEWTON, SHAPE_NONE, TypeParams, normalize_range_pointmass
from .physics_model_base import PhysicsModelBase


def _store(rng):
    return None if rng[0] == -np.inf and rng[1] == np.inf else [rng[0], rng[1]]


class PointMass(PhysicsModelBase):
    backends = ["newton", "euler"]

    def __init__(self, speed_range: Union[float, Tuple[float, float]] = None,
                 accel_range: Union[float, Tuple[float, float]] = None, interval: int = 100, delta_t: int = None,
                 backend: str = "newton"):
        self._speed = normalize_range_pointmass(speed_range)
        self._accel = normalize_range_pointmass(accel_range)
        self.speed_range, self.accel_range = _store(self._speed), _store(self._accel)
        self.interval = interval
        self.delta_t = self._effective_delta_t(delta_t, interval)
        if backend not in self.backends:
            logging.warning(f"Unsupported backend {backend}. Using `newton` instead.")
            backend = "newton"
        self.backend = backend

    def type_params(self, **shape) -> TypeParams:
        kw = dict(speed_lo=self._speed[0], speed_hi=self._speed[1], accel_lo=self._accel[0], accel_hi=self._accel[1],
                  model=MODEL_POINTMASS_EULER if self.backend == "euler" else MODEL_POINTMASS_NEWTON, shape=SHAPE_NONE)
        kw.update(shape)
        return TypeParams(**kw)

    def step_batch(self, x, y, heading, vx, vy, ax, ay, interval: int = None):
        """n pedestrians at once; ``x, y, heading, vx, vy`` (fp32 CUDA tensors) are advanced IN PLACE.
        Returns the speed tensor."""
        import torch

        interval = interval if interval is not None else self.interval
        action = torch.stack([ax.reshape(-1), ay.reshape(-1)], dim=1).contiguous()
        speed = torch.empty_like(x)
        self._launch(self.type_params(), interval, x.numel(), x, y, heading, speed, vx, vy, action, None)
        return speed

    def step(self, state: State, accel: Tuple[float, float], interval: int = None) -> State:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("tactics2d_b200 physics needs a CUDA device (no CPU implementation)")
        interval = interval if interval is not None else self.interval
        dev = torch.device("cuda", torch.cuda.current_device())
        vx, vy = state.velocity
        buf = torch.tensor([[state.x], [state.y], [state.heading], [vx], [vy], [accel[0]], [accel[1]]], dtype=torch.float32,
                           device=dev)
        self.step_batch(buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], buf[6], interval)
        o = buf.reshape(-1).cpu().tolist()
        if self.backend == "euler":   # :203-205: the euler State also records (ax, ay)
            return State(frame=state.frame + interval, x=o[0], y=o[1], heading=o[2], vx=o[3], vy=o[4], ax=accel[0], ay=accel[1])
        return State(frame=state.frame + interval, x=o[0], y=o[1], heading=o[2], vx=o[3], vy=o[4])

    def verify_state(self, state: State, last_state: State, interval: int = None) -> bool:
        """point_mass.py:234-258: the constant acceleration implied by the displacement must lie in range."""
        interval = state.frame - last_state.frame if interval is None else interval
        if interval == 0:
            return True
        dt = interval / 1000
        k = 2 / dt**2
        ax = (state.x - last_state.x - last_state.vx * dt) * k
        ay = (state.y - last_state.y - last_state.vy * dt) * k
        if self.accel_range is not None:
            a = float(np.hypot(ax, ay))
            if not self.accel_range[0] <= a <= self.accel_range[1]:
                return False
        return True
