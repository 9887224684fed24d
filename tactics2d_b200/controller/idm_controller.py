"""``IDMController`` (tactics2d/controller/idm_controller.py:16-155): Intelligent Driver Model, longitudinal only."""

from __future__ import annotations

from typing import Tuple

from .. import _lib
from .controller_base import CTRL_IDM, ControllerBase


class IDMController(ControllerBase):
    def __init__(self, desired_speed: float = 10.0, time_headway: float = 1.5, min_spacing: float = 2.0,
                 max_acceleration: float = 1.0, comfortable_deceleration: float = 3.0, delta: float = 4.0):
        self.desired_speed = desired_speed
        self.time_headway = time_headway
        self.min_spacing = min_spacing
        self.max_acceleration = max_acceleration
        self.comfortable_deceleration = comfortable_deceleration
        self.delta = delta

    def params(self):
        return _lib.ControllerParamsC(kind=CTRL_IDM, desired_speed=self.desired_speed, time_headway=self.time_headway,
                                      min_spacing=self.min_spacing, max_acceleration=self.max_acceleration,
                                      comfortable_deceleration=self.comfortable_deceleration, delta=self.delta)

    def step(self, ego_state, leading_state=None, **kwargs) -> Tuple[float, float]:
        """``(0.0, acceleration)``: free flow without a leader, car following with one (idm_controller.py:59-92)."""
        _, accel = self._step_one(ego_state, leading_state)
        return 0.0, accel

    def configure(self, **kwargs) -> None:
        for key, value in kwargs.items():
            if hasattr(self, key):
                setattr(self, key, value)
            else:
                raise AttributeError(f"IDMController has no parameter '{key}'")
