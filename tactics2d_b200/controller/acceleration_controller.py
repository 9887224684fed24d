"""``AccelerationController`` (tactics2d/controller/acceleration_controller.py:14-145): proportional cruise control,
adaptive cruise control when a ``front_state`` is given."""

from __future__ import annotations

from .. import _lib
from ..participant.trajectory import State
from .controller_base import CTRL_CRUISE, ControllerBase


class AccelerationController(ControllerBase):
    kp = 3.5
    speed_factor = 1.0
    accel_change_rate = 3.0
    max_accel = 1.5
    min_accel = -4.0
    interval = 2.0
    delta_t = 0.05

    DEFAULT_SAFETY_DISTANCE = 5.0
    MIN_TARGET_DISTANCE = 7.0
    MAX_TARGET_DISTANCE = 80.0

    def __init__(self, target_speed: float = 5.0):
        if target_speed < 0:
            raise ValueError("target_speed must be non-negative")
        self.target_speed = target_speed
        self._kp_interpolator = self.create_style_interpolator(4.5, 2.5)
        self._speed_factor_interpolator = self.create_style_interpolator(0.8, 1.2)
        self._accel_change_rate_interpolator = self.create_style_interpolator(3.0, 6.0)
        self._max_accel_interpolator = self.create_style_interpolator(1.5, 2.5)
        self._min_accel_interpolator = self.create_style_interpolator(-3.0, -5.0)
        self._interval_interpolator = self.create_style_interpolator(3.5, 1.5)

    def update_driving_style(self, style_id: float):
        if not isinstance(style_id, (int, float)):
            raise TypeError("style_id must be int or float")
        self.kp = self._kp_interpolator(style_id)
        self.speed_factor = self._speed_factor_interpolator(style_id)
        self.accel_change_rate = self._accel_change_rate_interpolator(style_id)
        self.max_accel = self._max_accel_interpolator(style_id)
        self.min_accel = self._min_accel_interpolator(style_id)
        self.interval = self._interval_interpolator(style_id)

    def _fill(self, row):
        row.target_speed, row.kp, row.accel_change_rate = self.target_speed, self.kp, self.accel_change_rate
        row.delta_t, row.max_accel, row.min_accel, row.interval = self.delta_t, self.max_accel, self.min_accel, self.interval
        return row

    def params(self):
        return self._fill(_lib.ControllerParamsC(kind=CTRL_CRUISE))

    def step(self, ego_state, **kwargs):
        """``(0.0, accel)`` (acceleration_controller.py:132-145)."""
        front_state = kwargs.get("front_state")
        if front_state is not None and not isinstance(front_state, State):
            raise TypeError("front_state must be a State instance")
        _, accel = self._step_one(ego_state, front_state)
        return 0.0, accel
