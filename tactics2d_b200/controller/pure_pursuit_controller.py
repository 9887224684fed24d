"""``PurePursuitController`` (tactics2d/controller/pure_pursuit_controller.py:14-98): pure-pursuit steering on a
waypoint polyline + the ``AccelerationController`` laws for the longitudinal command.

``waypoints`` is an array-like ``[V, 2]`` of vertices (or anything with ``.coords``, such as a shapely ``LineString``).
As in the reference (:90-92) the look-ahead point is ``waypoints.interpolate(max(speed * interval, min_pre_aiming_distance))``
- arc length measured from the path's first vertex."""

from __future__ import annotations

import numpy as np

from .. import _lib
from ..participant.trajectory import State
from .acceleration_controller import AccelerationController
from .controller_base import CTRL_PURE_PURSUIT, ControllerBase


class PurePursuitController(ControllerBase):
    interval = 1.0

    def __init__(self, min_pre_aiming_distance: float = 10.0, target_speed: float = 5.0):
        if min_pre_aiming_distance <= 0:
            raise ValueError("min_pre_aiming_distance must be positive")
        if target_speed < 0:
            raise ValueError("target_speed must be non-negative")
        self.min_pre_aiming_distance = min_pre_aiming_distance
        self._interval_interpolator = self.create_style_interpolator(2.0, 1.0)
        self._longitudinal_control = AccelerationController(target_speed)
        self.wheel_base = 2.637   # default of `step`, :76 (medium car)

    def update_driving_style(self, style_id: float):
        if not isinstance(style_id, (int, float)):
            raise TypeError("style_id must be int or float")
        self._longitudinal_control.update_driving_style(style_id)
        self.interval = self._interval_interpolator(style_id)

    def params(self):
        row = self._longitudinal_control._fill(_lib.ControllerParamsC(kind=CTRL_PURE_PURSUIT))
        row.min_pre_aiming_distance, row.pp_interval, row.wheel_base = self.min_pre_aiming_distance, self.interval, self.wheel_base
        return row

    def step(self, ego_state, waypoints, wheel_base: float = 2.637, **kwargs):
        """``(steering, accel)`` (pure_pursuit_controller.py:76-98)."""
        front_state = kwargs.get("front_state")
        if front_state is not None and not isinstance(front_state, State):
            raise TypeError("front_state must be a State instance")
        path = np.asarray(waypoints.coords if hasattr(waypoints, "coords") else waypoints, dtype=np.float32).reshape(-1, 2)
        return self._step_one(ego_state, front_state, path=path, wheel_base=wheel_base)
