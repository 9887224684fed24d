"""``Pedestrian`` (reference ``tactics2d/participant/element/pedestrian.py:19-162``): PointMass physics with
``speed_range = (-max_speed, max_speed)`` (which PointMass normalises to [0, max_speed]), a disc of radius
``width / 2`` as pose: ``get_pose`` returns ``((x, y), radius)`` (:138-149)."""

from __future__ import annotations

import logging
from typing import Any, Tuple

from ...types import SHAPE_CIRCLE, TypeParams
from ..trajectory import Trajectory
from .participant_base import ParticipantBase
from .participant_template import PEDESTRIAN_TEMPLATE


class Pedestrian(ParticipantBase):
    __annotations__ = {"type_": str, "length": float, "width": float, "height": float, "max_speed": float,
                       "max_accel": float, "verify": bool}
    _default_color = "#fd9644"

    def __init__(self, id_: Any, type_: str = "adult_male", trajectory: Trajectory = None, **kwargs):
        from ...physics import PointMass

        super().__init__(id_, type_, trajectory, **kwargs)
        self.load_from_template(type_ if type_ in PEDESTRIAN_TEMPLATE else "adult_male")
        self.speed_range = (-self.max_speed, self.max_speed)
        self.accel_range = (-self.max_accel, self.max_accel)
        if kwargs.get("physics_model") is None:
            self.physics_model = PointMass(speed_range=self.speed_range, accel_range=self.accel_range)
        else:
            self.physics_model = kwargs["physics_model"]
        self._radius = self.width / 2 if getattr(self, "width", None) is not None else 0

    @property
    def geometry(self) -> float:
        return self._radius

    def load_from_template(self, type_name: str, overwrite: bool = True, template: dict = None):
        template = PEDESTRIAN_TEMPLATE if template is None else template
        if type_name in template:
            for key, value in template[type_name].items():
                if overwrite or getattr(self, key) is None:
                    setattr(self, key, value)
        else:
            logging.warning(f"{type_name} is not in the template. Cannot auto-complete the empty attributes")

    def bind_trajectory(self, trajectory: Trajectory):
        if not isinstance(trajectory, Trajectory):
            raise TypeError("The trajectory must be an instance of Trajectory.")
        if self.verify and not self._verify_trajectory(trajectory):
            self.trajectory = Trajectory(self.id_)
            logging.warning(f"The trajectory is invalid. Pedestrian {self.id_} is not bound to the trajectory.")
        else:
            self.trajectory = trajectory

    def get_pose(self, frame: int = None) -> Tuple[Tuple[float, float], float]:
        return (self.trajectory.get_state(frame).location, self._radius)

    def type_params(self) -> TypeParams:
        row = self.physics_model.type_params(half_len=self.length / 2, half_wid=self.width / 2, radius=self._radius,
                                             shape=SHAPE_CIRCLE)
        row.name = self.type_
        return row
