"""``Cyclist`` (reference ``tactics2d/participant/element/cyclist.py:21-188``): template-loaded dimensions,
ranges steer +-max_steer / speed (0, max_speed) / accel (-max_decel, max_accel), kinematic bicycle with
lf = lr = L/2 when ``verify``, box pose like ``Vehicle``."""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from ...types import SHAPE_OBB, TypeParams
from ..trajectory import Trajectory
from .participant_base import ParticipantBase, box_corners, transform_box
from .participant_template import CYCLIST_TEMPLATE


class Cyclist(ParticipantBase):
    __annotations__ = {"type_": str, "length": float, "width": float, "height": float, "max_steer": float,
                       "max_speed": float, "max_accel": float, "max_decel": float, "verify": bool}
    _default_color = "#45aaf2"

    def __init__(self, id_: Any, type_: str = "cyclist", trajectory: Trajectory = None, **kwargs):
        super().__init__(id_, type_, trajectory, **kwargs)
        self.load_from_template(type_ if type_ in CYCLIST_TEMPLATE else "cyclist")
        self.steer_range = (-self.max_steer, self.max_steer)
        self.speed_range = (0, self.max_speed)
        self.accel_range = (-self.max_decel, self.max_accel)
        if not self.verify:
            self.physics_model = None
        elif kwargs.get("physics_model") is None:
            self.physics_model = self._default_model()
        else:
            self.physics_model = kwargs["physics_model"]
        self._bbox = box_corners(self.length, self.width)

    def _default_model(self):
        from ...physics import SingleTrackKinematics

        return SingleTrackKinematics(lf=self.length / 2, lr=self.length / 2, steer_range=self.steer_range,
                                     speed_range=self.speed_range, accel_range=self.accel_range)

    @property
    def geometry(self):
        return self._bbox

    def load_from_template(self, type_name: str, overwrite: bool = True, template: dict = None):
        template = CYCLIST_TEMPLATE if template is None else template
        if type_name in template:
            for key, value in template[type_name].items():
                if getattr(self, key, None) is None or overwrite:
                    setattr(self, key, value)
        else:
            logging.warning(f"{type_name} is not in the cyclist template. Cannot auto-complete the empty attributes")

    def bind_trajectory(self, trajectory: Trajectory):
        if not isinstance(trajectory, Trajectory):
            raise TypeError("The trajectory must be an instance of Trajectory.")
        if self.verify and not self._verify_trajectory(trajectory):
            self.trajectory = Trajectory(self.id_)
            logging.warning(f"The trajectory is invalid. Cyclist {self.id_} is not bound to the trajectory.")
        else:
            self.trajectory = trajectory

    def get_pose(self, frame: int = None) -> np.ndarray:
        return transform_box(self._bbox, self.trajectory.get_state(frame))

    def type_params(self) -> TypeParams:
        pm = self.physics_model or self._default_model()
        row = pm.type_params(half_len=self.length / 2, half_wid=self.width / 2, shape=SHAPE_OBB)
        row.name = self.type_
        return row
