"""``Vehicle`` - attributes, defaults, template loading, default physics model and ``get_pose`` of the
reference's ``tactics2d/participant/element/vehicle.py:21-308``."""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from ...types import SHAPE_OBB, TypeParams
from ..trajectory import State, Trajectory
from .participant_base import ParticipantBase, box_corners, transform_box
from .participant_template import EPA_MAPPING, EURO_SEGMENT_MAPPING, NCAP_MAPPING, VEHICLE_TEMPLATE


class Vehicle(ParticipantBase):
    __annotations__ = {"type_": str, "length": float, "width": float, "height": float, "kerb_weight": float,
                       "wheel_base": float, "front_overhang": float, "rear_overhang": float, "driven_mode": str,
                       "max_steer": float, "max_speed": float, "max_accel": float, "max_decel": float, "verify": bool}
    _default_color = "#2bcbba"
    _driven_modes = {"FWD", "RWD", "4WD", "AWD"}

    def __init__(self, id_: Any, type_: str = "medium_car", trajectory: Trajectory = None, **kwargs):
        super().__init__(id_, type_, trajectory, **kwargs)
        self.max_steer = np.round(np.pi / 6, 3) if self.max_steer is None else self.max_steer   # :107
        self.max_speed = 55.56 if self.max_speed is None else self.max_speed
        self.max_accel = 3.0 if self.max_accel is None else self.max_accel
        self.max_decel = 10.0 if self.max_decel is None else self.max_decel
        self.speed_range = (-16.67, self.max_speed)                                               # :111-113
        self.steer_range = (-self.max_steer, self.max_steer)
        self.accel_range = (-self.max_accel, self.max_accel)
        if self.driven_mode is None:
            self.driven_mode = "FWD"
        elif self.driven_mode not in self._driven_modes:
            self.driven_mode = "FWD"
            logging.warning("Invalid driven mode. The default mode FWD will be used.")
        if self.verify:
            if kwargs.get("physics_model") is not None:
                self.physics_model = kwargs["physics_model"]
            else:
                self._auto_construct_physics_model()
        self._bbox = box_corners(self.length, self.width) if None not in [self.length, self.width] else None

    @property
    def geometry(self):
        return self._bbox

    def _auto_construct_physics_model(self):
        """FWD only (vehicle.py:144-176): kinematic bicycle about the geometry centre."""
        from ...physics import SingleTrackKinematics

        if self.driven_mode != "FWD":
            return
        kw = dict(steer_range=self.steer_range, speed_range=self.speed_range, accel_range=self.accel_range)
        if None not in [self.front_overhang, self.rear_overhang]:
            self.physics_model = SingleTrackKinematics(lf=self.length / 2 - self.front_overhang,
                                                       lr=self.length / 2 - self.rear_overhang, **kw)
        elif self.length is not None:
            self.physics_model = SingleTrackKinematics(lf=self.length / 2, lr=self.length / 2, **kw)
        else:
            self.verify = False
            logging.info("Cannot construct a physics model for the vehicle. The state verification is turned off.")

    def load_from_template(self, type_name: str, overwrite: bool = True, template: dict = None):
        """vehicle.py:179-221, incl. the alias maps and ``max_accel = round(27.78 / t_0_100, 3)``."""
        template = VEHICLE_TEMPLATE if template is None else template
        for alias in (EURO_SEGMENT_MAPPING, EPA_MAPPING, NCAP_MAPPING):
            if type_name in alias:
                type_name = alias[type_name]
                break
        if type_name in template:
            for key, value in template[type_name].items():
                if key == "0_100_km/h":
                    if overwrite or self.max_accel is None:
                        self.max_accel = np.round(100 * 1000 / 3600 / value, 3)
                elif overwrite or getattr(self, key) is None:
                    setattr(self, key, value)
        else:
            logging.warning(f"{type_name} is not in the vehicle template. The default values will be used.")
        self.speed_range = (-16.67, self.max_speed)
        self.accel_range = (-self.max_decel, self.max_accel)
        if None not in [self.length, self.width]:
            self._bbox = box_corners(self.length, self.width)

    def add_state(self, state: State):
        if not self.verify or self.physics_model is None:
            self.trajectory.add_state(state)
        elif self.physics_model.verify_state(state, self.trajectory.current_state):
            self.trajectory.append_state(state)
        else:
            raise RuntimeError("Invalid state checked by the physics model %s." % (self.physics_model.__class__.__name__))

    def bind_trajectory(self, trajectory: Trajectory):
        if not isinstance(trajectory, Trajectory):
            raise TypeError("The trajectory must be an instance of Trajectory.")
        if self.verify and not self._verify_trajectory(trajectory):
            self.trajectory = Trajectory(self.id_)
            logging.warning(f"The trajectory is invalid. Vehicle {self.id_} is not bound to the trajectory.")
        else:
            self.trajectory = trajectory

    def get_pose(self, frame: int = None) -> np.ndarray:
        """The bounding box rotated and moved to the state (vehicle.py:263-281): (4, 2) corners."""
        return transform_box(self._bbox, self.trajectory.get_state(frame))

    def type_params(self, model: str = None) -> TypeParams:
        """Row of the kernels' type table equivalent to this vehicle (+ its physics model)."""
        pm = self.physics_model
        if pm is None:
            self._auto_construct_physics_model()
            pm = self.physics_model
        base = pm.type_params(half_len=self.length / 2, half_wid=self.width / 2, shape=SHAPE_OBB)
        base.name = self.type_
        return base
