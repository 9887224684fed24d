"""``Other`` (reference ``tactics2d/participant/element/other.py:19-149``): a participant of unknown type whose
shape is a box when length/width are known (a square when only one is), else a point."""

from __future__ import annotations

from typing import Any

import numpy as np

from ...types import MODEL_STATIC, SHAPE_NONE, SHAPE_OBB, TypeParams
from ..trajectory import State, Trajectory
from .participant_base import ParticipantBase, box_corners, transform_box


class Other(ParticipantBase):
    def __init__(self, id_: Any, type_: str = "unknown", trajectory: Trajectory = None, **kwargs):
        super().__init__(id_, type_, trajectory, **kwargs)

    @property
    def geometry(self):
        if self.length is not None and self.width is not None:
            return box_corners(self.length, self.width)
        if self.length is not None:
            return box_corners(self.length, self.length)
        if self.width is not None:
            return box_corners(self.width, self.width)
        return None

    def _verify_state(self, state: State) -> bool:
        return True

    def _verify_trajectory(self, trajectory: Trajectory) -> bool:
        return True

    def bind_trajectory(self, trajectory: Trajectory = None):
        if not isinstance(trajectory, Trajectory):
            raise TypeError(f"Expected a trajectory of type 'Trajectory', but got {type(trajectory)}.")
        self.trajectory = trajectory

    def get_pose(self, frame: int = None):
        geometry = self.geometry
        state = self.trajectory.get_state(frame)
        if geometry is None:
            return np.asarray(state.location, dtype=np.float64)   # the reference returns a shapely Point
        return transform_box(geometry, state)

    def type_params(self) -> TypeParams:
        g = self.geometry
        if g is None:
            return TypeParams(model=MODEL_STATIC, shape=SHAPE_NONE, name=self.type_)
        return TypeParams(half_len=float(g[0, 0]), half_wid=float(g[1, 1]), model=MODEL_STATIC, shape=SHAPE_OBB, name=self.type_)
