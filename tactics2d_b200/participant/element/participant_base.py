"""``ParticipantBase`` - the interface of the reference's
``tactics2d/participant/element/participant_base.py:14-246``: typed attributes coerced on assignment
(a value that cannot be converted becomes ``None`` with a warning, :80-103), a ``Trajectory``, an optional
physics model, ``get_pose / add_state / get_state(s) / is_active / reset``.

Poses are returned as plain arrays (shapely is not a dependency here): an ``(4, 2)`` float64 array of the
box corners in the reference's ring order for boxes, ``((x, y), radius)`` for pedestrians.  The batched
collision path never materialises these objects; ``type_params()`` is the bridge to the kernels' type table.
"""

from __future__ import annotations

import logging
from abc import ABC, abstractmethod
from typing import Any, List, Tuple

import numpy as np

from ..trajectory import State, Trajectory


class ParticipantBase(ABC):
    __annotations__ = {"type_": str, "length": float, "width": float, "height": float, "verify": bool}
    _default_color = (0, 0, 0, 255)

    def __init__(self, id_: Any, type_: str, trajectory: Trajectory = None, **kwargs):
        self.id_ = id_
        self.type_ = type_
        self.color = self._default_color if kwargs.get("color") is None else kwargs["color"]
        for key in self.__annotations__:
            if key != "type_":
                setattr(self, key, kwargs.get(key, None))
        self.verify = False if self.verify is None else self.verify
        self.physics_model = None
        if trajectory is not None:
            self.bind_trajectory(trajectory)
        else:
            self.trajectory = Trajectory(id_=self.id_)

    def __setattr__(self, name: str, value: Any):
        want = self.__annotations__.get(name)
        if want is not None and value is not None and not isinstance(value, want):
            try:
                value = want(value)
            except Exception:
                logging.warning(f"Cannot set {name} to {value}. Set to None instead.")
                value = None
        object.__setattr__(self, name, value)

    # ------------------------------------------------------------------ verification hooks
    def _verify_state(self, state: State) -> bool:
        return self.physics_model.verify_state(state, self.current_state) if self.verify else True

    def _verify_trajectory(self, trajectory: Trajectory) -> bool:
        return self.physics_model.verify_states(trajectory) if self.verify else True

    # ------------------------------------------------------------------ abstract surface
    @property
    @abstractmethod
    def geometry(self):
        """Local collision shape: (4, 2) corner array, a radius, or None."""

    @abstractmethod
    def bind_trajectory(self, trajectory: Trajectory = None):
        ...

    @abstractmethod
    def get_pose(self, frame: int = None):
        ...

    def get_trace(self, frame_range: Tuple[int, int] = None):
        """Centre-line points of the trajectory (the reference buffers them into a shapely ring)."""
        return self.trajectory.get_trace(frame_range)

    # ------------------------------------------------------------------ common behaviour
    @property
    def current_state(self) -> State:
        return self.trajectory.get_state()

    def is_active(self, frame: int) -> bool:
        if self.trajectory.first_frame is None:
            return False
        return self.trajectory.first_frame <= frame <= self.trajectory.last_frame

    def add_state(self, state: State):
        self.trajectory.add_state(state)

    def get_state(self, frame: int = None) -> State:
        if frame is None:
            return self.current_state
        if frame not in self.trajectory.history_states:
            raise KeyError(f"Time stamp {frame} is not found in the trajectory {self.id_}.")
        return self.trajectory.history_states[frame]

    def get_states(self, frame_range: Tuple[int] = None, frames: List[int] = None) -> List[State]:
        if frame_range is None and frames is None:
            return [self.trajectory.get_state(f) for f in self.trajectory.frames]
        if frame_range is not None:
            if len(frame_range) != 2:
                raise ValueError("The frame range must be a tuple with two elements.")
            lo, hi = frame_range
            return [self.trajectory.get_state(f) for f in self.trajectory.frames if lo <= f <= hi]
        return [self.trajectory.get_state(f) for f in sorted(frames)]

    def reset(self, state: State = None, keep_trajectory: bool = False):
        self.trajectory.reset(state, keep_trajectory)


def box_corners(length: float, width: float) -> np.ndarray:
    """Local ring [(+L/2,-W/2), (+L/2,+W/2), (-L/2,+W/2), (-L/2,-W/2)] (vehicle.py:133-140)."""
    hl, hw = 0.5 * length, 0.5 * width
    return np.array([[hl, -hw], [hl, hw], [-hl, hw], [-hl, -hw]], dtype=np.float64)


def transform_box(corners: np.ndarray, state: State) -> np.ndarray:
    """affine_transform(ring, [cos h, -sin h, sin h, cos h, x, y]) (vehicle.py:272-281)."""
    c, s = np.cos(state.heading), np.sin(state.heading)
    x = state.location[0] + corners[:, 0] * c - corners[:, 1] * s
    y = state.location[1] + corners[:, 0] * s + corners[:, 1] * c
    return np.stack([x, y], axis=1)
