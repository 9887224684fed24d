"""``Obstacle`` (reference ``tactics2d/participant/element/obstacle.py:14-19``): an ``Other`` whose
``get_state(frame)`` returns the state at the closest recorded frame."""

import numpy as np

from ..trajectory import State
from .other import Other


class Obstacle(Other):
    def get_state(self, frame: int = None) -> State:
        if frame is None:
            return self.current_state
        frames = np.array(self.trajectory.frames)
        return self.trajectory.get_state(int(frames[np.abs(frames - frame).argmin()]))
