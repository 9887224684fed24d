"""``State``: one participant at one time stamp.

Same record, property semantics and error behaviour as the reference's
``tactics2d/participant/trajectory/state.py:12-223`` (and as its tests pin them,
tests/test_participant.py:158-191,498-541):

* ten fields - ``frame`` (int, ms), ``x, y, heading``, optional ``vx, vy, speed, ax, ay, accel`` -
  each coerced to its declared type on assignment, ``ValueError`` when that fails (:108-126);
* ``speed``: the stored scalar if there is one, else ``|(vx, vy)|`` which is then *stored* (so a
  later change of ``vx`` alone does not alter it - pinned by the reference test at :532-536);
* ``velocity``: ``(vx, vy)`` if both are set, else ``speed * (cos, sin)(heading)`` (:148-165);
* ``accel``: ``|(ax, ay)|`` when both are set, else the norm of ``acceleration`` (:171-185) - i.e. the
  absolute value of the stored scalar, not the scalar itself;
* ``acceleration``: ``(ax, ay)`` or ``accel_scalar * (cos, sin)(heading)`` (:187-204).

In the batched engine a State is a *view* of one column of the SoA tensors (``BatchedWorld``); this
class is the host-side record the reference's signatures (``PhysicsModelBase.step(state, ...)``) speak.
"""

from __future__ import annotations

import math
from typing import Any, Optional, Tuple

_TYPES = {"frame": int, "x": float, "y": float, "heading": float, "vx": float, "vy": float, "_speed": float,
          "ax": float, "ay": float, "_accel": float}


class State:
    __annotations__ = dict(_TYPES)
    __slots__ = tuple(_TYPES)

    def __init__(self, frame: int, x: float = 0, y: float = 0, heading: float = 0, vx: float = None, vy: float = None,
                 speed: float = None, ax: float = None, ay: float = None, accel: float = None):
        for name, value in (("frame", frame), ("x", x), ("y", y), ("heading", heading), ("vx", vx), ("vy", vy),
                            ("_speed", speed), ("ax", ax), ("ay", ay), ("_accel", accel)):
            setattr(self, name, value)

    def __setattr__(self, name: str, value: Any) -> None:
        want = _TYPES.get(name)
        if want is not None and value is not None and not isinstance(value, want):
            try:
                value = want(value)
            except Exception:
                raise ValueError(f"Failed to convert {value} to the expected type of {name}: ({want}).")
        object.__setattr__(self, name, value)

    def __str__(self):
        return (f"{self.__class__.__name__}(frame={self.frame}, x={self.x}, y={self.y}, heading={self.heading}, "
                f"vx={self.vx}, vy={self.vy}, speed={self.speed}, ax={self.ax}, ay={self.ay}, accel={self.accel})")

    __repr__ = __str__

    @property
    def location(self) -> Tuple[float, float]:
        return (self.x, self.y)

    @property
    def speed(self) -> Optional[float]:
        if self._speed is None and self.vx is not None and self.vy is not None:
            self._speed = math.hypot(self.vx, self.vy)
        return self._speed

    @property
    def velocity(self) -> Optional[Tuple[float, float]]:
        if self.vx is not None and self.vy is not None:
            return (self.vx, self.vy)
        if self.speed is not None and self.heading is not None:
            return (self.speed * math.cos(self.heading), self.speed * math.sin(self.heading))
        return None

    @property
    def acceleration(self) -> Optional[Tuple[float, float]]:
        if self.ax is not None and self.ay is not None:
            return (self.ax, self.ay)
        if self._accel is not None and self.heading is not None:
            return (self._accel * math.cos(self.heading), self._accel * math.sin(self.heading))
        return None

    @property
    def accel(self) -> Optional[float]:
        a = self.acceleration
        return None if a is None else math.hypot(a[0], a[1])

    def set_heading(self, heading: float):
        self.heading = heading

    def set_velocity(self, vx: float, vy: float):
        self.vx = vx
        self.vy = vy

    def set_speed(self, speed: float):
        self._speed = speed

    def set_accel(self, ax: float, ay: float):
        self.ax = ax
        self.ay = ay
        self._accel = math.hypot(self.ax, self.ay)
