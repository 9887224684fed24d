"""Detectors on this hot path (reference ``tactics2d/traffic/event_detection/__init__.py:7-26``).  ``NoAction``
and ``Arrival`` (IoU based) are "next" rows of SURVEY.md section 8(f); ``OffRoute`` / ``OffLane`` are
unused / a stub in the reference."""

from .detectors import DynamicCollision, OutBound, StaticCollision, TimeExceed
from .event_base import EventBase

__all__ = ["EventBase", "DynamicCollision", "StaticCollision", "OutBound", "TimeExceed"]
