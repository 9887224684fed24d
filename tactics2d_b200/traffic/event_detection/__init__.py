"""Detectors on this hot path (reference ``tactics2d/traffic/event_detection/__init__.py:7-26``), including the
IoU based ``Arrival`` / ``NoAction``.  ``OffRoute`` / ``OffLane`` are unused / a
stub in the reference."""

from .detectors import Arrival, DynamicCollision, NoAction, OutBound, StaticCollision, TimeExceed
from .event_base import EventBase

__all__ = ["EventBase", "DynamicCollision", "StaticCollision", "OutBound", "TimeExceed", "Arrival", "NoAction"]
