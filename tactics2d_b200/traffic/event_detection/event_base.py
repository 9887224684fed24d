"""``EventBase`` (reference ``tactics2d/traffic/event_detection/event_base.py:10-19``): ``update`` / ``reset``."""

from abc import ABC, abstractmethod


class EventBase(ABC):
    @abstractmethod
    def update(self, *args, **kwargs):
        """Update the detector with the current information."""

    @abstractmethod
    def reset(self):
        """Reset the detector."""
