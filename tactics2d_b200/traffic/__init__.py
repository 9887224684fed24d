"""``tactics2d.traffic`` surface on this hot path (reference ``tactics2d/traffic/__init__.py:7-10``)."""

from .scenario_manager import BatchedScenarioManager, ScenarioManager
from .status import ScenarioStatus, TrafficStatus

__all__ = ["ScenarioManager", "BatchedScenarioManager", "ScenarioStatus", "TrafficStatus"]
