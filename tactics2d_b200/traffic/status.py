"""Status enums with the reference's names and codes (``tactics2d/traffic/status.py:10-61``); the kernels
write these integer codes into ``uint8`` tensors."""

from enum import IntEnum


class ScenarioStatus(IntEnum):
    """High-level status of a scenario (status.py:23-28)."""

    NORMAL = 1
    COMPLETED = 2
    TIME_EXCEEDED = 3
    OUT_BOUND = 4
    NO_ACTION = 5
    FAILED = 6


class TrafficStatus(IntEnum):
    """Low-level status of the (ego) participant (status.py:52-61)."""

    NORMAL = 1
    UNKNOWN = 2
    COLLISION_STATIC = 3
    COLLISION_DYNAMIC = 4
    OFF_ROUTE = 5
    OFF_LANE = 6
    VIOLATION_RETROGRADE = 7
    VIOLATION_NON_DRIVABLE = 8
    VIOLATION_TRAFFIC_LIGHT = 9
    VIOLATION_TRAFFIC_SIGN = 10
