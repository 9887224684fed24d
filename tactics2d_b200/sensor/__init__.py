"""Sensors on this path: the single-line lidar (the vector observation of the reference's ParkingEnv,
envs/parking.py:303-304,422-429)."""

from .lidar import SingleLineLidar

__all__ = ["SingleLineLidar"]
