"""Gym-style environments on this hot path.  The reference exports the single-ego ``ParkingEnv`` /
``RacingEnv`` (``tactics2d/envs/__init__.py:7-10``); their tick is what :class:`BatchedTrafficEnv` batches."""

from .batched_env import BatchedTrafficEnv, InvalidAction

__all__ = ["BatchedTrafficEnv", "InvalidAction"]
