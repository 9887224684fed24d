"""Dataset parsers on the path's input side: logged trajectories -> initial-state pools for ``BatchedWorld.reset``."""
from .parse_levelx import LevelXParser, initial_state_pool  # noqa: F401
