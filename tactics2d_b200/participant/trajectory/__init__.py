from .state import State
from .trajectory import Trajectory

__all__ = ["State", "Trajectory"]
