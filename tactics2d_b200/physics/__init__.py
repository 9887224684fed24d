"""Physics models with the reference's names and signatures (``tactics2d/physics/__init__.py:7-19``)."""

from .physics_model_base import PhysicsModelBase
from .point_mass import PointMass
from .single_track_drift import SingleTrackDrift
from .single_track_dynamics import SingleTrackDynamics
from .single_track_kinematics import SingleTrackKinematics

__all__ = ["PhysicsModelBase", "PointMass", "SingleTrackKinematics", "SingleTrackDynamics", "SingleTrackDrift"]
