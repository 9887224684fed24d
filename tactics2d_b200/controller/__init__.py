"""NPC controllers evaluated on the device (``tactics2d.controller`` surface).

``IDMController``, ``AccelerationController`` and ``PurePursuitController`` keep the reference's constructor arguments,
attributes, ``update_driving_style`` / ``configure`` and ``step(ego_state, ...) -> (steering, acceleration)``
(tactics2d/controller/*.py).  They are parameter holders: ``BatchedWorld.set_controllers`` turns a list of them into the
controller table of ``t2d_control``, which evaluates every controlled participant of every scenario in one launch;
``step`` on a single ``State`` goes through the same kernel with a batch of one.
"""

from .acceleration_controller import AccelerationController
from .controller_base import ControllerBase
from .idm_controller import IDMController
from .pure_pursuit_controller import PurePursuitController

__all__ = ["ControllerBase", "AccelerationController", "IDMController", "PurePursuitController"]
