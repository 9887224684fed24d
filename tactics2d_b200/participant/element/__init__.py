"""Participant classes with the reference's names (``tactics2d/participant/element/__init__.py:7-29``)."""

from .cyclist import Cyclist
from .obstacle import Obstacle
from .other import Other
from .participant_base import ParticipantBase
from .participant_template import list_cyclist_templates, list_pedestrian_templates, list_vehicle_templates
from .pedestrian import Pedestrian
from .vehicle import Vehicle

__all__ = ["ParticipantBase", "Pedestrian", "Cyclist", "Vehicle", "Other", "Obstacle", "list_vehicle_templates",
           "list_cyclist_templates", "list_pedestrian_templates"]
