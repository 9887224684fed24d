"""Batched event detectors with the reference's class names.

The reference's detectors take ONE shapely pose and return ONE bool
(``tactics2d/traffic/event_detection/{collision,out_bound,time_exceed}.py``).  Here ``update(world)`` takes a
:class:`tactics2d_b200.BatchedWorld` and returns device tensors for all N x M participants; the arithmetic
runs in the fused kernel (``t2d_step`` computes every detector in the same pass as the physics;
``t2d_check_events`` evaluates them alone).  Semantics kept from the reference:

* ``StaticCollision``  first static object in list order that the pose ``intersects`` (collision.py:37-43);
* ``DynamicCollision`` first other participant in list order whose pose ``intersects`` (collision.py:18-25;
  the reference method as written dereferences ``.geometry`` on shapely objects and cannot run - its intent,
  ego against every other pose with ``break``, is what is implemented, for every participant as the ego);
* ``OutBound``         ``not box.contains(pose)`` with box = (xmin, xmax, ymin, ymax) (out_bound.py:28-48);
* ``TimeExceed``       ``cnt_step += 1; cnt_step > max_step`` (time_exceed.py:26-33).
"""

from __future__ import annotations

from typing import Optional, Sequence

from .event_base import EventBase

F_DYNAMIC, F_STATIC, F_OUTBOUND = 1, 2, 4


def _events(world, fresh: bool):
    """Flags of the last ``step`` (they were computed in that pass) or a fresh ``check_events`` launch."""
    return world.check_events() if fresh else world._out


class DynamicCollision(EventBase):
    def __init__(self):
        super().__init__()

    def update(self, world, fresh: bool = True):
        """-> (collided bool [N, M], first-hit participant index int16 [N, M], -1 = none)."""
        r = _events(world, fresh)
        return (r.flags & F_DYNAMIC) != 0, r.hit_index

    def reset(self):
        return


class StaticCollision(EventBase):
    def __init__(self, static_objects: Optional[Sequence] = None):
        self.static_objects = static_objects   # [S, 4] segments (x1, y1, x2, y2) in list order

    def update(self, world, fresh: bool = True):
        """-> (collided bool [N, M], first-hit segment index int16 [N, M], -1 = none)."""
        r = _events(world, fresh)
        return (r.flags & F_STATIC) != 0, r.hit_segment

    def reset(self, static_objects=None, world=None):
        """Replace the static objects; with ``world`` given the map tile is re-staged on the device."""
        self.static_objects = static_objects
        if world is not None:
            world.set_map(static_objects, world.bounds)


class OutBound(EventBase):
    def __init__(self, boundary: tuple = None):
        self.map_boundary = boundary   # (xmin, xmax, ymin, ymax)

    def update(self, world, fresh: bool = True):
        r = _events(world, fresh)
        return (r.flags & F_OUTBOUND) != 0

    def reset(self, boundary: tuple = None, world=None):
        self.map_boundary = boundary
        if world is not None:
            world.set_map(world.segments, boundary)


class TimeExceed(EventBase):
    def __init__(self, max_step: int):
        self.max_step = max_step
        self.cnt_step = 0

    def update(self, world=None):
        """Scalar form (reference): ``++cnt > max_step``.  With a world: the per-scenario device counters,
        which ``t2d_step`` increments, compared against ``max_step`` -> bool [N]."""
        if world is None:
            self.cnt_step += 1
            return self.cnt_step > self.max_step
        return world.step_count > self.max_step

    def reset(self, world=None):
        self.cnt_step = 0
        if world is not None:
            world.step_count.zero_()


class Arrival(EventBase):
    """``Arrival`` (reference arrival.py:12-50): IoU of the ego's pose with the target area, completed when
    ``iou >= threshold`` (default 0.95).  Batched: ``reset(target, world)`` installs one target rectangle per
    scenario, ``update(world)`` returns ``(is_completed bool [N], iou fp32 [N])`` of the last ``step``."""

    def __init__(self, target_area=None, threshold: float = 0.95):
        self.target_area = target_area   # [N, 5] = (cx, cy, heading, half_len, half_wid)
        self.threshold = threshold

    def update(self, world):
        r = world._out
        if r.iou is None:
            raise RuntimeError("no target area: call Arrival.reset(target_area, world) / world.set_goal first")
        return r.iou >= self.threshold, r.iou

    def reset(self, target_area=None, world=None, no_action_max_step: int = 100):
        self.target_area = target_area
        if world is not None:
            world.set_goal(target_area, self.threshold, no_action_max_step)


class NoAction(EventBase):
    """``NoAction`` (reference no_action.py:12-58): counts consecutive ticks in which the ego's pose overlaps its
    previous pose with IoU > 0.999; fires when the count exceeds ``max_step``.  The counter lives on the device
    (``world.set_goal(..., no_action_max_step=max_step)``); ``update(world)`` reads it."""

    def __init__(self, max_step=100):
        self.max_step = max_step
        self.cnt_no_action = 0

    def update(self, world):
        if world._goal is None:
            raise RuntimeError("NoAction needs world.set_goal(target, no_action_max_step=...)")
        return world._goal["count"] > self.max_step

    def reset(self, world=None):
        self.cnt_no_action = 0
        if world is not None and world._goal is not None:
            world._goal["count"].zero_()
            world._goal["last_pose"].zero_()
