"""The reference's per-agent Python loop, restated (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

This is the CPU baseline ``bench.py`` times next to the GPU: the same work the reference does
for every participant of every scenario - one Python call per participant with NumPy *scalar*
float64 arithmetic, a fresh state record per step, then per-pose predicate calls in Python
``for`` loops with early ``break`` - i.e. the execution model of

* ``SingleTrackKinematics.step/_step``  tactics2d/physics/single_track_kinematics.py:178-198,126-176
* ``SingleTrackDynamics.step/_step``    tactics2d/physics/single_track_dynamics.py:231-251,140-229
* ``PointMass.step/_step_newton``       tactics2d/physics/point_mass.py:209-232,83-175
* ``Vehicle.get_pose``                  tactics2d/participant/element/vehicle.py:263-281
* ``DynamicCollision/StaticCollision.update``  tactics2d/traffic/event_detection/collision.py:18-25,37-43
* ``OutBound.update``                   tactics2d/traffic/event_detection/out_bound.py:37-48
* ``check_status``                      tactics2d/envs/parking.py:361-392

shapely/GEOS is not installable in this image, so ``intersects`` / ``contains`` are the closed-set
predicates of ``oracle.geometry`` evaluated per pair (labelled "restatement" in the bench line).
It is held to ``oracle.scenario`` (the vectorised restatement) by tests/test_oracle_scalar_port.py.
"""

from __future__ import annotations

import numpy as np

from . import geometry as G
from .scenario import (CIRCLE, DYNAMICS, F_DYNAMIC, F_OUTBOUND, F_STATIC, INACTIVE, KINEMATICS, NOSHAPE,
                       POINTMASS_EULER, POINTMASS_NEWTON, TABLE_FLOAT_FIELDS)

_G = 9.81


class _State:
    """The 10-field record the reference allocates per participant per step (state.py:47-96)."""

    __slots__ = ("frame", "x", "y", "heading", "vx", "vy", "speed", "accel")

    def __init__(self, frame, x, y, heading, vx=None, vy=None, speed=None, accel=None):
        self.frame, self.x, self.y, self.heading = int(frame), float(x), float(y), float(heading)
        self.vx, self.vy, self.speed, self.accel = vx, vy, speed, accel


def _clip(v, lo, hi):
    return np.clip(v, lo, hi)


def kinematics_step(s, accel, delta, p, interval, delta_t):
    accel = _clip(accel, p["accel_lo"], p["accel_hi"])
    delta = _clip(delta, p["steer_lo"], p["steer_hi"])
    L = p["lf"] + p["lr"]
    beta = np.arctan(p["lr"] / L * np.tan(delta))
    dt = float(delta_t) / 1000
    x, y, phi, v = s.x, s.y, s.heading, s.speed
    steps = [dt] * (interval // delta_t)
    if interval % delta_t > 0:
        steps.append(float(interval % delta_t) / 1000)
    for h in steps:
        dx = v * np.cos(phi + beta)
        dy = v * np.sin(phi + beta)
        dphi = v / L * np.tan(delta) * np.cos(beta)
        x += dx * h
        y += dy * h
        phi += dphi * h
        v += accel * h
        v = _clip(v, p["speed_lo"], p["speed_hi"])
    return _State(s.frame + interval, x, y, np.mod(phi, 2 * np.pi), v * np.cos(phi), v * np.sin(phi), v, accel)


def dynamics_step(s, accel, delta, p, interval, delta_t):
    accel = _clip(accel, p["accel_lo"], p["accel_hi"])
    delta = _clip(delta, p["steer_lo"], p["steer_hi"])
    lf, lr, L = p["lf"], p["lr"], p["lf"] + p["lr"]
    dt = float(delta_t) / 1000
    ff = (_G * lr - accel * p["mass_height"]) / L
    fr = (_G * lf + accel * p["mass_height"]) / L
    a1, a2 = lf * p["cf"] * ff, lr * p["cr"] * fr
    b1, b2 = lf**2 * p["cf"] * ff, lr**2 * p["cr"] * fr
    c1, c2 = p["cf"] * ff, p["cr"] * fr
    x, y, phi, v = s.x, s.y, s.heading, s.speed
    d_phi = v / L * np.tan(delta)
    beta = np.arctan(lr / lf * np.tan(delta))
    for _ in range(interval // delta_t):
        dx = v * np.cos(phi + beta)
        dy = v * np.sin(phi + beta)
        v_safe = v if np.abs(v) > 1e-6 else (1e-6 if v >= 0 else -1e-6)
        if np.abs(v) >= 0.1:
            dd_phi = p["mu"] * p["mass"] / p["I_z"] * (a1 * delta + (a2 - a1) * beta - (b1 + b2) * d_phi / v_safe)
            d_beta = p["mu"] / v_safe * (c1 * delta - (c2 + c1) * beta + (a2 - a1) * d_phi / v_safe) - d_phi
            d_phi += dd_phi * dt
        else:
            d_beta = lr / (1 + np.tan(delta) * lr / L) ** 2 / L / np.cos(delta) ** 2 * delta
            d_phi += v * np.cos(beta) / L * np.tan(delta) * dt
        x += dx * dt
        y += dy * dt
        v += accel * dt
        phi += d_phi * dt
        beta += d_beta * dt
        v = _clip(v, p["speed_lo"], p["speed_hi"])
    h = np.mod(phi, 2 * np.pi)
    return _State(s.frame + interval, x, y, h, v * np.cos(h), v * np.sin(h), v, accel)


def pointmass_newton_step(s, ax, ay, p, interval):
    from .physics import step_pointmass_newton

    o = step_pointmass_newton(s.x, s.y, s.vx, s.vy, ax, ay, (p["speed_lo"], p["speed_hi"]), interval)
    return _State(s.frame + interval, o["x"], o["y"], o["heading"], float(o["vx"]), float(o["vy"]), float(o["speed"]))


def _pose(s, p):
    """get_pose: the rotated box (vehicle.py:263-281) or the disc (pedestrian.py:138-149)."""
    if p["shape"] == CIRCLE:
        return ("c", s.x, s.y, p["radius"])
    return ("b", s.x, s.y, np.cos(s.heading), np.sin(s.heading), p["half_len"], p["half_wid"])


def _intersects(a, b):
    if a[0] == "b" and b[0] == "b":
        return bool(G.obb_obb(a[1], a[2], a[3], a[4], a[5], a[6], b[1], b[2], b[3], b[4], b[5], b[6]))
    if a[0] == "b":
        return bool(G.obb_circle(a[1], a[2], a[3], a[4], a[5], a[6], b[1], b[2], b[3]))
    if b[0] == "b":
        return bool(G.obb_circle(b[1], b[2], b[3], b[4], b[5], b[6], a[1], a[2], a[3]))
    return bool(G.circle_circle(a[1], a[2], a[3], b[1], b[2], b[3]))


def _hits_segment(a, sg):
    if a[0] == "b":
        return bool(G.obb_segment(a[1], a[2], a[3], a[4], a[5], a[6], sg[0], sg[1], sg[2], sg[3]))
    return bool(G.circle_segment(a[1], a[2], a[3], sg[0], sg[1], sg[2], sg[3]))


def _out_bound(a, bounds):
    if a[0] == "b":
        ex, ey = G.extents(a[3], a[4], a[5], a[6])
    else:
        ex = ey = a[3]
    return bool(G.out_of_bound(a[1], a[2], ex, ey, bounds))


def tick_scenarios(state, type_id, action, table, segments=None, bounds=None, interval=100, delta_t=5):
    """One tick of every participant of every scenario, the reference's way.  ``state``: dict of
    [N, M] arrays; returns (new_state dict of float64 arrays, flags, hit_index, hit_segment)."""
    N, M = type_id.shape
    out = {k: np.array(state[k], dtype=np.float64) for k in ("x", "y", "heading", "speed", "vx", "vy")}
    flags = np.zeros((N, M), np.uint8)
    hit_index = np.full((N, M), -1, np.int16)
    hit_segment = np.full((N, M), -1, np.int16)
    rows = [{k: float(np.float64(np.float32(table[k][t]))) for k in TABLE_FLOAT_FIELDS} | {"model": int(table["model"][t]), "shape": int(table["shape"][t])}
            for t in range(len(table["model"]))]
    seg = None if segments is None else [tuple(float(v) for v in s) for s in np.asarray(segments, dtype=np.float64)]
    for n in range(N):
        poses = [None] * M
        for m in range(M):   # ScenarioManager.update: physics_model.step + add_state
            t = int(type_id[n, m])
            if t == INACTIVE:
                continue
            p = rows[t]
            s = _State(0, state["x"][n, m], state["y"][n, m], state["heading"][n, m], float(state["vx"][n, m]),
                       float(state["vy"][n, m]), float(state["speed"][n, m]))
            a0, a1 = float(action[n, m, 0]), float(action[n, m, 1])
            if p["model"] == KINEMATICS:
                s = kinematics_step(s, a0, a1, p, interval, delta_t)
            elif p["model"] == DYNAMICS:
                s = dynamics_step(s, a0, a1, p, interval, delta_t)
            elif p["model"] == POINTMASS_NEWTON:
                s = pointmass_newton_step(s, a0, a1, p, interval)
            elif p["model"] == POINTMASS_EULER:
                from .physics import step_pointmass_euler

                o = step_pointmass_euler(s.x, s.y, s.heading, s.vx, s.vy, a0, a1, (p["speed_lo"], p["speed_hi"]), interval, delta_t)
                s = _State(interval, o["x"], o["y"], o["heading"], float(o["vx"]), float(o["vy"]), float(o["speed"]))
            out["x"][n, m], out["y"][n, m], out["heading"][n, m] = s.x, s.y, s.heading
            out["speed"][n, m], out["vx"][n, m], out["vy"][n, m] = s.speed, s.vx, s.vy
            if p["shape"] != NOSHAPE:
                poses[m] = _pose(s, p)
        for m in range(M):   # check_status for every participant as the ego
            a = poses[m]
            if a is None:
                continue
            if bounds is not None and _out_bound(a, bounds):
                flags[n, m] |= F_OUTBOUND
            if seg is not None:
                for k, sg in enumerate(seg):        # StaticCollision.update: list order, break
                    if _hits_segment(a, sg):
                        flags[n, m] |= F_STATIC
                        hit_segment[n, m] = k
                        break
            for j in range(M):                      # DynamicCollision.update: list order, break
                if j == m or poses[j] is None:
                    continue
                if _intersects(a, poses[j]):
                    flags[n, m] |= F_DYNAMIC
                    hit_index[n, m] = j
                    break
    return out, flags, hit_index, hit_segment
