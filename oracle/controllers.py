"""Float64 restatement of the reference's NPC controllers (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this package; the product path
(``tactics2d_b200``) never does.

Restated, one vectorised NumPy function per reference method:

* ``IDMController.step`` / ``_idm_acceleration`` - ``tactics2d/controller/idm_controller.py:59-141``;
* ``AccelerationController._cruise_control`` / ``_adaptive_cruise_control`` / ``step`` -
  ``tactics2d/controller/acceleration_controller.py:82-145``;
* ``PurePursuitController._lateral_control`` / ``step`` - ``tactics2d/controller/pure_pursuit_controller.py:51-98``;
  its ``waypoints.interpolate(d)`` is shapely's ``LineString.interpolate`` (third-party, ``shapely>=2.0.7,<2.1.0``,
  not installable here): the point at arc length ``d`` from the FIRST vertex, clamped to the last vertex
  (``d`` is never negative on this path: it is ``max(speed * interval, min_pre_aiming_distance) > 0``).

PINNED: ``tests/golden/controllers.npz`` holds outputs of the unmodified reference classes (``oracle/make_golden.py``
imports the three modules with a stand-in for ``shapely.geometry`` that provides only ``LineString.interpolate`` /
``Point`` - so IDM and the two longitudinal laws are the reference's own arithmetic end to end, and pure pursuit is the
reference's arithmetic around the restated interpolation).

``accel_last`` quirk kept: ``State.accel`` is ``|a|`` (``participant/trajectory/state.py:171-185`` returns the norm of
``(a cos h, a sin h)``), so the rate limit of the longitudinal laws is centred on the magnitude of the last applied
acceleration.  The reference raises ``TypeError`` when the state carries no acceleration yet; the batched path defines
that first tick as ``accel_last = 0``.
"""

from __future__ import annotations

import numpy as np

EXTERNAL, IDM, CRUISE, PURE_PURSUIT = 0, 1, 2, 3

# defaults of the reference classes, in the order of ``t2d_controller_params``
IDM_DEFAULTS = dict(desired_speed=10.0, time_headway=1.5, min_spacing=2.0, max_acceleration=1.0,
                    comfortable_deceleration=3.0, delta=4.0)                       # idm_controller.py:33-41
ACC_DEFAULTS = dict(target_speed=5.0, kp=3.5, accel_change_rate=3.0, delta_t=0.05, max_accel=1.5, min_accel=-4.0,
                    interval=2.0)                                                  # acceleration_controller.py:33-45
PP_DEFAULTS = dict(min_pre_aiming_distance=10.0, pp_interval=1.0, wheel_base=2.637)   # pure_pursuit_controller.py:26-28,76
SAFETY_DISTANCE, MIN_TARGET_DISTANCE, MAX_TARGET_DISTANCE = 5.0, 7.0, 80.0          # acceleration_controller.py:42-45


def _clip(x, lo, hi):
    return np.minimum(np.maximum(x, lo), hi)      # np.clip: NaN propagates


def idm(v, x, y, has_lead, v_lead, x_lead, y_lead, p):
    """idm_controller.py:59-141 -> acceleration (steering is always 0.0, :92)."""
    v, x, y, v_lead, x_lead, y_lead = (np.asarray(a, np.float64) for a in (v, x, y, v_lead, x_lead, y_lead))
    vd, T, s0, am, b, de = (np.float64(p[k]) for k in ("desired_speed", "time_headway", "min_spacing", "max_acceleration",
                                                       "comfortable_deceleration", "delta"))
    with np.errstate(all="ignore"):
        if vd > 0:
            ratio = (v / vd) ** de                                                  # :77-79, :127
            free = am * (1.0 - ratio)
            ratio_follow = ratio
        else:
            free = np.where(v > 0, -b, 0.0)                                         # :82
            ratio_follow = np.where(v > 0, 1.0, 0.0)                                # :130
        dist = np.hypot(x_lead - x, y_lead - y)                                     # :107-109
        dv = v_lead - v                                                             # :112
        s_star = s0 + v * T + (v * dv) / (2.0 * np.sqrt(am * b))                    # :116-120
        s_star = np.maximum(s_star, s0)                                             # :121
        follow = np.where(dist > 0, am * (1.0 - ratio_follow - (s_star / np.where(dist > 0, dist, 1.0)) ** 2), -b)   # :125-138
        a = np.where(has_lead, follow, free)
        return _clip(a, -b, am)                                                     # :89


def cruise(v, accel_last, p):
    """acceleration_controller.py:82-102."""
    v, accel_last = np.asarray(v, np.float64), np.asarray(accel_last, np.float64)
    a = (np.float64(p["target_speed"]) - v) / np.float64(p["kp"])
    w = np.float64(p["accel_change_rate"]) * np.float64(p["delta_t"])
    a = _clip(a, accel_last - w, accel_last + w)
    return _clip(a, np.float64(p["min_accel"]), np.float64(p["max_accel"]))


def adaptive_cruise(v, x, y, accel_last, v_lead, x_lead, y_lead, accel_lead, p):
    """acceleration_controller.py:104-130."""
    v, x, y, accel_last, v_lead, x_lead, y_lead, accel_lead = (
        np.asarray(a, np.float64) for a in (v, x, y, accel_last, v_lead, x_lead, y_lead, accel_lead))
    kp = np.float64(p["kp"])
    d_front = np.hypot(x - x_lead, y - y_lead)
    d_target = _clip(v * np.float64(p["interval"]) + SAFETY_DISTANCE, MIN_TARGET_DISTANCE, MAX_TARGET_DISTANCE)
    rel_speed = v_lead - v
    rel_target_speed = (d_target - d_front) / kp
    rel_accel = (rel_target_speed - rel_speed) / kp
    a = accel_lead - rel_accel
    w = np.float64(p["accel_change_rate"]) * np.float64(p["delta_t"])
    a = _clip(a, accel_last - w, accel_last + w)
    return _clip(a, np.float64(p["min_accel"]), np.float64(p["max_accel"]))


def interpolate(path, d):
    """shapely ``LineString.interpolate(d)``, d >= 0: walk the segments from the first vertex."""
    path = np.asarray(path, np.float64)
    seg = np.hypot(*(path[1:] - path[:-1]).T)
    acc = 0.0
    for i, L in enumerate(seg):
        if d <= acc + L and L > 0:
            t = (d - acc) / L
            return path[i] + t * (path[i + 1] - path[i])
        acc += L
    return path[-1].copy()


def pure_pursuit_steering(x, y, heading, v, path, p):
    """pure_pursuit_controller.py:51-74,90-92 for one participant (scalar float64)."""
    d = max(float(v) * float(p["pp_interval"]), float(p["min_pre_aiming_distance"]))   # :90-91
    px, py = interpolate(path, d)                                                      # :92
    with np.errstate(all="ignore"):
        ang = np.arctan2(py - y, px - x)                                               # :62-64
        dist = np.hypot(py - y, px - x)                                                # :65-67
        return float(np.arctan(np.float64(2.0) * np.float64(p["wheel_base"]) * np.sin(ang - heading) / np.float64(dist)))   # :68-70


def applied_accel_magnitude(action, type_id, table, steer_first=False):
    """|a| the physics applied for this tick's raw action (what the next tick's ``State.accel`` returns):
    bicycles clip to the type's accel range (single_track_kinematics.py:192); point masses take (ax, ay) unclipped."""
    action = np.asarray(action, np.float64)
    tid = np.where(type_id == 255, 0, type_id).astype(np.int64)
    model = np.asarray(table["model"])[tid]
    lo = np.asarray(table["accel_lo"], np.float64)[tid]
    hi = np.asarray(table["accel_hi"], np.float64)[tid]
    a = action[..., 1] if steer_first else action[..., 0]
    bic = np.abs(_clip(a, lo, hi))
    pm = np.hypot(action[..., 0], action[..., 1])
    out = np.where(model == 4, 0.0, np.where((model == 2) | (model == 3), pm, bic))   # 4 = static: no acceleration
    return np.where(type_id == 255, 0.0, out)


def control_tick(state, type_id, table, action, ctrl_id, ctrl_table, lead_index, path_id, paths, last_accel,
                 steer_first=False):
    """One batched controller pass: returns ``(action', last_accel')``.

    ``action`` [N, M, 2] float32 holds the externally supplied actions; rows whose ``ctrl_id`` selects a controller
    are overwritten with ``(accel, steer)`` (``(steer, accel)`` when ``steer_first``), rounded to float32.
    ``last_accel'`` is ``applied_accel_magnitude`` of the final action buffer, for every participant."""
    x, y, h, v = (np.asarray(state[k], np.float64) for k in ("x", "y", "heading", "speed"))
    N, M = x.shape
    out = np.array(action, np.float32, copy=True)
    la = np.asarray(last_accel, np.float64)
    for n in range(N):
        for m in range(M):
            cid = int(ctrl_id[n, m])
            if cid == 255 or int(type_id[n, m]) == 255:
                continue
            p = ctrl_table[cid]
            kind = int(p["kind"])
            if kind == EXTERNAL:
                continue
            li = -1 if lead_index is None else int(lead_index[n, m])
            has = 0 <= li < M and int(type_id[n, li]) != 255 and li != m
            xl, yl, vl, al = (x[n, li], y[n, li], v[n, li], la[n, li]) if has else (0.0, 0.0, 0.0, 0.0)
            steer = 0.0
            if kind == IDM:
                acc = float(idm(v[n, m], x[n, m], y[n, m], has, vl, xl, yl, p))
            else:
                acc = float(adaptive_cruise(v[n, m], x[n, m], y[n, m], la[n, m], vl, xl, yl, al, p) if has
                            else cruise(v[n, m], la[n, m], p))
                if kind == PURE_PURSUIT:
                    pid = -1 if path_id is None else int(path_id[n, m])
                    if 0 <= pid < len(paths):
                        steer = pure_pursuit_steering(x[n, m], y[n, m], h[n, m], v[n, m], paths[pid], p)
            out[n, m] = (steer, acc) if steer_first else (acc, steer)
    return out, applied_accel_magnitude(out, type_id, table, steer_first).astype(np.float32)
