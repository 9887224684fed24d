"""Float64 restatement of one batched scenario tick (TEST INFRASTRUCTURE ONLY).

The tick the reference performs for ONE ego in ``_ParkingScenarioManager.update`` /
``check_status`` (envs/parking.py:352-392), stated for N scenarios x M participants:

1. ``update``  (parking.py:352-356): ``cnt_step += 1``; every participant's physics
   model steps its current state (``oracle.physics``), the new state is appended.
2. ``check_status`` (parking.py:361-392), on the NEW poses, first hit returns:
   time-exceed (time_exceed.py:26-33) -> no-action (no_action.py:32-53, ``goal_events``)
   -> out-of-bound (out_bound.py:37-48) -> static collision (collision.py:37-43, first
   object in list order, ``break``) -> [dynamic collision (collision.py:18-25: the ego
   against every other participant in list order, ``break``) - unused by the reference
   envs, appended here after static] -> arrival (arrival.py:32-47, ``goal_events``).
3. ``terminated/truncated`` (parking.py:243-248): done = status != NORMAL.

Extensions the reference leaves undefined and this build defines:
every participant is checked (not only the ego); pedestrians collide as discs
(``Pedestrian.get_pose`` is ``((x, y), width/2)``, pedestrian.py:138-149).

Status codes: traffic/status.py:23-28 (ScenarioStatus), :52-61 (TrafficStatus).
"""

from __future__ import annotations

import numpy as np

from . import geometry as G
from . import physics as P

# model ids / shape ids of the type table (mirrors include/t2d_b200.h)
KINEMATICS, DYNAMICS, POINTMASS_NEWTON, POINTMASS_EULER, STATIC = 0, 1, 2, 3, 4
OBB, CIRCLE, NOSHAPE = 0, 1, 2
INACTIVE = 255

# flag bits of the per-participant event byte
F_DYNAMIC, F_STATIC, F_OUTBOUND = 1, 2, 4

# ScenarioStatus, traffic/status.py:23-28
NORMAL, COMPLETED, TIME_EXCEEDED, OUT_BOUND, NO_ACTION, FAILED = 1, 2, 3, 4, 5, 6

TABLE_FLOAT_FIELDS = ("half_len", "half_wid", "radius", "lf", "lr", "steer_lo", "steer_hi",
                      "speed_lo", "speed_hi", "accel_lo", "accel_hi", "mass", "mass_height",
                      "mu", "I_z", "cf", "cr")


DRIFT_FIELDS = ("wheel_radius", "T_sb", "T_se", "I_yw")
DRIFT = 5


def _gather(table, type_id):
    tid = np.where(type_id == INACTIVE, 0, type_id).astype(np.int64)
    out = {k: np.asarray(table[k], dtype=np.float64)[tid] for k in TABLE_FLOAT_FIELDS}
    for k in DRIFT_FIELDS:      # SingleTrackDrift rows only; older tables do not carry them
        if k in table:
            out[k] = np.asarray(table[k], dtype=np.float64)[tid]
    out["model"] = np.asarray(table["model"])[tid]
    out["shape"] = np.asarray(table["shape"])[tid]
    return out


def physics_tick(state, type_id, action, table, interval=100, delta_t=5, steer_first=False):
    """Step every participant with its own model.  ``state``: dict of [N, M] arrays
    x, y, heading, speed, vx, vy.  ``action``: [N, M, 2] = (accel, steer) for the
    bicycles (``(steer, accel)`` when ``steer_first`` - the env action order,
    parking.py:239), (ax, ay) for point masses.  Returns float64 arrays."""
    p = _gather(table, type_id)
    a0 = np.asarray(action[..., 0], dtype=np.float64)
    a1 = np.asarray(action[..., 1], dtype=np.float64)
    acc, ste = (a1, a0) if steer_first else (a0, a1)
    keys = ("x", "y", "heading", "speed", "vx", "vy") + tuple(k for k in ("omega_wf", "omega_wr") if k in state)
    s = {k: np.asarray(state[k], dtype=np.float64) for k in keys}
    out = {k: s[k].copy() for k in s}
    active = type_id != INACTIVE
    rng = lambda a, b: (p[a], p[b])

    def put(mask, res):
        for k in out:
            if k in res:          # only the drift model returns wheel speeds
                out[k] = np.where(mask, res[k], out[k])

    m = active & (p["model"] == KINEMATICS)
    if m.any():
        put(m, P.step_kinematics(s["x"], s["y"], s["heading"], s["speed"], acc, ste, p["lf"], p["lr"],
                                 rng("steer_lo", "steer_hi"), rng("speed_lo", "speed_hi"),
                                 rng("accel_lo", "accel_hi"), interval, delta_t))
    m = active & (p["model"] == DYNAMICS)
    if m.any():
        put(m, P.step_dynamics(s["x"], s["y"], s["heading"], s["speed"], acc, ste, p["lf"], p["lr"],
                               p["mass"], p["mass_height"], p["mu"], p["I_z"], p["cf"], p["cr"],
                               rng("steer_lo", "steer_hi"), rng("speed_lo", "speed_hi"),
                               rng("accel_lo", "accel_hi"), interval, delta_t))
    m = active & (p["model"] == DRIFT)
    if m.any():     # wheel speeds travel in state["omega_wf"], state["omega_wr"] (single_track_drift.py:467-499)
        put(m, P.step_drift(s["x"], s["y"], s["heading"], s["speed"], s["omega_wf"], s["omega_wr"], acc, ste, p["lf"], p["lr"],
                            p["mass"], p["wheel_radius"], p["T_sb"], p["T_se"], p["I_z"], p["I_yw"],
                            rng("steer_lo", "steer_hi"), rng("speed_lo", "speed_hi"), rng("accel_lo", "accel_hi"),
                            interval, delta_t))
    m = active & (p["model"] == POINTMASS_NEWTON)
    if m.any():
        put(m, P.step_pointmass_newton(s["x"], s["y"], s["vx"], s["vy"], a0, a1,
                                       rng("speed_lo", "speed_hi"), interval))
    m = active & (p["model"] == POINTMASS_EULER)
    if m.any():
        put(m, P.step_pointmass_euler(s["x"], s["y"], s["heading"], s["vx"], s["vy"], a0, a1,
                                      rng("speed_lo", "speed_hi"), interval, delta_t))
    return out


def events(x, y, heading, type_id, table, segments=None, bounds=None, chunk=64, poly_start=None):
    """Collision / out-of-bound events on the given poses.

    Returns ``flags`` uint8 [N, M] (bit0 dynamic, bit1 static, bit2 out-of-bound),
    ``hit_index`` int16 [N, M] (lowest colliding participant index or -1:
    collision.py:18-25 iterates in list order and breaks on the first hit) and
    ``hit_segment`` int16 [N, M] (lowest colliding map segment or -1, collision.py:37-43).

    ``poly_start`` [P + 1]: the segments [poly_start[p], poly_start[p + 1]) are the edges of a closed Area polygon.  The
    reference tests ``agent_pose.intersects(static_object.geometry)`` object by object (collision.py:37-43): a polygon is
    hit when one of its edges meets the pose OR the pose lies inside it (a pose inside and clear of every edge contains its
    own centre, so "centre in polygon" decides); ``hit_segment`` then names the first OBJECT hit by its first segment.
    """
    x, y, heading = (np.asarray(a, dtype=np.float64) for a in (x, y, heading))
    N, M = x.shape
    p = _gather(table, type_id)
    solid = (type_id != INACTIVE) & (p["shape"] != NOSHAPE)
    circ = p["shape"] == CIRCLE
    c, s = np.cos(heading), np.sin(heading)
    hl, hw, r = p["half_len"], p["half_wid"], p["radius"]
    flags = np.zeros((N, M), np.uint8)
    hit_index = np.full((N, M), -1, np.int16)
    hit_segment = np.full((N, M), -1, np.int16)

    for n0 in range(0, N, chunk):
        sl = slice(n0, min(N, n0 + chunk))
        A = lambda a: a[sl, :, None]
        B = lambda a: a[sl, None, :]
        oo = G.obb_obb(A(x), A(y), A(c), A(s), A(hl), A(hw), B(x), B(y), B(c), B(s), B(hl), B(hw))
        oc = G.obb_circle(A(x), A(y), A(c), A(s), A(hl), A(hw), B(x), B(y), B(r))
        co = G.obb_circle(B(x), B(y), B(c), B(s), B(hl), B(hw), A(x), A(y), A(r))
        cc = G.circle_circle(A(x), A(y), A(r), B(x), B(y), B(r))
        ca, cb = A(circ), B(circ)
        hit = np.where(ca, np.where(cb, cc, co), np.where(cb, oc, oo))
        hit &= A(solid) & B(solid) & ~np.eye(M, dtype=bool)[None]
        any_hit = hit.any(-1)
        hit_index[sl] = np.where(any_hit, hit.argmax(-1), -1).astype(np.int16)
        flags[sl] |= np.where(any_hit, F_DYNAMIC, 0).astype(np.uint8)

        if segments is not None and len(segments) > 0:
            seg = np.asarray(segments, dtype=np.float64)
            S = lambda k: seg[None, None, :, k]
            os_ = G.obb_segment(A(x), A(y), A(c), A(s), A(hl), A(hw), S(0), S(1), S(2), S(3))
            cs = G.circle_segment(A(x), A(y), A(r), S(0), S(1), S(2), S(3))
            sh = np.where(ca, cs, os_) & A(solid)
            any_s = sh.any(-1)
            first = np.where(any_s, sh.argmax(-1), np.iinfo(np.int32).max).astype(np.int64)
            if poly_start is not None and len(poly_start) > 1:
                ps = np.asarray(poly_start, dtype=np.int64)
                obj_first = np.arange(len(seg), dtype=np.int64)
                for p0, p1 in zip(ps[:-1], ps[1:]):
                    obj_first[p0:p1] = p0
                first = np.where(any_s, obj_first[np.minimum(first, len(seg) - 1)], first)
                for p0, p1 in zip(ps[:-1], ps[1:]):
                    inside = G.point_in_ring(x[sl], y[sl], seg[p0:p1]) & solid[sl]
                    first = np.where(inside & (p0 < first), p0, first)
                any_s = first < np.iinfo(np.int32).max
            hit_segment[sl] = np.where(any_s, first, -1).astype(np.int16)
            flags[sl] |= np.where(any_s, F_STATIC, 0).astype(np.uint8)

    if bounds is not None:
        ex, ey = G.extents(c, s, hl, hw)
        ex = np.where(circ, r, ex)
        ey = np.where(circ, r, ey)
        ob = G.out_of_bound(x, y, ex, ey, bounds) & solid
        flags |= np.where(ob, F_OUTBOUND, 0).astype(np.uint8)
    return flags, hit_index, hit_segment


def status(flags, type_id, step_count_new, max_step=0, ego_only=True):
    """check_status priority chain, parking.py:361-392 -> (scenario_status, done) uint8 [N].

    ``ego_only``: only participant 0 decides (the reference's single ego);
    otherwise any active participant's event ends the scenario."""
    active = type_id != INACTIVE
    f = np.where(active, flags, 0)
    if ego_only:
        agg = f[:, 0]
    else:
        agg = np.bitwise_or.reduce(f, axis=1)
    st = np.full(flags.shape[0], NORMAL, np.uint8)
    st = np.where((agg & F_DYNAMIC) != 0, FAILED, st)
    st = np.where((agg & F_STATIC) != 0, FAILED, st)
    st = np.where((agg & F_OUTBOUND) != 0, OUT_BOUND, st)
    if max_step and max_step > 0:
        st = np.where(step_count_new > max_step, TIME_EXCEEDED, st)  # time_exceed.py:31-33
    st = st.astype(np.uint8)
    return st, (st != NORMAL).astype(np.uint8)


def tick(state, type_id, action, table, step_count, segments=None, bounds=None, interval=100,
         delta_t=5, max_step=0, ego_only=True, steer_first=False):
    """One full tick in float64 end to end (free-running oracle)."""
    new = physics_tick(state, type_id, action, table, interval, delta_t, steer_first)
    flags, hi, hs = events(new["x"], new["y"], new["heading"], type_id, table, segments, bounds)
    cnt = np.asarray(step_count) + 1
    st, done = status(flags, type_id, cnt, max_step, ego_only)
    return new, flags, hi, hs, st, done, cnt


def goal_events(x, y, heading, type_id, table, target, last_pose, count, threshold=0.95, no_action_max=100):
    """Arrival.update (arrival.py:32-47) and NoAction.update (no_action.py:32-53) for the ego (participant 0) of
    every scenario, on the given (new) poses.  ``last_pose``: [N, 4] (x, y, heading, valid), ``count``: [N].
    Returns (arrived bool [N], no_action bool [N], iou [N], new_last_pose, new_count)."""
    N = x.shape[0]
    p = _gather(table, type_id)
    arrived = np.zeros(N, bool)
    noact = np.zeros(N, bool)
    iou = np.zeros(N, np.float64)
    new_last = np.array(last_pose, dtype=np.float64)
    new_count = np.array(count, dtype=np.int64)
    for n in range(N):
        if type_id[n, 0] == INACTIVE or p["shape"][n, 0] != OBB:
            continue
        ex, ey, eh = float(x[n, 0]), float(y[n, 0]), float(heading[n, 0])
        el, ew = float(p["half_len"][n, 0]), float(p["half_wid"][n, 0])
        if no_action_max > 0:
            if last_pose[n, 3] != 0:
                i0 = G.rect_iou(ex, ey, eh, el, ew, float(last_pose[n, 0]), float(last_pose[n, 1]), float(last_pose[n, 2]), el, ew)
                new_count[n] = new_count[n] + 1 if i0 > 0.999 else 0
            noact[n] = new_count[n] > no_action_max
        new_last[n] = (ex, ey, eh, 1.0)
        t = [float(v) for v in target[n]]
        iou[n] = G.rect_iou(ex, ey, eh, el, ew, *t)
        arrived[n] = iou[n] >= threshold
    return arrived, noact, iou, new_last, new_count


def status_with_goal(flags, type_id, step_count_new, arrived, noact, max_step=0, ego_only=True):
    """The full check_status chain (parking.py:361-392); later assignments have higher priority:
    completed < collision < out of bound < no action < time exceeded."""
    active = type_id != INACTIVE
    f = np.where(active, flags, 0)
    agg = f[:, 0] if ego_only else np.bitwise_or.reduce(f, axis=1)
    st = np.full(flags.shape[0], NORMAL, np.uint8)
    st = np.where(arrived, COMPLETED, st)                       # parking.py:387-390
    st = np.where((agg & (F_DYNAMIC | F_STATIC)) != 0, FAILED, st)   # :381-385
    st = np.where((agg & F_OUTBOUND) != 0, OUT_BOUND, st)       # :376-379
    st = np.where(noact, NO_ACTION, st)                         # :371-374
    if max_step and max_step > 0:
        st = np.where(step_count_new > max_step, TIME_EXCEEDED, st)   # :366-369
    st = st.astype(np.uint8)
    return st, (st != NORMAL).astype(np.uint8)


def env_epilogue(flags, scn_status, step_count, max_step, iou=None, ego_xy=None, target=None, max_iou=None, min_dist=None,
                 reset_trackers=True):
    """``ParkingEnv.step`` after ``check_status`` (envs/parking.py:240-256) with ``_get_reward`` (:148-190), per scenario, in
    float64.  ``flags`` [N, M] event bytes, ``scn_status`` [N]; with a goal: ``iou`` [N], ``ego_xy`` [N, 2], ``target``
    [N, 5] and the per-episode extrema ``max_iou`` / ``min_dist`` [N] (updated copies are returned).
    Returns dict(reward, terminated, truncated, done, traffic_status, max_iou, min_dist)."""
    flags = np.asarray(flags)
    N = flags.shape[0]
    traffic = np.where(flags & F_STATIC, 3, np.where(flags & F_DYNAMIC, 4, 1)).astype(np.uint8)   # status.py:52-61
    reward = np.zeros(N, np.float64)
    term = np.zeros(N, bool)
    trunc = np.zeros(N, bool)
    max_iou = None if max_iou is None else np.array(max_iou, np.float64)
    min_dist = None if min_dist is None else np.array(min_dist, np.float64)
    for n in range(N):
        st = int(scn_status[n])
        ego_ts = int(traffic[n, 0]) if st == FAILED else 1      # check_status returns at the first detector that fires
        term[n] = st == COMPLETED                                # :243-244
        trunc[n] = (not term[n]) and (st != NORMAL or ego_ts != 1)   # :245-248
        if ego_ts in (3, 4):
            r = -5.0                                             # :151-152
        elif st in (TIME_EXCEEDED, NO_ACTION):
            r = -1.0                                             # :153-157
        elif st == OUT_BOUND:
            r = -5.0                                             # :158-159
        elif st == COMPLETED:
            r = 5.0                                              # :160-161
        else:
            r = -np.tanh(float(step_count[n]) / max_step) * 0.001 if max_step and max_step > 0 else 0.0   # :163
            if iou is not None and max_iou is not None:
                r += float(iou[n]) if max_iou[n] == -np.inf else float(iou[n]) - max_iou[n]   # :164-169
                max_iou[n] = max(max_iou[n], float(iou[n]))                                   # :170
            if target is not None and min_dist is not None:
                d = float(np.hypot(ego_xy[n, 0] - target[n, 0], ego_xy[n, 1] - target[n, 1]))   # :172-185
                if d < min_dist[n]:
                    if np.isfinite(min_dist[n]):
                        r += (min_dist[n] - d) * 0.1             # :186-187
                    min_dist[n] = d                              # :188
        reward[n] = r
        if reset_trackers and (term[n] or trunc[n]):
            if max_iou is not None:
                max_iou[n] = -np.inf
            if min_dist is not None:
                min_dist[n] = np.inf
    return dict(reward=reward, terminated=term, truncated=trunc, done=(term | trunc).astype(np.uint8), traffic_status=traffic,
                max_iou=max_iou, min_dist=min_dist)
