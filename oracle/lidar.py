"""Float64 restatement of the reference's single-line lidar scan (TEST INFRASTRUCTURE ONLY).

``SingleLineLidar._scan_obstacles`` (tactics2d/sensor/lidar.py:128-221), vectorised NumPy exactly as the reference
writes it: obstacle rings -> ego frame (``_rotate_and_filter_obstacles``, :105-126) -> edges -> determinant
intersection with the beam lines -> the four 1e-8-slack box filters -> min distance, clip, range -> inf.

PINNED to the unmodified reference: ``oracle/make_golden.py`` loads the reference's ``sensor/lidar.py`` by path (with
stand-ins for the shapely ring container, ``affine_transform`` and ``Map``, the only things it needs from packages that
are not installable here), runs ``_scan_obstacles`` on 12 scenes x 4 beam / range settings and writes
``tests/golden/lidar.npz``; ``tests/test_oracle_lidar.py`` holds this restatement to those scans at 1e-12.  (No reference
test pins a scan: tests/test_sensor.py needs data files that are not in the reference tree.)  The arithmetic below
restates the NumPy part; the shapely part (``affine_transform``, ``distance``) is replaced by its definition.  Obstacles are passed as closed rings [K+1, 2] or open polylines [2, 2]
(a map segment); the distance filter of :122-125 only prunes whole obstacles and never changes the scan.
"""

from __future__ import annotations

import numpy as np


def scan(ego, rings, n_beams: int, max_range: float):
    """ego = (x, y, heading); rings = list of [P, 2] float64 point arrays (consecutive points are edges)."""
    x0, y0, theta0 = ego
    a0, b0 = np.cos(theta0), np.sin(theta0)
    x_off = -x0 * a0 - y0 * b0           # :118-119
    y_off = x0 * b0 - y0 * a0
    x1s, x2s, y1s, y2s = [], [], [], []
    for ring in rings:
        r = np.asarray(ring, dtype=np.float64)
        rx = a0 * r[:, 0] + b0 * r[:, 1] + x_off   # affine [a, b, -b, a, xoff, yoff]  :120
        ry = -b0 * r[:, 0] + a0 * r[:, 1] + y_off
        x1s.extend(rx[:-1]); x2s.extend(rx[1:]); y1s.extend(ry[:-1]); y2s.extend(ry[1:])
    if len(x1s) == 0:
        return np.full(n_beams, np.inf)
    theta = np.linspace(0, 2 * np.pi, n_beams, endpoint=False)   # :160
    a = np.sin(theta).reshape(-1, 1)
    b = -np.cos(theta).reshape(-1, 1)
    c = 0
    x1s, x2s, y1s, y2s = (np.array(v).reshape(1, -1) for v in (x1s, x2s, y1s, y2s))
    d = (y2s - y1s)                                               # :183-185
    e = (x1s - x2s)
    f = (y1s * x2s - x1s * y2s)
    det = a * e - b * d                                           # :188
    parallel = det == 0
    det = np.where(parallel, 1.0, det)
    raw_x = (b * f - c * e) / det
    raw_y = (c * d - a * f) / det
    tmp_inf = max_range * 10
    tz = 1e-8
    lx = (np.cos(theta) * max_range).reshape(-1, 1)
    ly = (np.sin(theta) * max_range).reshape(-1, 1)
    raw_x = np.where(raw_x > np.maximum(tz, lx) + tz, tmp_inf, raw_x)    # :201-204
    raw_x = np.where(raw_x < np.minimum(-tz, lx) - tz, tmp_inf, raw_x)
    raw_y = np.where(raw_y > np.maximum(tz, ly) + tz, tmp_inf, raw_y)
    raw_y = np.where(raw_y < np.minimum(-tz, ly) - tz, tmp_inf, raw_y)
    raw_x = np.where(raw_x > np.maximum(x1s, x2s) + tz, tmp_inf, raw_x)  # :206-209
    raw_x = np.where(raw_x < np.minimum(x1s, x2s) - tz, tmp_inf, raw_x)
    raw_y = np.where(raw_y > np.maximum(y1s, y2s) + tz, tmp_inf, raw_y)
    raw_y = np.where(raw_y < np.minimum(y1s, y2s) - tz, tmp_inf, raw_y)
    raw_x = np.where(parallel, tmp_inf, raw_x)                            # :211
    obs = np.min(np.sqrt(raw_x**2 + raw_y**2), axis=1)
    obs = np.clip(obs, 0, max_range)
    return np.where(obs == max_range, np.inf, obs)


def scan_world(x, y, heading, type_id, table, segments, n_beams, max_range):
    """Ego (participant 0) scans of every scenario of a batched world (float64 inputs = the fp32 device state)."""
    from . import geometry as G
    from .scenario import INACTIVE, OBB

    N, M = x.shape
    out = np.full((N, n_beams), np.inf)
    seg_rings = [] if segments is None else [np.asarray(s, dtype=np.float64).reshape(2, 2) for s in segments]
    for n in range(N):
        if type_id[n, 0] == INACTIVE:
            continue
        rings = list(seg_rings)
        for j in range(1, M):
            t = type_id[n, j]
            if t == INACTIVE or table["shape"][t] != OBB:
                continue
            c = G.obb_corners(float(x[n, j]), float(y[n, j]), float(heading[n, j]), float(table["half_len"][t]), float(table["half_wid"][t]))
            rings.append(np.concatenate([c, c[:1]], 0))
        out[n] = scan((float(x[n, 0]), float(y[n, 0]), float(heading[n, 0])), rings, n_beams, max_range)
    return out
