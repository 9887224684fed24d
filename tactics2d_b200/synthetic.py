"""Seeded synthetic scenario batches for the BASELINE.json configs (host-side NumPy).

The reference has no scenario generator for N x M multi-agent batches (its generators build one
parking lot / racing track, map/generator/*.py, through the global ``np.random``); these are the
synthetic inputs SURVEY.md section 8(d) specifies.  Everything is generated on the CPU from an
explicit seed and rounded to fp32, so the CPU oracle and the GPU see identical bits.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from .types import TYPE_INACTIVE, TypeTable


@dataclass
class Scene:
    table: TypeTable
    x: np.ndarray          # fp32 [N, M]
    y: np.ndarray
    heading: np.ndarray
    speed: np.ndarray
    vx: np.ndarray
    vy: np.ndarray
    type_id: np.ndarray    # uint8 [N, M]
    segments: Optional[np.ndarray] = None   # fp32 [S, 4]
    bounds: Optional[Tuple[float, float, float, float]] = None
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def shape(self):
        return self.x.shape

    def state(self) -> dict:
        return dict(x=self.x, y=self.y, heading=self.heading, speed=self.speed, vx=self.vx, vy=self.vy)


def _finish(table, x, y, h, v, tid, segments, bounds, name, **meta) -> Scene:
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    x, y, h, v = f(x), f(y), f(h), f(v)
    vx = f(v.astype(np.float64) * np.cos(h.astype(np.float64)))
    vy = f(v.astype(np.float64) * np.sin(h.astype(np.float64)))
    seg = None if segments is None else np.ascontiguousarray(segments, dtype=np.float32).reshape(-1, 4)
    return Scene(table, x, y, h, v, vx, vy, np.ascontiguousarray(tid, dtype=np.uint8), seg, bounds, name, meta)


def grid_wall_segments(size: float = 200.0, pitch: float = 50.0, wall: float = 16.0) -> np.ndarray:
    """The "synthetic grid map" of config 2: axis-aligned wall pieces of length ``wall`` centred on
    every edge of a ``pitch`` lattice over a ``size`` x ``size`` arena (gaps between pieces let
    traffic through).  The reference's GridMapGenerator is a cost grid, not geometry
    (map/generator/generate_grid_map.py:10-42), so this build defines the map itself."""
    n = int(round(size / pitch))
    segs = []
    for i in range(n + 1):
        c = i * pitch
        for j in range(n):
            m = (j + 0.5) * pitch
            segs.append((c, m - wall / 2, c, m + wall / 2))   # vertical piece on x = c
            segs.append((m - wall / 2, c, m + wall / 2, c))   # horizontal piece on y = c
    return np.asarray(segs, dtype=np.float32)


def random_actions(seed: int, shape, accel=(-4.0, 3.0), steer=(-0.6, 0.6)) -> np.ndarray:
    """Per-step actions [N, M, 2] = (accel, steer); the ranges exceed the models' limits on purpose
    so that clipping is exercised."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(accel[0], accel[1], shape)
    d = rng.uniform(steer[0], steer[1], shape)
    return np.ascontiguousarray(np.stack([a, d], -1), dtype=np.float32)


def config1(seed: int = 0) -> Scene:
    """C1 (parity gate): 1 scenario x 8 medium cars, SingleTrackKinematics, empty map, bounds
    (-100, 100, -100, 100); 2 x 4 lattice with 6 m pitch and +-1 m jitter so that pairs overlap."""
    rng = np.random.default_rng(seed)
    table = TypeTable.from_templates("kinematics")
    ix, iy = np.meshgrid(np.arange(4), np.arange(2))
    x = (ix.reshape(1, 8) - 1.5) * 6.0 + rng.uniform(-1, 1, (1, 8))
    y = (iy.reshape(1, 8) - 0.5) * 6.0 + rng.uniform(-1, 1, (1, 8))
    h = rng.uniform(0, 2 * np.pi, (1, 8))
    v = rng.uniform(0, 10, (1, 8))
    tid = np.full((1, 8), table.index("medium_car"))
    return _finish(table, x, y, h, v, tid, None, (-100.0, 100.0, -100.0, 100.0), "C1 1x8 kinematics, empty map")


def _arena(rng, n, m, size, jitter, vmax, table, type_choices, heading=None):
    side = int(np.ceil(np.sqrt(m)))
    pitch = size / side
    jitter = pitch / 2 if jitter is None else jitter
    k = np.arange(m)
    gx, gy = (k % side + 0.5) * pitch, (k // side + 0.5) * pitch
    x = gx[None] + rng.uniform(-jitter, jitter, (n, m))
    y = gy[None] + rng.uniform(-jitter, jitter, (n, m))
    h = rng.uniform(0, 2 * np.pi, (n, m)) if heading is None else heading
    v = rng.uniform(0, vmax, (n, m))
    tid = rng.choice(np.asarray(type_choices), size=(n, m))
    return x, y, h, v, tid


def config2(n: int = 4096, m: int = 64, seed: int = 1, size: float = 200.0, jitter: float = None) -> Scene:
    """C2: N x M SingleTrackKinematics vehicles (types sampled from the 9 VEHICLE_TEMPLATE rows) in a
    200 m arena with the synthetic grid map; jitter tuned for a few percent of colliding participants."""
    rng = np.random.default_rng(seed)
    table = TypeTable.vehicles("kinematics")   # the 9 VEHICLE_TEMPLATE rows: a kinematics-only table
    x, y, h, v, tid = _arena(rng, n, m, size, jitter, 15.0, table, list(range(9)))
    return _finish(table, x, y, h, v, tid, grid_wall_segments(size), (-8.0, size + 8.0, -8.0, size + 8.0),
                   f"C2 {n}x{m} kinematics + OBB collision, synthetic grid map")


def config3(n: int = 4096, m: int = 64, seed: int = 3, segments=None, bounds=None) -> Scene:
    """C3: SingleTrackDynamics vehicles driving along a highD-like straight road at 20-40 m/s (away
    from the stiff |v| < 0.5 region); ``segments`` = the map's collidable polylines (highD tiles)."""
    rng = np.random.default_rng(seed)
    table = TypeTable.from_templates("dynamics")
    if bounds is None:
        bounds = (0.0, 668.0, -30.0, 2.0)
    x0, x1, y0, y1 = bounds
    lanes = np.linspace(y0 + 4.0, y1 - 4.0, 8)
    slots = m // 8 + (m % 8 > 0)
    k = np.arange(m)
    x = (x0 + 10.0) + (k // 8 + 0.5)[None] * ((x1 - x0 - 20.0) / slots) + rng.uniform(-6, 6, (n, m))
    y = lanes[k % 8][None] + rng.uniform(-0.8, 0.8, (n, m))
    h = np.where(k % 8 < 4, 0.0, np.pi)[None] + rng.uniform(-0.03, 0.03, (n, m))
    h = np.mod(h, 2 * np.pi)
    v = rng.uniform(20, 40, (n, m))
    tid = rng.integers(0, 9, (n, m))
    return _finish(table, x, y, h, v, tid, segments, bounds, f"C3 {n}x{m} dynamics + map polylines")


def config4(n: int = 16384, m: int = 32, seed: int = 4, segments=None, bounds=None, size: float = 150.0) -> Scene:
    """C4: mixed traffic - 60 % vehicles (kinematics), 20 % cyclists (kinematics, lf = lr = L/2),
    20 % pedestrians (PointMass newton, disc of width/2) on an inD-like intersection."""
    rng = np.random.default_rng(seed)
    table = TypeTable.from_templates("kinematics")
    x, y, h, v, _ = _arena(rng, n, m, size, None, 8.0, table, [0])
    u = rng.uniform(0, 1, (n, m))
    tid = np.where(u < 0.6, rng.integers(0, 9, (n, m)), np.where(u < 0.8, rng.integers(9, 12, (n, m)), rng.integers(12, 16, (n, m))))
    v = np.where(tid >= 12, rng.uniform(0, 2.5, (n, m)), v)
    if bounds is None:
        bounds = (-8.0, size + 8.0, -8.0, size + 8.0)
    else:
        x = x + bounds[0]
        y = y + bounds[2]
    return _finish(table, x, y, h, v, tid, segments, bounds, f"C4 {n}x{m} mixed vehicle/cyclist/pedestrian")


def config5(n: int = 65536, m: int = 128, seed: int = 5, segments=None, bounds=None, size: float = 280.0) -> Scene:
    """C5: broadphase stress - N x 128 kinematic vehicles, rounD-like map."""
    rng = np.random.default_rng(seed)
    table = TypeTable.vehicles("kinematics")
    x, y, h, v, tid = _arena(rng, n, m, size, None, 15.0, table, list(range(9)))
    if bounds is None:
        bounds = (-8.0, size + 8.0, -8.0, size + 8.0)
    else:
        x = x + bounds[0]
        y = y + bounds[2]
    return _finish(table, x, y, h, v, tid, segments, bounds, f"C5 {n}x{m} kinematics + broadphase stress")


def with_inactive(scene: Scene, fraction: float, seed: int = 0) -> Scene:
    """Mark a random subset of slots inactive (ragged scenarios)."""
    rng = np.random.default_rng(seed)
    tid = scene.type_id.copy()
    tid[rng.uniform(0, 1, tid.shape) < fraction] = TYPE_INACTIVE
    return Scene(scene.table, scene.x, scene.y, scene.heading, scene.speed, scene.vx, scene.vy, tid, scene.segments,
                 scene.bounds, scene.name + f" ({fraction:.0%} inactive)", dict(scene.meta))
