"""The lidar restatement (oracle/lidar.py) against scans of the unmodified reference's ``SingleLineLidar._scan_obstacles``
(tests/golden/lidar.npz, written by oracle/make_golden.py: the reference's own NumPy ray / edge arithmetic, run with
stand-ins for the shapely ring container and ``affine_transform``)."""

import os

import numpy as np
import pytest

from oracle import geometry as G
from oracle import lidar as OL

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lidar.npz"))


def _rings(k):
    rings = [np.concatenate([w, w[:1]], 0) for w in GOLD["walls"]]
    for x, y, h, hl, hw in GOLD["others"][k]:
        c = G.obb_corners(x, y, h, hl, hw)
        rings.append(np.concatenate([c, c[:1]], 0))
    return rings


@pytest.mark.parametrize("n_beams,max_range", [(360, 20.0), (500, 12.0), (37, 30.0), (1100, 9.0)])
def test_scan_equals_reference(n_beams, max_range):
    want = GOLD[f"scan_{n_beams}_{int(max_range)}"]
    hits = 0
    for k, ego in enumerate(GOLD["ego"]):
        got = OL.scan(tuple(ego), _rings(k), n_beams, max_range)
        assert np.array_equal(np.isinf(got), np.isinf(want[k]))
        ok = np.isfinite(want[k])
        # (the reference drops obstacles farther than the range before the scan, :122-125; that never changes a beam)
        np.testing.assert_allclose(got[ok], want[k][ok], rtol=1e-12, atol=1e-12)
        hits += int(ok.sum())
    assert hits > 0.2 * want.size


def test_scan_world_matches_per_ego_scan():
    """``scan_world`` (what the GPU tests call) = the per-ego ``scan`` on the same poses."""
    k, n_beams, max_range = 4, 360, 20.0
    ego, others = GOLD["ego"][k], GOLD["others"][k]
    m = 1 + len(others)
    x = np.concatenate([[ego[0]], others[:, 0]])[None]
    y = np.concatenate([[ego[1]], others[:, 1]])[None]
    h = np.concatenate([[ego[2]], others[:, 2]])[None]
    # a type table with one row per participant (each has its own size)
    table = dict(shape=np.zeros(m, np.int32), half_len=np.concatenate([[2.0], others[:, 3]]), half_wid=np.concatenate([[0.9], others[:, 4]]))
    tid = np.arange(m, dtype=np.uint8)[None]
    segs = np.concatenate([np.concatenate([w, np.roll(w, -1, 0)], 1) for w in GOLD["walls"]], 0)     # ring edges as segments
    got = OL.scan_world(x, y, h, tid, table, segs, n_beams, max_range)[0]
    want = GOLD["scan_360_20"][k]
    assert np.array_equal(np.isinf(got), np.isinf(want))
    ok = np.isfinite(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-12)
