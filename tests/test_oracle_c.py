"""The C oracle (oracle/c/oracle_tick.c) agrees with the NumPy oracle, hence with the reference."""

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import scenario as O
from tactics2d_b200 import synthetic


@pytest.mark.parametrize("maker,kw", [
    (synthetic.config2, dict(n=24, m=64, seed=2, size=90.0)),
    (synthetic.config3, dict(n=8, m=32, seed=1)),
    (synthetic.config4, dict(n=40, m=32, seed=3, size=50.0)),
])
@pytest.mark.parametrize("interval,delta_t", [(100, 5), (33, 10)])
def test_c_oracle_matches_numpy(maker, kw, interval, delta_t):
    scene = synthetic.with_inactive(maker(**kw), 0.1, seed=1)
    if scene.segments is None:
        scene.segments = synthetic.grid_wall_segments(50.0, 25.0, 9.0)
    table = scene.table.as_oracle_table()
    act = synthetic.random_actions(5, scene.shape, accel=(-4, 4), steer=(-0.7, 0.7))
    a = CO.physics(scene.state(), scene.type_id, act, table, interval, delta_t)
    b = O.physics_tick(scene.state(), scene.type_id, act, table, interval, delta_t)
    for k in b:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12, atol=1e-12, err_msg=k)
    fa = CO.events(b["x"], b["y"], b["heading"], scene.type_id, table, scene.segments, scene.bounds)
    fb = O.events(b["x"], b["y"], b["heading"], scene.type_id, table, scene.segments, scene.bounds)
    for u, v in zip(fa, fb):
        assert np.array_equal(u, v)
    assert (fb[0] & 1).any()


def test_c_oracle_steer_first():
    scene = synthetic.config4(4, 32, seed=3, size=50.0)
    table = scene.table.as_oracle_table()
    act = synthetic.random_actions(5, scene.shape)
    a = CO.physics(scene.state(), scene.type_id, act, table, steer_first=True)
    b = O.physics_tick(scene.state(), scene.type_id, act, table, steer_first=True)
    for k in b:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12, atol=1e-12)
