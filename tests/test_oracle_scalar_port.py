"""The per-agent scalar port (CPU baseline) agrees with the vectorised oracle."""

import numpy as np

from oracle import scalar_port as SP
from oracle import scenario as O
from tactics2d_b200 import synthetic


def _check(scene, interval=100, delta_t=5):
    act = synthetic.random_actions(3, scene.shape, accel=(-4, 3), steer=(-0.7, 0.7))
    table = scene.table.as_oracle_table()
    new, fl, hi, hs = SP.tick_scenarios(scene.state(), scene.type_id, act, table, scene.segments, scene.bounds, interval, delta_t)
    ref = O.physics_tick(scene.state(), scene.type_id, act, table, interval, delta_t)
    for k in ref:
        np.testing.assert_allclose(new[k], ref[k], rtol=1e-12, atol=1e-12, err_msg=k)
    rfl, rhi, rhs = O.events(ref["x"], ref["y"], ref["heading"], scene.type_id, table, scene.segments, scene.bounds)
    assert np.array_equal(fl, rfl) and np.array_equal(hi, rhi) and np.array_equal(hs, rhs)
    return fl


def test_scalar_port_kinematics_gridmap():
    fl = _check(synthetic.config2(6, 24, seed=2, size=40.0))
    assert (fl & 1).any() and (fl & 2).any()


def test_scalar_port_mixed_and_ragged():
    scene = synthetic.with_inactive(synthetic.config4(5, 16, seed=3, size=30.0, segments=synthetic.grid_wall_segments(30.0, 15.0, 6.0)), 0.15)
    _check(scene, 50, 3)


def test_scalar_port_dynamics():
    _check(synthetic.config3(2, 16, seed=1))
