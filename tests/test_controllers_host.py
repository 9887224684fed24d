"""Host-side behaviour of the controller facades, holding them to the facts the reference's own
tests/test_controllers.py checks (construction defaults and validation, driving-style interpolation, configure) plus the
parameter rows they hand to the device table.  The laws themselves run on the GPU (tests/test_gpu_controllers.py)."""

import pytest

from tactics2d_b200.controller import AccelerationController, ControllerBase, IDMController, PurePursuitController
from tactics2d_b200.controller.controller_base import CTRL_CRUISE, CTRL_IDM, CTRL_PURE_PURSUIT


def test_controller_base_is_abstract_and_interpolates_styles():      # reference: TestControllerBase
    with pytest.raises(TypeError):
        ControllerBase()

    class Concrete(ControllerBase):
        def step(self, ego_state, **kwargs):
            return 0.0, 0.0

    f = Concrete().create_style_interpolator(1.0, 2.0)
    assert (f(-1.0), f(1.0), f(0.0)) == (1.0, 2.0, 1.5)
    assert (f(-2.0), f(2.0)) == (1.0, 2.0)                              # outside [-1, 1]: the end values
    with pytest.raises(AttributeError, match="has no parameter"):
        Concrete().configure(nothing=1)


def test_acceleration_controller_defaults_styles_and_row():         # reference: TestAccelerationController
    c = AccelerationController()
    assert (c.target_speed, c.kp, c.max_accel, c.min_accel, c.accel_change_rate, c.interval, c.delta_t) == (5.0, 3.5, 1.5, -4.0, 3.0, 2.0, 0.05)
    assert AccelerationController(target_speed=10.0).target_speed == 10.0
    with pytest.raises(ValueError, match="target_speed must be non-negative"):
        AccelerationController(target_speed=-1.0)
    c.update_driving_style(-1.0)
    assert (c.kp, c.speed_factor, c.max_accel, c.min_accel, c.interval) == (4.5, 0.8, 1.5, -3.0, 3.5)
    c.update_driving_style(1.0)
    assert (c.kp, c.speed_factor, c.accel_change_rate, c.max_accel, c.min_accel, c.interval) == (2.5, 1.2, 6.0, 2.5, -5.0, 1.5)
    with pytest.raises(TypeError):
        c.update_driving_style("fast")
    row = c.params()
    assert row.kind == CTRL_CRUISE and (row.kp, row.max_accel, row.min_accel, row.interval) == (2.5, 2.5, -5.0, 1.5)
    assert row.target_speed == 5.0 and abs(row.delta_t - 0.05) < 1e-8


def test_idm_controller_defaults_configure_and_row():               # reference: TestIDMController
    c = IDMController()
    assert (c.desired_speed, c.time_headway, c.min_spacing, c.max_acceleration, c.comfortable_deceleration, c.delta) == (10.0, 1.5, 2.0, 1.0, 3.0, 4.0)
    c = IDMController(desired_speed=15.0, time_headway=2.0, min_spacing=3.0, max_acceleration=2.0, comfortable_deceleration=4.0, delta=2.0)
    assert (c.desired_speed, c.time_headway, c.min_spacing, c.max_acceleration, c.comfortable_deceleration, c.delta) == (15.0, 2.0, 3.0, 2.0, 4.0, 2.0)
    c.configure(desired_speed=12.0, max_acceleration=1.5)
    assert (c.desired_speed, c.max_acceleration) == (12.0, 1.5)
    with pytest.raises(AttributeError, match="has no parameter"):
        c.configure(invalid_param=1.0)
    row = c.params()
    assert row.kind == CTRL_IDM and (row.desired_speed, row.max_acceleration, row.delta) == (12.0, 1.5, 2.0)


def test_pure_pursuit_controller_defaults_styles_and_row():         # reference: TestPurePursuitController
    c = PurePursuitController()
    assert c.min_pre_aiming_distance == 10.0 and c.interval == 1.0 and c._longitudinal_control.target_speed == 5.0
    c = PurePursuitController(min_pre_aiming_distance=5.0, target_speed=8.0)
    assert c.min_pre_aiming_distance == 5.0 and c._longitudinal_control.target_speed == 8.0
    with pytest.raises(ValueError, match="min_pre_aiming_distance must be positive"):
        PurePursuitController(min_pre_aiming_distance=0)
    with pytest.raises(ValueError, match="target_speed must be non-negative"):
        PurePursuitController(target_speed=-1.0)
    c.update_driving_style(-1.0)
    assert c.interval == 2.0 and c._longitudinal_control.kp == 4.5     # the longitudinal law follows the style too
    c.update_driving_style(1.0)
    assert c.interval == 1.0
    row = c.params()
    assert row.kind == CTRL_PURE_PURSUIT and row.min_pre_aiming_distance == 5.0 and row.pp_interval == 1.0
    assert abs(row.wheel_base - 2.637) < 1e-6 and row.target_speed == 8.0 and row.kp == 2.5
