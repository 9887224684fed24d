"""The controller restatement (oracle/controllers.py) against outputs of the unmodified reference classes
(tests/golden/controllers.npz, written by oracle/make_golden.py)."""

import os

import numpy as np
import pytest

from oracle import controllers as C

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "controllers.npz"))
IDM_KEYS = ("desired_speed", "time_headway", "min_spacing", "max_acceleration", "comfortable_deceleration", "delta")
ACC_KEYS = ("target_speed", "kp", "accel_change_rate", "delta_t", "max_accel", "min_accel", "interval")
PP_KEYS = ACC_KEYS + ("min_pre_aiming_distance", "pp_interval", "wheel_base")


def _close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    np.testing.assert_allclose(a[ok], b[ok], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_idm_matches_reference(ci):
    p = dict(zip(IDM_KEYS, G["idm_cfgs"][ci]))
    e, l = G["ego"], G["lead"]
    _close(C.idm(e[:, 3], e[:, 0], e[:, 1], False, 0.0, 0.0, 0.0, p), G[f"idm{ci}_free"])
    _close(C.idm(e[:, 3], e[:, 0], e[:, 1], True, l[:, 3], l[:, 0], l[:, 1], p), G[f"idm{ci}_follow"])


@pytest.mark.parametrize("si", [0, 1, 2, 3])
def test_acceleration_controller_matches_reference(si):
    p = dict(zip(ACC_KEYS, G["acc_cfgs"][si]))
    e, l = G["ego"], G["lead"]
    # State.accel is the magnitude of the stored acceleration (state.py:171-185)
    _close(C.cruise(e[:, 3], np.abs(e[:, 4]), p), G[f"acc{si}_cruise"])
    _close(C.adaptive_cruise(e[:, 3], e[:, 0], e[:, 1], np.abs(e[:, 4]), l[:, 3], l[:, 0], l[:, 1], np.abs(l[:, 4]), p),
           G[f"acc{si}_follow"])


@pytest.mark.parametrize("si", [0, 1])
def test_pure_pursuit_matches_reference(si):
    p = dict(zip(PP_KEYS, G["pp_cfgs"][si]))
    e, l, pid = G["ego"], G["lead"], G["path_id"]
    paths = [G[f"path{k}"] for k in range(3)]
    steer = [C.pure_pursuit_steering(r[0], r[1], r[2], r[3], paths[k], p) for r, k in zip(e, pid)]
    _close(steer, G[f"pp{si}_steer"])
    _close(C.cruise(e[:, 3], np.abs(e[:, 4]), p), G[f"pp{si}_accel"])
    _close(C.adaptive_cruise(e[:, 3], e[:, 0], e[:, 1], np.abs(e[:, 4]), l[:, 3], l[:, 0], l[:, 1], np.abs(l[:, 4]), p),
           G[f"pp{si}_accel_follow"])


def test_interpolate_walks_from_the_first_vertex_and_clamps():
    path = np.array([[0.0, 0.0], [3.0, 4.0], [3.0, 10.0]])
    assert np.allclose(C.interpolate(path, 0.0), [0, 0])
    assert np.allclose(C.interpolate(path, 2.5), [1.5, 2.0])
    assert np.allclose(C.interpolate(path, 5.0), [3, 4])
    assert np.allclose(C.interpolate(path, 8.0), [3, 7])
    assert np.allclose(C.interpolate(path, 100.0), [3, 10])
