"""On-device NPC controllers (t2d_control) against the unmodified reference's outputs (tests/golden/controllers.npz) and
against the float64 restatement on batched, multi-tick scenes."""

import os

import numpy as np
import pytest

from oracle import controllers as OC
from oracle import scenario as O

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "controllers.npz"))
IDM_KEYS = ("desired_speed", "time_headway", "min_spacing", "max_acceleration", "comfortable_deceleration", "delta")
ACC_KEYS = ("target_speed", "kp", "accel_change_rate", "delta_t", "max_accel", "min_accel", "interval")
PP_KEYS = ACC_KEYS + ("min_pre_aiming_distance", "pp_interval", "wheel_base")


def _row(kind, **kw):
    from tactics2d_b200 import _lib

    return _lib.ControllerParamsC(kind=kind, **{k: float(v) for k, v in kw.items()})


def _pair_world(device, steer_first=False):
    """256 scenarios x (ego, leader) holding the golden file's states."""
    import torch  # noqa: F401

    from tactics2d_b200 import BatchedWorld
    from tactics2d_b200.types import TypeParams, TypeTable

    e, l = G["ego"], G["lead"]
    n = len(e)
    w = BatchedWorld(n, 2, TypeTable([TypeParams()]), device=device, steer_first=steer_first)
    st = np.stack([e, l], 1)                        # [n, 2, 5]
    w.set_state(st[..., 0], st[..., 1], st[..., 2], st[..., 3], type_id=np.zeros((n, 2), np.uint8))
    return w, np.abs(st[..., 4]).astype(np.float32)   # State.accel = |a|


def _run(w, rows, last_accel, follow, path_id=None):
    import torch

    n = w.N
    w.set_controllers(rows, ctrl_id=np.tile(np.array([[0, 255]], np.uint8), (n, 1)),
                      lead_index=np.tile(np.array([[1 if follow else -1, -1]], np.int16), (n, 1)),
                      path_id=None if path_id is None else np.stack([path_id, np.full(n, -1)], 1).astype(np.int16),
                      last_accel=last_accel)
    marker = torch.full((n, 2, 2), 7.5, dtype=torch.float32, device=w.device)
    act = w.control(marker).cpu().numpy()
    assert np.all(act[:, 1] == 7.5)                  # the uncontrolled participant keeps the caller's action
    return act[:, 0]


def _close(got, want, rtol=2e-5, atol=2e-6):
    want = np.asarray(want, np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok].astype(np.float64), want[ok], rtol=rtol, atol=atol)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_idm_equals_reference_outputs(cuda_device, ci):
    w, la = _pair_world(cuda_device)
    rows = [_row(1, **dict(zip(IDM_KEYS, G["idm_cfgs"][ci])))]
    free = _run(w, rows, la, follow=False)
    _close(free[:, 0], G[f"idm{ci}_free"])
    assert np.all(free[:, 1] == 0.0)                 # IDM never steers (idm_controller.py:92)
    _close(_run(w, rows, la, follow=True)[:, 0], G[f"idm{ci}_follow"])


@pytest.mark.parametrize("si", [0, 1, 2, 3])
def test_acceleration_controller_equals_reference_outputs(cuda_device, si):
    w, la = _pair_world(cuda_device)
    rows = [_row(2, **dict(zip(ACC_KEYS, G["acc_cfgs"][si])))]
    _close(_run(w, rows, la, follow=False)[:, 0], G[f"acc{si}_cruise"])
    _close(_run(w, rows, la, follow=True)[:, 0], G[f"acc{si}_follow"])


@pytest.mark.parametrize("si", [0, 1])
def test_pure_pursuit_equals_reference_outputs(cuda_device, si):
    w, la = _pair_world(cuda_device, steer_first=True)
    w.set_paths([G[f"path{k}"] for k in range(3)])
    rows = [_row(3, **dict(zip(PP_KEYS, G["pp_cfgs"][si])))]
    a = _run(w, rows, la, follow=False, path_id=G["path_id"])
    _close(a[:, 0], G[f"pp{si}_steer"], rtol=2e-5, atol=1e-5)      # (steer, accel): the env's action order
    _close(a[:, 1], G[f"pp{si}_accel"])
    _close(_run(w, rows, la, follow=True, path_id=G["path_id"])[:, 1], G[f"pp{si}_accel_follow"])


@pytest.mark.parametrize("n,m,steer_first", [(48, 32, False), (9, 100, True), (5, 3, False)])
def test_control_then_step_rollout_matches_restatement(cuda_device, n, m, steer_first):
    """control -> step for several ticks on a mixed scene: actions of the controlled participants and last_accel of
    everybody equal the restatement evaluated on the state the GPU holds; uncontrolled rows are untouched."""
    import torch

    from tactics2d_b200 import BatchedWorld, synthetic
    from tactics2d_b200.controller import AccelerationController, IDMController, PurePursuitController

    scene = synthetic.with_inactive(synthetic.config4(n, m, seed=5), 0.15, seed=6) if m > 3 else synthetic.config4(n, m, seed=5)
    w = BatchedWorld(n, m, scene.table, device=cuda_device, steer_first=steer_first)
    w.set_map(scene.segments, scene.bounds)
    w.set_state(scene.x, scene.y, scene.heading, scene.speed, vx=scene.vx, vy=scene.vy, type_id=scene.type_id)
    rng = np.random.default_rng(11)
    pp = PurePursuitController(min_pre_aiming_distance=6.0, target_speed=9.0)
    pp.update_driving_style(-0.4)
    acc = AccelerationController(target_speed=12.0)
    ctrls = [IDMController(), IDMController(desired_speed=25.0, max_acceleration=2.0), acc, pp]
    paths = [np.array([[0, 0], [50, 10], [120, 10]], np.float32), np.array([[10, -40], [10, 90], [-60, 160], [-60, 400]], np.float32)]
    w.set_paths(paths)
    ctrl_id = rng.choice([255, 0, 1, 2, 3], size=(n, m), p=[0.2, 0.2, 0.2, 0.2, 0.2]).astype(np.uint8)
    ctrl_id[:, 0] = 255                                               # the ego is driven by the caller
    lead = rng.integers(-1, m, size=(n, m)).astype(np.int16)
    pid = rng.integers(-1, len(paths), size=(n, m)).astype(np.int16)
    w.set_controllers(ctrls, ctrl_id, lead, pid)
    table = scene.table.as_oracle_table()
    ctab = []
    for c in ctrls:
        r = c.params()
        ctab.append({k: getattr(r, k) for k, _ in r._fields_})
    la = np.zeros((n, m), np.float32)
    controlled = (ctrl_id != 255) & (scene.type_id != 255)
    for t in range(5):
        ext = synthetic.random_actions(900 + t, (n, m))
        before = w.state_numpy()
        want_act, want_la = OC.control_tick(before, scene.type_id, table, ext, ctrl_id, ctab, lead, pid,
                                            [p.astype(np.float64) for p in paths], la, steer_first)
        act = w.control(torch.from_numpy(ext).to(cuda_device))
        got = act.cpu().numpy()
        assert np.array_equal(got[~controlled], ext[~controlled])
        np.testing.assert_allclose(got[controlled], want_act[controlled], rtol=3e-6, atol=3e-6)
        got_la = w.last_accel.cpu().numpy()
        np.testing.assert_allclose(got_la, want_la, rtol=3e-6, atol=3e-6)
        la = got_la
        w.step(act)
    assert controlled.sum() > 0


def test_single_state_controller_api(cuda_device):
    """The reference's per-object call: controller.step(ego_state, ...) -> (steering, acceleration)."""
    from tactics2d_b200.controller import AccelerationController, IDMController, PurePursuitController
    from tactics2d_b200.participant.trajectory import State

    e, l = G["ego"], G["lead"]
    idm = IDMController()
    for i in (0, 30, 77):
        s, a = idm.step(State(0, x=e[i, 0], y=e[i, 1], heading=e[i, 2], speed=e[i, 3]))
        assert s == 0.0 and a == pytest.approx(G["idm0_free"][i], rel=2e-5, abs=2e-6)
        s, a = idm.step(State(0, x=e[i, 0], y=e[i, 1], heading=e[i, 2], speed=e[i, 3]),
                        State(0, x=l[i, 0], y=l[i, 1], heading=l[i, 2], speed=l[i, 3]))
        assert a == pytest.approx(G["idm0_follow"][i], rel=2e-5, abs=2e-6)
    acc = AccelerationController(target_speed=8.0)
    i = 40
    ego = State(0, x=e[i, 0], y=e[i, 1], heading=e[i, 2], speed=e[i, 3], accel=e[i, 4])
    front = State(0, x=l[i, 0], y=l[i, 1], heading=l[i, 2], speed=l[i, 3], accel=l[i, 4])
    assert acc.step(ego)[1] == pytest.approx(G["acc0_cruise"][i], rel=2e-5, abs=2e-6)
    assert acc.step(ego, front_state=front)[1] == pytest.approx(G["acc0_follow"][i], rel=2e-5, abs=2e-6)
    with pytest.raises(TypeError):
        acc.step(State(0, x=0.0, y=0.0, heading=0.0, speed=1.0))     # no acceleration in the state
    with pytest.raises(TypeError):
        acc.step(ego, front_state="not a state")
    pp = PurePursuitController()
    k = int(G["path_id"][i])
    steer, a = pp.step(ego, G[f"path{k}"], wheel_base=2.637)
    assert steer == pytest.approx(G["pp0_steer"][i], rel=2e-5, abs=1e-5)
    assert a == pytest.approx(G["pp0_accel"][i], rel=2e-5, abs=2e-6)
    with pytest.raises(ValueError):
        PurePursuitController(min_pre_aiming_distance=0.0)
    with pytest.raises(ValueError):
        AccelerationController(target_speed=-1.0)
    with pytest.raises(AttributeError):
        idm.configure(no_such_parameter=1.0)
    acc.update_driving_style(5.0)                                     # clamps to the aggressive end values
    assert (acc.kp, acc.max_accel, acc.min_accel, acc.interval) == (2.5, 2.5, -5.0, 1.5)
