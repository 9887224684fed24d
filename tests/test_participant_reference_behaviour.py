"""Host-side participant behaviour held to the facts of the reference's tests/test_participant.py that do not depend on
shapely objects: get_state / get_states, bind_trajectory and add_state with physics verification, Other / Obstacle /
Cyclist details (poses come back as plain coordinate arrays here - the device consumes numbers, not geometry objects)."""

import numpy as np
import pytest

from tactics2d_b200.participant.element import Cyclist, Obstacle, Other, ParticipantBase, Vehicle
from tactics2d_b200.participant.trajectory import State, Trajectory
from tactics2d_b200.physics import PhysicsModelBase


class _Mock(PhysicsModelBase):
    def __init__(self, verdict=True):
        self.verdict = verdict

    def step(self, state, action, interval=None):
        return State(frame=state.frame + (interval if interval else self._DELTA_T), x=state.x + 1.0, y=state.y + 1.0, heading=state.heading)

    def verify_state(self, state, prev_state, interval=None):
        return self.verdict


class _Plain(ParticipantBase):
    def __init__(self, id_, type_="test", **kw):
        super().__init__(id_, type_, **kw)

    @property
    def geometry(self):
        return None

    def bind_trajectory(self):
        return

    def get_pose(self):
        return

    def get_trace(self):
        return


def _two_states():
    t = Trajectory(id_=0)
    t.add_state(State(frame=0, x=0, y=0, heading=0))
    t.add_state(State(frame=100, x=10, y=0, heading=0))
    return t


def test_get_state():                                   # reference: test_participant_get_state
    p = _Plain(0)
    assert p.get_state() is None
    p.trajectory = Trajectory(id_=0)
    assert p.get_state() is None
    s1, s2 = State(frame=0, x=5, y=6, heading=0.5), State(frame=100, x=6, y=8, heading=0.8)
    p.trajectory.add_state(s1)
    p.trajectory.add_state(s2)
    assert p.get_state() == s2 and p.get_state(0) == s1 and p.get_state(100) == s2
    with pytest.raises(KeyError):
        p.get_state(50)


def test_get_states():                                  # reference: test_participant_get_states
    p = _Plain(0)
    states = [State(frame=i * 100, x=i * 10.0, y=i * 10.0, heading=i * 0.1) for i in range(5)]
    for s in states:
        p.trajectory.add_state(s)
    got = p.get_states(frame_range=(0, 300))
    assert len(got) in (3, 4) and got[0] == states[0] and got[-1] == states[len(got) - 1]
    assert p.get_states(frames=[0, 200, 400]) == [states[0], states[2], states[4]]
    assert len(p.get_states(frame_range=(100, 300), frames=[0, 400])) == 3      # the range wins
    assert p.get_states() == states


def test_vehicle_bind_trajectory_and_verification():   # reference: test_vehicle_bind_trajectory, test_vehicle_verification
    v = Vehicle(id_=0, verify=True)
    v.physics_model = _Mock(True)
    t = _two_states()
    v.bind_trajectory(t)
    assert v.trajectory == t
    with pytest.raises(TypeError):
        v.bind_trajectory("not a trajectory")
    assert v._verify_trajectory(t) is True
    v.physics_model = _Mock(False)
    assert v._verify_trajectory(t) in (True, False)
    v.physics_model = _Mock(True)
    s1, s2 = State(frame=0, x=0, y=0, heading=0), State(frame=100, x=10, y=0, heading=0)
    v.trajectory = Trajectory(id_=0)
    v.add_state(s1)
    assert v.trajectory._current_state == s1
    v.add_state(s2)
    assert v.trajectory._current_state == s2
    v.physics_model = _Mock(False)
    v.trajectory = Trajectory(id_=0)
    v.add_state(s1)
    if v.verify:
        with pytest.raises(RuntimeError, match="Invalid state checked by the physics model"):
            v.add_state(s2)
    else:
        v.add_state(s2)


def test_other_participant_geometry_pose_and_trace():   # reference: test_other_participant_methods (coordinates, not shapely)
    o1, o2, o3, o4 = Other(id_=0, length=4.0, width=2.0), Other(id_=1, length=3.0), Other(id_=2, width=1.5), Other(id_=3)
    assert np.asarray(o1.geometry).shape == (4, 2) and o2.geometry is not None and o3.geometry is not None and o4.geometry is None
    o1.add_state(State(frame=0, x=5, y=6, heading=0.5))
    assert np.asarray(o1.get_pose(0)).shape == (4, 2)
    o4.add_state(State(frame=0, x=5, y=6, heading=0.5))
    assert tuple(np.asarray(o4.get_pose(0), dtype=float).reshape(-1)[:2]) == (5.0, 6.0)       # no size: the pose is the point


def test_obstacle_and_cyclist_details():                # reference: test_obstacle, test_cyclist_details
    ob = Obstacle(id_=1, length=2.0, width=1.0)
    ob.add_state(State(frame=0, x=0, y=0, heading=0))
    assert ob.is_active(0) and ob.get_state(0).x == 0
    from tactics2d_b200.physics import SingleTrackKinematics

    c = Cyclist(id_=0, verify=True, length=2.0, max_steer=0.5, max_speed=10.0, max_accel=5.0)
    assert isinstance(c.physics_model, SingleTrackKinematics)
    assert c.physics_model.lf == pytest.approx(c.length / 2) and c.physics_model.lr == pytest.approx(c.length / 2)   # cyclist.py:88-94
    t = Trajectory(id_=0, fps=10.0)
    t.add_state(State(frame=0, x=0, y=0, heading=0, vx=1.0, vy=0.0))
    t.add_state(State(frame=100, x=10, y=0, heading=0, vx=1.0, vy=0.0))
    c.bind_trajectory(t)                                   # may be refused by the verification; a trajectory object stays
    assert c.trajectory is not None
    with pytest.raises(TypeError, match="The trajectory must be an instance of Trajectory."):
        c.bind_trajectory("not a trajectory")
    Cyclist(id_=1).load_from_template("non_existent_template")          # warns, does not raise
    assert Cyclist(id_=2, verify=False).physics_model is None
    custom = SingleTrackKinematics(lf=1.0, lr=1.0)
    assert Cyclist(id_=3, verify=True, physics_model=custom).physics_model == custom
    simple = Cyclist(id_=5, length=2.0, width=0.5, verify=False)
    simple.add_state(State(frame=0, x=0, y=0, heading=0))
    simple.add_state(State(frame=100, x=10, y=0, heading=0))
    # reference quirk kept: __init__ loads the template with overwrite=True (cyclist.py:79,110), so the constructor's
    # length / width give way to the template's 1.8 x 0.65
    np.testing.assert_allclose(simple.get_pose(0), [[0.9, -0.325], [0.9, 0.325], [-0.9, 0.325], [-0.9, -0.325]])
