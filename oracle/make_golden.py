"""Generate golden physics vectors from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run here, where ``/root/reference`` exists:

    python oracle/make_golden.py            # writes tests/golden/physics_*.npz

The reference's ``tactics2d.physics`` and ``tactics2d.participant.trajectory`` import
with NumPy alone; nothing else of the reference is importable in
this image (shapely / gymnasium absent).  The vectors are committed so that the GPU
box (which has no ``/root/reference``) can hold both the oracle and the CUDA path to
the reference's own numbers.  Action scripts replayed below are the reference test
suite's ``VEHICLE_ACTION_LIST`` / ``PEDESTRIAN_ACTION_LIST`` (tests/test_physics.py:52-73).
"""

from __future__ import annotations

import os
import sys

import numpy as np

REF = os.environ.get("T2D_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# medium_car, participant_template.py:78-95 ; ranges as Vehicle.load_from_template sets them
MEDIUM = dict(lf=4.284 / 2 - 0.880, lr=4.284 / 2 - 0.767, mass=1620.0, mass_height=1.452 / 2)
RANGES = dict(steer_range=(-0.524, 0.524), speed_range=(-16.67, 69.44), accel_range=(-11.0, 3.121))

VEHICLE_ACTION_LIST = [((0, 0), 1000), ((1, 0), 1000), ((-1, 0), 1000), ((4, 0), 1000),
                       ((-4, 0), 1000), ((15, 0), 2000), ((-15, 0), 500), ((1, 0), 1000),
                       ((0.1, 0.3), 5000), ((0.1, -0.3), 5000), ((0.1, 0.6), 5000),
                       ((0.1, -0.6), 5000)]
PEDESTRIAN_ACTION_LIST = [((0, 0), 100), ((1, 0), 500), ((-1, 0), 500), ((1, 0), 500),
                          ((0, 1), 500), ((0, -1), 500), ((1, 1), 500), ((2, 2), 500),
                          ((-2, -2), 2000), ((-1, 2), 500), ((2, -1), 500)]


def main():
    sys.path.insert(0, REF)
    from tactics2d.participant.trajectory import State
    from tactics2d.physics import PointMass, SingleTrackDynamics, SingleTrackKinematics

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260924)

    def rec_bicycle(model, states, actions, interval):
        out = np.zeros((len(states), 8))
        for i, ((x, y, h, v), (a, d)) in enumerate(zip(states, actions)):
            s, a_c, d_c = model.step(State(0, x=x, y=y, heading=h, speed=v), a, d, interval)
            vx, vy = s.velocity
            out[i] = (s.x, s.y, s.heading, s.speed, vx, vy, a_c, d_c)
        return out

    n = 192
    states = np.stack([rng.uniform(-500, 500, n), rng.uniform(-500, 500, n),
                       rng.uniform(0, 2 * np.pi, n), rng.uniform(-15, 60, n)], 1)
    # a slice of low / zero speeds and out-of-range speeds
    states[:16, 3] = rng.uniform(-0.3, 0.3, 16)
    states[16:20, 3] = 0.0
    states[20:24, 3] = (80.0, -30.0, 69.44, -16.67)
    actions = np.stack([rng.uniform(-14, 6, n), rng.uniform(-0.8, 0.8, n)], 1)
    actions[24:28] = 0.0

    cases = {}
    for name, kw in [("con", RANGES), ("unc", dict())]:
        for interval, delta_t in [(100, 5), (9, 5), (50, 3), (100, None), (33, 10)]:
            m = SingleTrackKinematics(lf=MEDIUM["lf"], lr=MEDIUM["lr"], interval=interval,
                                      delta_t=delta_t, **kw)
            cases[f"kin_{name}_{interval}_{delta_t}"] = rec_bicycle(m, states, actions, interval)
            m = SingleTrackDynamics(lf=MEDIUM["lf"], lr=MEDIUM["lr"], mass=MEDIUM["mass"],
                                    mass_height=MEDIUM["mass_height"], interval=interval,
                                    delta_t=delta_t, **kw)
            cases[f"dyn_{name}_{interval}_{delta_t}"] = rec_bicycle(m, states, actions, interval)
    np.savez(os.path.join(OUT, "physics_bicycle.npz"), states=states, actions=actions,
             lf=MEDIUM["lf"], lr=MEDIUM["lr"], mass=MEDIUM["mass"],
             mass_height=MEDIUM["mass_height"], steer_range=RANGES["steer_range"],
             speed_range=RANGES["speed_range"], accel_range=RANGES["accel_range"], **cases)

    # ---- point mass
    pstates = np.stack([rng.uniform(-50, 50, n), rng.uniform(-50, 50, n),
                        rng.uniform(-5, 5, n), rng.uniform(-5, 5, n)], 1)  # x y vx vy
    pstates[:8, 2:] = 0.0
    pstates[8:16, 2:] *= 2.0  # some above the 7 m/s limit already
    pact = np.stack([rng.uniform(-4, 4, n), rng.uniform(-4, 4, n)], 1)
    pact[:4] = 0.0
    pact[16:20] = 1e-7
    pcases = {}
    for name, sr in [("ped", (-7.0, 7.0)), ("band", (1.0, 3.0)), ("flt", 4.0), ("unc", None)]:
        for backend in ("newton", "euler"):
            for interval, delta_t in [(100, 5), (9, 5), (50, 3)]:
                m = PointMass(speed_range=sr, accel_range=(-1.5, 1.5), interval=interval,
                              delta_t=delta_t, backend=backend)
                out = np.zeros((n, 6))
                for i in range(n):
                    x, y, vx, vy = pstates[i]
                    st = State(0, x=x, y=y, heading=float(np.arctan2(vy, vx)), vx=vx, vy=vy)
                    s = m.step(st, tuple(pact[i]), interval)
                    out[i] = (s.x, s.y, s.heading, s.vx, s.vy, s.speed)
                pcases[f"pm_{name}_{backend}_{interval}_{delta_t}"] = out
    np.savez(os.path.join(OUT, "physics_pointmass.npz"), states=pstates, actions=pact, **pcases)

    # ---- scripted rollouts (tests/test_physics.py:77-110 simulate_actions), free-running
    roll = {}
    for tag, cls, extra in [("kin", SingleTrackKinematics, {}),
                            ("dyn", SingleTrackDynamics, dict(mass=MEDIUM["mass"],
                                                              mass_height=MEDIUM["mass_height"]))]:
        for interval, delta_t in [(100, 5), (50, 3), (9, 5)]:
            m = cls(lf=MEDIUM["lf"], lr=MEDIUM["lr"], interval=interval, delta_t=delta_t,
                    **extra, **RANGES)
            s = State(0, x=10.0, y=10.0, heading=0.3, speed=5.0)
            traj = [(s.x, s.y, s.heading, s.speed)]
            acts = []
            for action, duration in VEHICLE_ACTION_LIST:
                for _ in np.arange(0, duration, interval):
                    s, _, _ = m.step(s, action[0], action[1], interval)
                    traj.append((s.x, s.y, s.heading, s.speed))
                    acts.append(action)
            roll[f"{tag}_{interval}_{delta_t}_traj"] = np.array(traj)
            roll[f"{tag}_{interval}_{delta_t}_act"] = np.array(acts, dtype=np.float64)
    for backend in ("newton", "euler"):
        m = PointMass(speed_range=(-7.0, 7.0), accel_range=(-1.5, 1.5), interval=100,
                      delta_t=5, backend=backend)
        s = State(0, x=0.0, y=0.0, heading=0.0, vx=0.0, vy=0.0)
        traj = [(s.x, s.y, s.heading, s.vx, s.vy)]
        acts = []
        for action, duration in PEDESTRIAN_ACTION_LIST:
            for _ in np.arange(0, duration, 100):
                s = m.step(s, action, 100)
                traj.append((s.x, s.y, s.heading, s.vx, s.vy))
                acts.append(action)
        roll[f"pm_{backend}_traj"] = np.array(traj)
        roll[f"pm_{backend}_act"] = np.array(acts, dtype=np.float64)
    np.savez(os.path.join(OUT, "physics_rollouts.npz"), **roll)

    # ---- verify_state truth table (single_track_kinematics.py:200-250)
    m = SingleTrackKinematics(lf=MEDIUM["lf"], lr=MEDIUM["lr"], **RANGES)
    vs_in, vs_out = [], []
    for _ in range(256):
        last = (rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(0, 2 * np.pi),
                rng.uniform(-5, 30))
        a, d = rng.uniform(-12, 4), rng.uniform(-0.6, 0.6)
        s, _, _ = m.step(State(0, x=last[0], y=last[1], heading=last[2], speed=last[3]), a, d, 100)
        cand = np.array([s.x, s.y, s.heading, s.speed]) + rng.normal(0, 1, 4) * rng.choice(
            [0.0, 1e-3, 0.05, 0.5])
        cand[2] = np.mod(cand[2], 2 * np.pi)
        ok = m.verify_state(State(100, x=cand[0], y=cand[1], heading=cand[2], speed=cand[3]),
                            State(0, x=last[0], y=last[1], heading=last[2], speed=last[3]), 100)
        vs_in.append(np.concatenate([last, cand]))
        vs_out.append(bool(ok))
    np.savez(os.path.join(OUT, "verify_state.npz"), inputs=np.array(vs_in),
             valid=np.array(vs_out))
    print("golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
