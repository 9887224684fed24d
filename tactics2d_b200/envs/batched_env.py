"""``BatchedTrafficEnv`` - Gym-style ``step`` / ``reset`` over N scenarios x M participants.

Keeps the contract of the reference's single-ego envs (``tactics2d/envs/parking.py:219-298``,
``racing.py:145-202``):

* ``reset(seed, options) -> (observation, info)``; ``step(action) -> (observation, reward, terminated,
  truncated, info)``;
* the action order is ``[steering, accel]`` (parking.py:239); out-of-range actions are clipped by the physics
  model, not rejected (single_track_kinematics.py:192-193);
* ``terminated`` iff the scenario status is COMPLETED; ``truncated`` iff the scenario or the ego's traffic
  status is not NORMAL (parking.py:243-248); without a ``target`` a scenario never completes, so every ``done``
  is a truncation;
* the status priority time-exceed -> out-of-bound -> collision (parking.py:361-392);
* the reward chain of ``ParkingEnv._get_reward`` (parking.py:148-190), in its order: -5 collision, -1 time exceeded / no
  action, -5 out of bound, +5 completed, else the time penalty ``-tanh(t / max_step) * 0.001`` + the IoU gain over the
  episode's best + 0.1 x the progress towards the target centre (``_max_iou`` / ``_min_dist_to_target`` are kept per
  scenario on the device).  Two deliberate differences, both where the reference misbehaves: its check_status stores
  NO_ACTION into the *traffic* status (parking.py:372), here it is a scenario status and earns the -1 the reward chain
  intends; its first step adds ``(inf - d) * 0.1`` to the reward, here the first step only records the distance.

What differs, deliberately: the environment is *vectorised* (every quantity has a leading N axis and lives on
the GPU), all M participants are simulated (the ego is participant 0; the others take ``npc_action`` or zeros),
the observation is the state tensors themselves (the reference renders a BEV image, which is outside this
path), and scenarios are drawn from a pool of initial states instead of the reference's map generators.
The reference envs construct a ``render_manager`` that is commented out at this commit and crash on the
first ``update``; this class follows their documented contract, not the crash.
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from ..traffic import BatchedScenarioManager, ScenarioStatus, TrafficStatus
from ..world import BatchedWorld


class InvalidAction(Exception):
    """Raised when an action does not have the batched action shape (parking.py:235-236 raises it for
    actions outside the action space)."""


class BatchedTrafficEnv:
    metadata = {"render_modes": []}

    def __init__(self, scene, device="cuda:0", max_step: int = 1000, step_size: int = 100, delta_t: int = 5,
                 any_participant: bool = False, auto_reset: bool = True, target=None, arrival_threshold: float = 0.95,
                 no_action_max_step: int = 100):
        """``scene``: a :class:`tactics2d_b200.synthetic.Scene` (initial states, types, map tile, bounds);
        ``target``: optional [N, 5] target areas (cx, cy, heading, half_len, half_wid) for the egos - enables the
        ``Arrival`` (-> COMPLETED / ``terminated``) and ``NoAction`` detectors and the IoU reward terms of
        ``ParkingEnv._get_reward`` (parking.py:148-190)."""
        import torch

        self.scene = scene
        n, m = scene.shape
        self.num_envs, self.num_participants = n, m
        self.max_step = int(max_step)
        self.auto_reset = auto_reset
        self.world = BatchedWorld(n, m, scene.table, device=device, interval=step_size, delta_t=delta_t, max_step=max_step,
                                  any_participant=any_participant, steer_first=True)
        self.world.set_map(scene.segments, scene.bounds)
        self.scenario_manager = BatchedScenarioManager(self.world, max_step=max_step, step_size=step_size)
        dev = self.world.device
        self._pool = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.state().items()}
        self._type_id = torch.from_numpy(scene.type_id).to(dev)
        self.scenario_manager.set_initial_state(self._pool)
        self._action = torch.zeros((n, m, 2), dtype=torch.float32, device=dev)
        self._rng = np.random.default_rng(0)
        if target is not None:
            self.world.set_goal(target, arrival_threshold, no_action_max_step)
        self.observation_space = {"shape": (n, m, 6), "dtype": "float32"}
        self.action_space = {"shape": (n, 2), "low": (-np.inf, -np.inf), "high": (np.inf, np.inf)}

    # ------------------------------------------------------------------ helpers
    def _obs(self):
        return self.scenario_manager.get_observation()

    def _info(self, status, traffic, flags, hit_index, hit_segment):
        return {"scenario_status": status, "traffic_status": traffic, "flags": flags, "hit_index": hit_index,
                "hit_segment": hit_segment, "step_count": self.world.step_count}

    # ------------------------------------------------------------------ gym surface
    def reset(self, seed: int = None, options: dict = None):
        import torch

        if seed is not None:
            self._rng = np.random.default_rng(seed)
        perm = None
        if options and options.get("shuffle"):
            perm = torch.from_numpy(self._rng.permutation(self.num_envs).astype(np.int32)).to(self.world.device)
        self.world.type_id.copy_(self._type_id)
        self.scenario_manager.reset(pool_index=perm)
        self.world.reset_env_trackers()
        status = torch.full((self.num_envs,), int(ScenarioStatus.NORMAL), dtype=torch.uint8, device=self.world.device)
        traffic = torch.full((self.num_envs, self.num_participants), int(TrafficStatus.NORMAL), dtype=torch.uint8,
                             device=self.world.device)
        o = self.world._out
        return self._obs(), self._info(status, traffic, torch.zeros_like(o.flags), torch.full_like(o.hit_index, -1),
                                       torch.full_like(o.hit_segment, -1))

    def step(self, action, npc_action=None):
        """``action``: fp32 device tensor [N, 2] = (steering, accel) of the ego (participant 0), or [N, M, 2] for
        all participants; ``npc_action`` [N, M-1, 2] optionally drives the others (otherwise they keep their rows of the
        internal action array: zeros, or what ``world.set_controllers`` computes on the device every tick).

        Launches per call: the controllers (if set), the fused tick, the env epilogue (reward / terminated / truncated /
        TrafficStatus / done in one kernel) and the masked reset - no elementwise PyTorch.  The tensors in the returned
        tuple and in ``info`` are views of buffers owned by the env: they hold this step's values until the next ``step``."""
        w = self.world
        if action.dim() == 3:
            if tuple(action.shape) != (self.num_envs, self.num_participants, 2):
                raise InvalidAction(f"Action of shape {tuple(action.shape)} is not in the action space.")
            full = action.contiguous()
            w.set_ego_action(None)
        else:
            if tuple(action.shape) != (self.num_envs, 2):
                raise InvalidAction(f"Action of shape {tuple(action.shape)} is not in the action space.")
            full = self._action
            w.set_ego_action(action.to(device=w.device, dtype=full.dtype).contiguous())   # read by the kernels, not scattered
            if npc_action is not None:
                full[:, 1:, :] = npc_action
        if w.last_accel is not None:
            w.control(full)
        self.scenario_manager.update(full)
        status, traffic = self.scenario_manager.check_status()
        e = self.scenario_manager.env_result
        r = w.result
        info = self._info(status, traffic, r.flags, r.hit_index, r.hit_segment)
        if r.iou is not None:
            info["iou"] = r.iou
        if self.auto_reset:
            self.scenario_manager.reset(mask=e.done)
        return self._obs(), e.reward, e.terminated, e.truncated, info

    def render(self):
        raise NotImplementedError("rendering is outside this hot path")

    def close(self):
        self.world.close()
