"""Batched step-env state holders for the classic_control reachers.

The dynamics themselves run inside the fused CUDA kernel (fancy_gym_b200/csrc/fg_rollout.cuh);
these classes own the per-env device state (joint angles / velocities / step counters / task
context), the spaces, the constructor kwargs of the reference envs and the reset-time context
sampling:
  HoleReacherEnv      fancy_gym/envs/classic_control/hole_reacher/hole_reacher.py
  ViaPointReacherEnv  fancy_gym/envs/classic_control/viapoint_reacher/viapoint_reacher.py
  SimpleReacherEnv    fancy_gym/envs/classic_control/simple_reacher/simple_reacher.py
  (base: base_reacher/base_reacher.py, base_reacher_direct.py, base_reacher_torque.py)

reset(seed=s) seeds env i with s + i (gymnasium's vector-env convention).  With
context_sampler='numpy' (default) every env draws its context from
Generator(PCG64(SeedSequence(seed_i))) in the reference's draw order, so env i is the reference env
reset with seed s + i.  context_sampler='device' draws the same distributions with torch's Philox
generator on the GPU (for large batches; not stream-compatible with numpy).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional, Union

import numpy as np
import torch

from ... import _lib
from ...utils.gym_compat import Box, Env


def _np_rng(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class BaseReacherEnv(Env):
    env_kind = -1
    torque = False
    n_ctx_obs = 0            # task-specific obs entries between velocity and step counter

    def __init__(self, n_links: int, random_start: bool = True, allow_self_collision: bool = False,
                 num_envs: int = 1, device: Union[str, torch.device, None] = None, context_sampler: str = "numpy",
                 render_mode: Optional[str] = None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        if not 1 <= n_links <= _lib.FG_MAX_DOF:
            raise ValueError(f"n_links must be in 1..{_lib.FG_MAX_DOF}")
        if context_sampler not in ("numpy", "device"):
            raise ValueError("context_sampler must be 'numpy' or 'device'")
        self.n_links = int(n_links)
        self.num_envs = int(num_envs)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.random_start = random_start
        self.allow_self_collision = allow_self_collision
        self.context_sampler = context_sampler
        self.render_mode = render_mode
        self._dt = 0.01                                   # base_reacher.py:21
        self._start_pos = np.hstack([[np.pi / 2], np.zeros(self.n_links - 1)])   # base_reacher.py:34
        bound = 1000.0 if self.torque else 2 * np.pi      # base_reacher_torque.py:16 / base_reacher_direct.py:16
        ab = np.ones(self.n_links) * bound
        self.action_space = Box(low=-ab, high=ab, shape=ab.shape, batch=self._batch_or_none())
        self.observation_space = self._make_observation_space()
        B, n = self.num_envs, self.n_links
        dev = self.device
        self.q = torch.zeros(B, n, dtype=torch.float64, device=dev)
        self.v = torch.zeros(B, n, dtype=torch.float64, device=dev)
        self.steps = torch.zeros(B, dtype=torch.int32, device=dev)
        self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
        self.ctx = torch.zeros(B, 4, dtype=torch.float64, device=dev)
        self._seed_rngs = None
        self._torch_gen = None
        self._was_reset = False

    def _batch_or_none(self):
        return self.num_envs if self.num_envs > 1 else None

    # ---- spaces --------------------------------------------------------------------------------
    def _state_bound(self):
        n = self.n_links
        return np.hstack([[np.pi] * n, [np.pi] * n, [np.inf] * n, [np.inf] * self.n_ctx_obs, [np.inf]])

    def _make_observation_space(self):
        sb = self._state_bound()
        return Box(low=-sb, high=sb, shape=sb.shape, batch=self._batch_or_none())

    # ---- properties the MP wrappers expose (raw_interface_wrapper.py:24-53) ----------------------
    @property
    def dt(self):
        return self._dt

    @property
    def current_pos(self):
        return self.q.clone()

    @property
    def current_vel(self):
        return self.v.clone()

    # ---- reset ---------------------------------------------------------------------------------
    def _seeds(self, seed):
        B = self.num_envs
        if seed is None:
            if self._seed_rngs is None:
                ss = np.random.SeedSequence()
                self._seed_rngs = [np.random.Generator(np.random.PCG64(s)) for s in ss.spawn(B)]
            return None
        seeds = np.asarray(seed).reshape(-1)
        if seeds.size == 1:
            seeds = int(seeds[0]) + np.arange(B)
        if seeds.size != B:
            raise ValueError(f"need one seed or {B} seeds")
        self._seed_rngs = None
        return [int(s) for s in seeds]

    def _sample_numpy(self, seeds) -> Dict[str, np.ndarray]:
        raise NotImplementedError

    def _sample_device(self, gen) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def reset(self, *, seed=None, options: Optional[Dict[str, Any]] = None):
        """-> (obs [B, O] float32 tensor on the device, {}).  options['contexts'] (dict of arrays)
        bypasses sampling; options['random_start'] as in base_reacher.py:77-80."""
        options = options or {}
        B, n, dev = self.num_envs, self.n_links, self.device
        random_start = options.get("random_start", self.random_start)
        if "contexts" in options:
            c = {k: torch.as_tensor(np.asarray(v), dtype=torch.float64) for k, v in options["contexts"].items()}
        elif self.context_sampler == "numpy":
            seeds = self._seeds(seed)
            if seeds is None:
                rngs = self._seed_rngs
            else:
                rngs = [_np_rng(s) for s in seeds]
                self._seed_rngs = rngs
            c = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in self._sample_numpy(rngs, seeds, random_start).items()}
        else:
            if seed is not None or self._torch_gen is None:
                self._torch_gen = torch.Generator(device=dev)
                self._torch_gen.manual_seed(int(np.asarray(seed).reshape(-1)[0]) if seed is not None
                                            else int(np.random.SeedSequence().entropy % (2 ** 62)))
            c = self._sample_device(self._torch_gen, random_start)
        q0 = c.pop("q0").to(dev)
        self.q.zero_()
        self.q[:, 0] = q0
        self.v.zero_()
        self.steps.zero_()
        self.done.zero_()
        self._set_ctx({k: v.to(dev) for k, v in c.items()})
        self._was_reset = True
        return self.get_obs(), {}

    def _set_ctx(self, c):
        raise NotImplementedError

    def _first_joint(self, rng, random_start):
        # base_reacher.py:81-86 (random start angle of the first joint, the arm is straight)
        return rng.uniform(np.pi / 4, 3 * np.pi / 4) if random_start else self._start_pos[0]

    # ---- observation of the current state (reset-time; per-step observations come from the kernel)
    def end_effector(self):
        th = torch.cumsum(self.q, dim=1)
        return torch.stack([torch.cos(th).sum(1), torch.sin(th).sum(1)], dim=1)

    def _task_obs(self, ee):
        raise NotImplementedError

    def get_obs(self):
        ee = self.end_effector()
        parts = [torch.cos(self.q), torch.sin(self.q), self.v, *self._task_obs(ee), self.steps.to(torch.float64)[:, None]]
        return torch.cat(parts, dim=1).to(torch.float32)

    def close(self):
        pass


class HoleReacherEnv(BaseReacherEnv):
    env_kind = _lib.ENV_HOLE_REACHER
    n_ctx_obs = 3    # hole width, ee - goal (2)

    def __init__(self, n_links: int, hole_x: Union[None, float] = None, hole_depth: Union[None, float] = None,
                 hole_width: float = 1., random_start: bool = False, allow_self_collision: bool = False,
                 allow_wall_collision: bool = False, collision_penalty: float = 1000, rew_fct: str = "simple", **kwargs):
        if rew_fct not in ("simple", "vel_acc", "unbounded"):
            raise ValueError("Unknown reward function {}".format(rew_fct))      # hole_reacher.py:57-58
        self.initial_x, self.initial_width, self.initial_depth = hole_x, hole_width, hole_depth
        self.allow_wall_collision = allow_wall_collision
        self.collision_penalty = collision_penalty
        self.rew_fct = rew_fct
        self.rew_fct_code = ("simple", "vel_acc", "unbounded").index(rew_fct)     # fg_config.rew_fct
        super().__init__(n_links, random_start, allow_self_collision, **kwargs)

    def _sample_numpy(self, rngs, seeds, random_start):
        # hole_reacher.py:79-112 (_generate_hole) then base_reacher.py:73-93 on the same stream
        B = self.num_envs
        out = {k: np.zeros(B) for k in ("x", "width", "depth", "q0")}
        for i, rng in enumerate(rngs):
            width = rng.uniform(0.15, 0.5) if self.initial_width is None else float(self.initial_width)
            if self.initial_x is None:
                direction = rng.choice([-1, 1])
                x = direction * rng.uniform(width / 2, 3.5)
            else:
                x = float(self.initial_x)
            depth = rng.uniform(1, 1) if self.initial_depth is None else float(self.initial_depth)
            out["x"][i], out["width"][i], out["depth"][i] = x, width, depth
            out["q0"][i] = self._first_joint(rng, random_start)
        return out

    def _sample_device(self, gen, random_start):
        B, dev = self.num_envs, self.device
        u = torch.rand(B, 4, generator=gen, device=dev, dtype=torch.float64)
        width = 0.15 + 0.35 * u[:, 0] if self.initial_width is None else torch.full((B,), float(self.initial_width), device=dev, dtype=torch.float64)
        if self.initial_x is None:
            direction = torch.where(u[:, 1] < 0.5, -1.0, 1.0)
            x = direction * (width / 2 + (3.5 - width / 2) * u[:, 2])
        else:
            x = torch.full((B,), float(self.initial_x), device=dev, dtype=torch.float64)
        depth = torch.full((B,), 1.0 if self.initial_depth is None else float(self.initial_depth), device=dev, dtype=torch.float64)
        q0 = np.pi / 4 + (np.pi / 2) * u[:, 3] if random_start else torch.full((B,), self._start_pos[0], device=dev, dtype=torch.float64)
        return dict(x=x, width=width, depth=depth, q0=q0)

    def _set_ctx(self, c):
        self.ctx.zero_()
        self.ctx[:, 0], self.ctx[:, 1], self.ctx[:, 2] = c["x"], c["width"], c["depth"]

    def _task_obs(self, ee):
        goal = torch.stack([self.ctx[:, 0], -self.ctx[:, 2]], dim=1)      # hole_reacher.py:100
        return [self.ctx[:, 1:2], ee - goal]


class ViaPointReacherEnv(BaseReacherEnv):
    env_kind = _lib.ENV_VIAPOINT_REACHER
    n_ctx_obs = 4    # ee - via (2), ee - goal (2)

    def __init__(self, n_links, random_start: bool = False, via_target: Union[None, Iterable] = None,
                 target: Union[None, Iterable] = None, allow_self_collision=False, collision_penalty=1000, **kwargs):
        self.intitial_target = target
        self.initial_via_target = via_target
        self.collision_penalty = collision_penalty
        super().__init__(n_links, random_start, allow_self_collision, **kwargs)

    def _goal_draws(self, rng):
        total = float(self.n_links)
        if self.initial_via_target is None:      # viapoint_reacher.py:59-64
            via = np.array([total, total])
            while np.linalg.norm(via) >= 0.5 * total:
                via = rng.uniform(low=-0.5 * total, high=0.5 * total, size=2)
        else:
            via = np.array(self.initial_via_target, dtype=np.float64)
        if self.intitial_target is None:         # viapoint_reacher.py:66-72
            goal = np.array([total, total])
            while np.linalg.norm(goal) >= total or np.linalg.norm(goal) <= 0.5 * total:
                goal = rng.uniform(low=-total, high=total, size=2)
        else:
            goal = np.array(self.intitial_target, dtype=np.float64)
        return via, goal

    def _sample_numpy(self, rngs, seeds, random_start):
        # viapoint_reacher.py:45-53: seeded reset (draws the start angle) -> _generate_goal on the same stream
        # -> seeded reset again: start angle = first variate, goal from the variates after it (App. A.6-Q4)
        B = self.num_envs
        out = dict(via=np.zeros((B, 2)), goal=np.zeros((B, 2)), q0=np.zeros(B))
        for i, rng in enumerate(rngs):
            if seeds is None:                          # unseeded: goal, start, goal, start on one running stream
                self._goal_draws(rng)
                self._first_joint(rng, random_start)
                out["via"][i], out["goal"][i] = self._goal_draws(rng)
                out["q0"][i] = self._first_joint(rng, random_start)
                continue
            q0 = self._first_joint(rng, random_start)
            out["via"][i], out["goal"][i] = self._goal_draws(rng)
            out["q0"][i] = q0
            rngs[i] = _np_rng(seeds[i])                # the second seeded reset restarts the stream
            self._first_joint(rngs[i], random_start)
        return out

    def _sample_device(self, gen, random_start):
        B, dev, total = self.num_envs, self.device, float(self.n_links)

        def ring(lo, hi, half):
            out = torch.empty(B, 2, device=dev, dtype=torch.float64)
            todo = torch.ones(B, dtype=torch.bool, device=dev)
            while bool(todo.any()):
                cand = (torch.rand(B, 2, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * half
                nrm = cand.norm(dim=1)
                ok = todo & (nrm < hi) & (nrm > lo)
                out[ok] = cand[ok]
                todo &= ~ok
            return out
        via = ring(-1.0, 0.5 * total, 0.5 * total) if self.initial_via_target is None else \
            torch.as_tensor(np.asarray(self.initial_via_target, dtype=np.float64), device=dev).expand(B, 2)
        goal = ring(0.5 * total, total, total) if self.intitial_target is None else \
            torch.as_tensor(np.asarray(self.intitial_target, dtype=np.float64), device=dev).expand(B, 2)
        q0 = np.pi / 4 + (np.pi / 2) * torch.rand(B, generator=gen, device=dev, dtype=torch.float64) if random_start \
            else torch.full((B,), self._start_pos[0], device=dev, dtype=torch.float64)
        return dict(via=via, goal=goal, q0=q0)

    def _set_ctx(self, c):
        self.ctx[:, 0:2] = c["via"]
        self.ctx[:, 2:4] = c["goal"]

    def _task_obs(self, ee):
        return [ee - self.ctx[:, 0:2], ee - self.ctx[:, 2:4]]


class SimpleReacherEnv(BaseReacherEnv):
    env_kind = _lib.ENV_SIMPLE_REACHER
    torque = True
    n_ctx_obs = 2    # ee - goal

    def __init__(self, n_links: int, target: Union[None, Iterable] = None, random_start: bool = True,
                 allow_self_collision: bool = False, **kwargs):
        self.inital_target = target
        super().__init__(n_links, random_start, allow_self_collision, **kwargs)
        self._start_pos = np.zeros(self.n_links)       # simple_reacher.py:29

    def _goal_draw(self, rng):
        if self.inital_target is None:                  # simple_reacher.py:87-94
            total = float(self.n_links)
            goal = np.array([total, total])
            while np.linalg.norm(goal) >= total:
                goal = rng.uniform(low=-total, high=total, size=2)
            return goal
        return np.array(self.inital_target, dtype=np.float64)

    def _sample_numpy(self, rngs, seeds, random_start):
        B = self.num_envs
        out = dict(goal=np.zeros((B, 2)), q0=np.zeros(B))
        for i, rng in enumerate(rngs):
            if seeds is None:
                self._goal_draw(rng)
                self._first_joint(rng, random_start)
                out["goal"][i] = self._goal_draw(rng)
                out["q0"][i] = self._first_joint(rng, random_start)
                continue
            q0 = self._first_joint(rng, random_start)
            out["goal"][i] = self._goal_draw(rng)
            out["q0"][i] = q0
            rngs[i] = _np_rng(seeds[i])
            self._first_joint(rngs[i], random_start)
        return out

    def _sample_device(self, gen, random_start):
        B, dev, total = self.num_envs, self.device, float(self.n_links)
        if self.inital_target is None:
            goal = torch.empty(B, 2, device=dev, dtype=torch.float64)
            todo = torch.ones(B, dtype=torch.bool, device=dev)
            while bool(todo.any()):
                cand = (torch.rand(B, 2, generator=gen, device=dev, dtype=torch.float64) * 2 - 1) * total
                ok = todo & (cand.norm(dim=1) < total)
                goal[ok] = cand[ok]
                todo &= ~ok
        else:
            goal = torch.as_tensor(np.asarray(self.inital_target, dtype=np.float64), device=dev).expand(B, 2)
        q0 = np.pi / 4 + (np.pi / 2) * torch.rand(B, generator=gen, device=dev, dtype=torch.float64) if random_start \
            else torch.zeros(B, device=dev, dtype=torch.float64)
        return dict(goal=goal, q0=q0)

    def _set_ctx(self, c):
        self.ctx.zero_()
        self.ctx[:, 0:2] = c["goal"]

    def _task_obs(self, ee):
        return [ee - self.ctx[:, 0:2]]
