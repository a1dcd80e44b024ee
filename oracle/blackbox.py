"""Restatement of the reference's black-box episode loop, batched over environments.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
fancy_gym/black_box/black_box_wrapper.py:96-120 (get_trajectory), :150-217 (step), :222-229
(reset) and the controllers fancy_gym/black_box/controller/{pd,vel,pos}_controller.py on top of
oracle/reacher.py (env half, pinned) and oracle/mp.py (MP half, unpinned).

The TimeLimit wrapper that gymnasium puts around every registered env (max_episode_steps=200,
App. A.6-Q12) is folded in: `truncated` becomes True on the max_episode_steps-th env step.
"""
from __future__ import annotations

import copy

import numpy as np

from . import mp as omp
from .reacher import BatchedReacher

# ---- resolved configs of the twelve classic_control black-box ids: registry defaults (fancy_gym/envs/registry.py:62-129)
# ---- merged with each env's mp_wrapper.mp_config through nested_update, incl. the `_type` replace quirk (:264-277) ----
_MP_DEFAULTS = {
    "ProMP": dict(traj=dict(trajectory_generator_type="promp"), phase=dict(phase_generator_type="linear"),
                  basis=dict(basis_generator_type="zero_rbf", num_basis=5, num_basis_zero_start=1, basis_bandwidth_factor=3.0)),
    "DMP": dict(traj=dict(trajectory_generator_type="dmp"), phase=dict(phase_generator_type="exp"),
                basis=dict(basis_generator_type="rbf", num_basis=5)),
    "ProDMP": dict(traj=dict(trajectory_generator_type="prodmp", duration=2.0, weights_scale=1.0),
                   phase=dict(phase_generator_type="exp", tau=1.5),
                   basis=dict(basis_generator_type="prodmp", alpha=10, num_basis=5)),
}
_DEFAULT_CTRL = dict(controller_type="motor", p_gains=1.0, d_gains=0.1)

# envs/__init__.py:38-87 (registration kwargs) and */mp_wrapper.py (mp_config)
_ENVS = {
    "HoleReacher-v0": dict(
        env=dict(kind="hole", n_links=5, random_start=True, allow_self_collision=False, allow_wall_collision=False,
                 hole_width=None, hole_depth=1, hole_x=None, collision_penalty=100),
        mp_config={"ProMP": dict(ctrl=dict(controller_type="velocity"), traj=dict(weights_scale=2)),
                   "DMP": dict(ctrl=dict(controller_type="velocity"), traj=dict(weights_scale=500), phase=dict(alpha_phase=2.5)),
                   "ProDMP": {}}),
    "ViaPointReacher-v0": dict(
        env=dict(kind="viapoint", n_links=5, allow_self_collision=False, collision_penalty=1000),
        mp_config={"ProMP": dict(ctrl=dict(controller_type="velocity")),
                   "DMP": dict(ctrl=dict(controller_type="velocity"), traj=dict(weights_scale=50), phase=dict(alpha_phase=2)),
                   "ProDMP": {}}),
    "SimpleReacher-v0": dict(
        env=dict(kind="simple", n_links=2),
        mp_config={"ProMP": dict(ctrl=dict(p_gains=0.6, d_gains=0.075)),
                   "DMP": dict(ctrl=dict(p_gains=0.6, d_gains=0.075), traj=dict(weights_scale=50), phase=dict(alpha_phase=2)),
                   "ProDMP": {}}),
}
_ENVS["LongSimpleReacher-v0"] = dict(env=dict(kind="simple", n_links=5), mp_config=_ENVS["SimpleReacher-v0"]["mp_config"])


def _merge(base, update):
    if any(k.endswith("_type") for k in update):
        return dict(update)
    out = dict(base)
    out.update(update)
    return out


def _resolve(name, mp_type):
    d, e = _MP_DEFAULTS[mp_type], _ENVS[name]
    own = e["mp_config"][mp_type]
    return dict(env=dict(e["env"]), traj=_merge(d["traj"], own.get("traj", {})), phase=_merge(d["phase"], own.get("phase", {})),
                basis=_merge(d["basis"], own.get("basis", {})), ctrl=_merge(_DEFAULT_CTRL, own.get("ctrl", {})))


RESOLVED = {f"fancy_{mp}/{name}": _resolve(name, mp) for name in _ENVS for mp in _MP_DEFAULTS}


def controller_action(ctrl, des_pos, des_vel, c_pos, c_vel):
    """pd_controller.py:21-29, vel_controller.py:8-9, pos_controller.py:8-9."""
    t = ctrl["controller_type"].lower()
    if t == "velocity":
        return des_vel
    if t == "position":
        return des_pos
    if t == "motor":
        if des_pos.shape != c_pos.shape or des_vel.shape != c_vel.shape:
            raise ValueError("Mismatch in dimension between desired and current state")
        p, d = ctrl.get("p_gains", 1), ctrl.get("d_gains", 0.5)
        return p * (des_pos - c_pos) + d * (des_vel - c_vel)
    raise ValueError(f"Specified controller type {t} not supported")


class BlackBoxOracle:
    def __init__(self, env: BatchedReacher, traj_gen: omp.MPBase, ctrl: dict, duration=None,
                 max_episode_steps=200, verbose=1, learn_sub_trajectories=False, replanning_schedule=None,
                 reward_aggregation=np.sum, max_planning_times=np.inf, condition_on_desired=False):
        self.env = env
        self.traj_gen = traj_gen
        self.ctrl = ctrl
        self.dt = env.dt
        self.max_episode_steps = max_episode_steps
        self.duration = duration if duration is not None else max_episode_steps * env.dt
        self.verbose = verbose
        self.learn_sub_trajectories = learn_sub_trajectories
        self.do_replanning = replanning_schedule is not None
        self.replanning_schedule = replanning_schedule or (lambda *x: False)
        self.reward_aggregation = reward_aggregation
        self.max_planning_times = max_planning_times
        self.condition_on_desired = condition_on_desired
        self.return_context_observation = not (learn_sub_trajectories or self.do_replanning)
        self.traj_gen.set_duration(self.duration, self.dt)
        self.bounds = self.traj_gen.get_params_bounds()

    # black_box_wrapper.py:222-229
    def reset(self, seeds=None, contexts=None):
        self.current_traj_steps = 0
        self.plan_steps = 0
        self.traj_gen.reset()
        self.condition_pos = None
        self.condition_vel = None
        obs = self.env.reset(seeds=seeds, contexts=contexts)
        self.elapsed = np.zeros(self.env.B, dtype=np.int64)     # TimeLimit._elapsed_steps
        self.done = np.zeros(self.env.B, bool)
        self._last_obs = obs.copy()
        self._last_info = {}
        return self.observation(obs)

    # black_box_wrapper.py:89-94 (+ utils/wrappers.py:49-63 TimeAwareObservation when replanning)
    def observation(self, obs):
        if self.return_context_observation:
            return obs[:, self.env.context_mask()]
        t = (self.elapsed / self.max_episode_steps)[:, None]
        return np.concatenate([obs, t], axis=1).astype(obs.dtype)

    # black_box_wrapper.py:96-120
    def get_trajectory(self, action):
        duration = self.duration
        if self.learn_sub_trajectories:
            duration = None
            self.traj_gen.reset()
        clipped = np.clip(action, self.bounds[0], self.bounds[1])
        self.traj_gen.set_params(clipped)
        init_time = np.array(0 if not self.do_replanning else self.current_traj_steps * self.dt)
        # per env: the recorded desired state once THAT env's loop has broken, its current state before (each env of the
        # batch is its own wrapper in the reference: :110-111)
        cpos, cvel = self.env.current_pos, self.env.current_vel
        if self.condition_pos is not None:
            m = self.condition_has[:, None]
            cpos = np.where(m, self.condition_pos.astype(np.float64), cpos)      # (float32 -> float64 is exact)
            cvel = np.where(m, self.condition_vel.astype(np.float64), cvel)
        self.traj_gen.set_initial_conditions(init_time, cpos, cvel)
        self.traj_gen.set_duration(duration, self.dt)
        return self.traj_gen.get_traj_pos(), self.traj_gen.get_traj_vel()

    # black_box_wrapper.py:150-217, for all envs of the batch in lock-step
    def step(self, action):
        env, B = self.env, self.env.B
        action = np.asarray(action)
        if action.ndim == 1:
            action = action[None]
        position, velocity = self.get_trajectory(action)
        if position.ndim == 2:
            position, velocity = position[None], velocity[None]
        position = np.broadcast_to(position, (B, *position.shape[1:]))
        velocity = np.broadcast_to(velocity, (B, *velocity.shape[1:]))
        T = position.shape[1]
        rewards = np.zeros((B, T))
        length = np.zeros(B, dtype=np.int64)
        terminated = np.zeros(B, bool)
        truncated = np.zeros(B, bool)
        min_margin = np.full(B, np.inf)
        last_obs = self._last_obs
        last_info = self._last_info
        step_actions = np.zeros((B, T, env.n_links)) if self.verbose >= 2 else None
        od = env.obs_dim + (0 if self.return_context_observation else 1)
        step_obs = np.zeros((B, T, od), dtype=np.float32) if self.verbose >= 2 else None
        live = ~self.done            # envs that finished their episode in an earlier call stay frozen
        brk = ~live
        self.plan_steps += 1
        new_cond_pos = None if self.condition_pos is None else np.array(self.condition_pos, copy=True)
        new_cond_vel = None if self.condition_vel is None else np.array(self.condition_vel, copy=True)
        for t in range(T):
            run = ~brk
            if not run.any():
                break
            pos, vel = position[:, t], velocity[:, t]
            a = controller_action(self.ctrl, pos, vel, env.current_pos, env.current_vel)
            a = np.clip(a, env.action_low, env.action_high)
            frozen = _snapshot(env) if not run.all() else None
            obs, r, term, info = env.step(a)
            if frozen is not None:
                _restore(env, frozen, ~run)
            self.elapsed[run] += 1
            trunc = self.elapsed >= self.max_episode_steps
            rewards[run, t] = r[run]
            length[run] = t + 1
            min_margin[run] = np.minimum(min_margin[run], info["margin"][run])
            if not last_info:
                last_info = {k: v.copy() for k, v in info.items() if k != "margin"}
            last_obs[run] = obs[run]
            for k in last_info:
                last_info[k][run] = info[k][run]
            if self.verbose >= 2:
                step_actions[run, t] = a[run]
                step_obs[run, t] = (obs if self.return_context_observation else self.observation(obs))[run]
            terminated[run] = term[run]
            truncated[run] = trunc[run]
            replan = bool(self.replanning_schedule(env.current_pos, env.current_vel, obs, a,
                                                   t + 1 + self.current_traj_steps)) \
                and self.plan_steps < self.max_planning_times
            stop = run & (term | trunc | replan)
            if self.condition_on_desired and stop.any():      # only a BREAK records the desired state (:196-201)
                if new_cond_pos is None:
                    new_cond_pos = np.zeros_like(pos)
                    new_cond_vel = np.zeros_like(vel)
                    self.condition_has = np.zeros(B, bool)
                new_cond_pos[stop] = pos[stop]
                new_cond_vel[stop] = vel[stop]
                self.condition_has = self.condition_has | stop
            n_valid = getattr(self.traj_gen, "n_valid", None)
            if n_valid is not None:          # ragged sub-trajectories: an env's plan simply ends after its own number of
                stop = stop | (run & (t + 1 >= n_valid))      # steps (the reference's for loop runs out: no break)
            brk = brk | stop
            self.done = self.done | (run & (term | trunc))
        if self.condition_on_desired:
            self.condition_pos, self.condition_vel = new_cond_pos, new_cond_vel
        # all live envs of a batch advance by the same number of steps unless they terminated
        adv = length[live & ~self.done]
        self.current_traj_steps += int(adv[0]) if adv.size else int(length.max(initial=0))
        self._last_obs, self._last_info = last_obs, last_info
        ret = np.array([self.reward_aggregation(rewards[b, :length[b]]) if length[b] else 0.0 for b in range(B)])
        infos = dict(last_info)
        infos["trajectory_length"] = length
        infos["min_margin"] = min_margin
        if self.verbose >= 2:
            infos.update(positions=position, velocities=velocity, step_actions=step_actions,
                         step_observations=step_obs, step_rewards=rewards)
        return self.observation(last_obs), ret, terminated, truncated, infos


def _snapshot(env):
    return {k: copy.copy(getattr(env, k, None)) for k in ("q", "v", "acc", "steps", "J", "ee_latch")}


def _restore(env, snap, mask):
    for k, old in snap.items():
        cur = getattr(env, k, None)
        if old is None or cur is None:
            continue
        if cur.dtype != old.dtype:
            old = old.astype(cur.dtype)
        cur[mask] = old[mask]


def make_oracle(env_id, mode="mirror", mp_overrides=None, **bb_kwargs):
    """Builds the oracle for one of the twelve classic_control ids (RESOLVED above; `mp_overrides` = {"env": {...},
    "traj": ..., "phase": ..., "basis": ..., "ctrl": ...} updates the sections), mirroring
    make_bb (fancy_gym/utils/make_env_helpers.py:68-136): duration = max_episode_steps * dt,
    tau defaults to the duration, learn_sub_trajectories implies learn_tau, default bounds."""
    cfg = copy.deepcopy(RESOLVED[env_id])
    for k, v in (mp_overrides or {}).items():
        cfg[k] = _merge(cfg[k], v)        # (a section that names a `*_type` replaces the base, registry.py:272-274)
    env = BatchedReacher(**cfg["env"])
    duration = 200 * env.dt
    phase = dict(cfg["phase"])
    if phase.get("tau") is None:
        phase["tau"] = duration
    if bb_kwargs.get("learn_sub_trajectories") is not None:
        phase["learn_tau"] = True
    if phase.get("learn_tau") and phase.get("tau_bound") is None:
        phase["tau_bound"] = [env.dt * 2, duration]
    if phase.get("learn_delay") and phase.get("delay_bound") is None:
        phase["delay_bound"] = [0, duration - env.dt * 2]
    pg = omp.get_phase_generator(mode=mode, **phase)
    bg = omp.get_basis_generator(phase_generator=pg, **cfg["basis"])
    tg = omp.get_trajectory_generator(action_dim=env.n_links, basis_generator=bg, **cfg["traj"])
    return BlackBoxOracle(env, tg, cfg["ctrl"], duration=duration, **bb_kwargs)
