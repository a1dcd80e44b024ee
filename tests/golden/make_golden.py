"""Generates tests/golden/*.npz FROM THE REFERENCE'S OWN FILES and pins the oracle to them.

Run here (this container has /root/reference; the GPU box does not):
    python tests/golden/make_golden.py

What it does
  1. env_kat.npz     the reference's classic_control env classes (imported unmodified through
                     oracle/ref_loader.py) are stepped with deterministic fp32 actions
                     a_t[d] = A*sin(0.05 t + d); per-step obs / reward / termination are stored
                     together with the sampled contexts (reset(seed)).
  2. bb_*.npz        the reference's own BlackBoxWrapper.step loop + controllers + TimeLimit
                     semantics + TimeAwareObservation, driven by the oracle's 'shipped' MP
                     restatement (mp_pytorch itself is absent: see oracle/mp.py header), for the
                     three BASELINE envs incl. replanning / condition_on_desired.
  3. asserts that oracle/reacher.py and oracle/blackbox.py reproduce every stored array
     (bit-exact for obs/flags/lengths, <= 8e-16 relative for float64 rewards: numpy's 1-D
     `linalg.norm` and the batched sqrt(dx^2+dy^2) may differ by one ulp, i.e. up to three in the square).
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mp as omp  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402
from oracle.blackbox import RESOLVED, make_oracle  # noqa: E402
from oracle.reacher import BatchedReacher  # noqa: E402

_HOLE = dict(n_links=5, random_start=True, hole_width=None, hole_depth=1, hole_x=None, collision_penalty=100)
ENV_CASES = [
    # key, reference env name, oracle kind, oracle kwargs, seeds, amplitudes, constructor overrides given to BOTH sides
    ("HoleReacher-v0", "HoleReacher-v0", "hole", _HOLE, [0, 1, 2, 3, 4, 5, 6, 7], [0.3, 1.5, 6.0, 3.0, 0.1, 2.0, 4.0, 1.0], {}),
    ("ViaPointReacher-v0", "ViaPointReacher-v0", "viapoint", dict(n_links=5, collision_penalty=1000), [0, 1, 2, 3],
     [0.3, 6.0, 1.5, 3.0], {}),
    ("SimpleReacher-v0", "SimpleReacher-v0", "simple", dict(n_links=2), [0, 1, 2, 3], [5.0, 20.0, 100.0, 1.0], {}),
    ("LongSimpleReacher-v0", "LongSimpleReacher-v0", "simple", dict(n_links=5), [0, 1], [5.0, 50.0], {}),
    # the other two HoleReacher reward functions (hole_reacher.py:48-58)
    ("HoleReacher-v0+vel_acc", "HoleReacher-v0", "hole", _HOLE, [0, 1, 2, 3, 4, 5], [0.3, 1.5, 6.0, 0.1, 2.0, 1.0], dict(rew_fct="vel_acc")),
    ("HoleReacher-v0+unbounded", "HoleReacher-v0", "hole", _HOLE, [0, 1, 2, 3, 4, 5], [0.3, 1.5, 6.0, 0.1, 2.0, 1.0], dict(rew_fct="unbounded")),
    # constructor options: collisions allowed, fixed hole, fixed start
    ("HoleReacher-v0+free", "HoleReacher-v0", "hole", _HOLE, [0, 1, 2, 3], [6.0, 3.0, 4.0, 2.0],
     dict(allow_self_collision=True, allow_wall_collision=True)),
    ("HoleReacher-v0+fixed", "HoleReacher-v0", "hole", _HOLE, [0, 1, 2], [0.3, 1.5, 3.0],
     dict(hole_x=1.5, hole_width=0.3, hole_depth=0.8, random_start=False)),
    ("ViaPointReacher-v0+fixed", "ViaPointReacher-v0", "viapoint", dict(n_links=5, collision_penalty=1000), [0, 1], [0.3, 6.0],
     dict(via_target=(1.0, 1.0), target=(3.0, 2.0), random_start=True)),
    ("ViaPointReacher-v0+free", "ViaPointReacher-v0", "viapoint", dict(n_links=5, collision_penalty=1000), [0, 1], [6.0, 3.0],
     dict(allow_self_collision=True)),
    ("SimpleReacher-v0+fixed", "SimpleReacher-v0", "simple", dict(n_links=2), [0, 1], [5.0, 20.0],
     dict(target=(0.5, 1.0), random_start=False)),
]


def close64(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        # one ulp in a distance becomes up to three ulps (6.7e-16 relative) in its square, the reward's distance term
        ok = same_inf | (np.abs(a - b) <= 8e-16 * np.maximum(np.abs(a), np.abs(b)))
    return bool(np.all(ok))


def gen_env_kat(ns):
    out = {}
    for key, name, kind, kw, seeds, amps, over in ENV_CASES:
        kw = {**kw, **over}
        n = kw["n_links"]
        T = 200
        obs_dim = None
        rec = dict(obs0=[], obs=[], rew=[], term=[], length=[])
        ctx_rec = []
        for s, A in zip(seeds, amps):
            env = rl.make_step_env(ns, name, **over)
            ob0, _ = env.reset(seed=s)
            obs_dim = ob0.shape[0]
            obs = np.zeros((T, obs_dim), np.float32)
            rew = np.zeros(T)
            term = np.zeros(T, bool)
            L = T
            for t in range(T):
                a = (A * np.sin(0.05 * t + np.arange(n))).astype(np.float32)
                ob, r, te, tr, info = env.step(a)
                obs[t], rew[t], term[t] = ob, r, te
                if te:
                    L = t + 1
                    break
            rec["obs0"].append(ob0); rec["obs"].append(obs); rec["rew"].append(rew)
            rec["term"].append(term); rec["length"].append(L)
            if kind == "hole":
                ctx_rec.append([env._tmp_x, env._tmp_width, float(env._tmp_depth), env._start_pos[0]])
            elif kind == "viapoint":
                ctx_rec.append([*env._via_point, *env._goal, env._start_pos[0]])
            else:
                ctx_rec.append([*env._goal, env._start_pos[0]])
        name = key
        key = key.replace("-", "_")
        out[f"{key}/seeds"] = np.array(seeds)
        out[f"{key}/amps"] = np.array(amps)
        out[f"{key}/ctx"] = np.array(ctx_rec, dtype=np.float64)
        for k, v in rec.items():
            out[f"{key}/{k}"] = np.stack(v) if k != "length" else np.array(v)

        # ---- pin the oracle ----
        o = BatchedReacher(kind, **kw)
        ob0 = o.reset(seeds=seeds)
        assert np.array_equal(ob0, out[f"{key}/obs0"]), name
        for t in range(T):
            a = np.stack([(A * np.sin(0.05 * t + np.arange(n))).astype(np.float32) for A in amps])
            ob, r, te, info = o.step(a)
            for i in range(len(seeds)):
                if t < out[f"{key}/length"][i]:
                    assert np.array_equal(ob[i], out[f"{key}/obs"][i, t]), (name, i, t)
                    assert close64(r[i], out[f"{key}/rew"][i, t]), (name, i, t, r[i], out[f"{key}/rew"][i, t])
                    assert te[i] == out[f"{key}/term"][i, t], (name, i, t)
        print(f"env_kat {name}: oracle == reference files on {len(seeds)} seeds, lengths {out[f'{key}/length']}")
    np.savez_compressed(os.path.join(HERE, "env_kat.npz"), **out)


class TorchTrajGen:
    """Adapts an oracle.mp generator to the torch-tensor interface BlackBoxWrapper calls
    (black_box_wrapper.py:57,62-65,102,106,113-118,124,226)."""

    def __init__(self, tg):
        import torch
        self.torch = torch
        self.tg = tg
        self.phase_gn = tg.phase_gn

    def set_duration(self, duration, dt):
        self.tg.set_duration(duration, dt)

    def set_params(self, p):
        self.tg.set_params(np.asarray(p))

    def set_initial_conditions(self, t, pos, vel):
        self.tg.set_initial_conditions(t, pos, vel)

    def get_traj_pos(self):
        return self.torch.from_numpy(np.ascontiguousarray(self.tg.get_traj_pos()))

    def get_traj_vel(self):
        return self.torch.from_numpy(np.ascontiguousarray(self.tg.get_traj_vel()))

    def get_params_bounds(self):
        return self.torch.from_numpy(self.tg.get_params_bounds())

    def reset(self):
        self.tg.reset()


def resolved_cfg(env_id, mp_over=None):
    """the id's resolved config with section overrides ({"ctrl": ..., "basis": ..., "phase": ..., "traj": ...}) applied the
    way make_oracle applies them"""
    import copy
    from oracle.blackbox import _merge
    cfg = copy.deepcopy(RESOLVED[env_id])
    for k, v in (mp_over or {}).items():
        cfg[k] = _merge(cfg[k], v)
    return cfg


_SECTION = {"ctrl": "controller_kwargs", "basis": "basis_generator_kwargs", "phase": "phase_generator_kwargs",
            "traj": "trajectory_generator_kwargs"}


def mp_config_override_of(env_id, mp_over, bb_kwargs=None):
    """the same overrides as a fancy_gym `mp_config_override` (whole sections, so that the `_type` replace quirk of
    nested_update, registry.py:272-274, does not drop keys)"""
    cfg = resolved_cfg(env_id, mp_over)
    out = {_SECTION[k]: dict(cfg[k]) for k in (mp_over or {})}
    if bb_kwargs is not None:
        out["black_box_kwargs"] = dict(bb_kwargs)
    return out


def build_reference_bb(ns, env_id, mode, bb_kwargs, env_over=None, mp_over=None):
    """reference BlackBoxWrapper(MPWrapper([TimeAwareObservation](TimeLimit(env)))) as make_bb
    (utils/make_env_helpers.py:68-136) would assemble it, with the oracle MP as traj_gen."""
    cfg = resolved_cfg(env_id, mp_over)
    name = env_id.split("/")[1]
    env = ns.TimeLimit(rl.make_step_env(ns, name, **(env_over or {})), 200)
    if bb_kwargs.get("replanning_schedule") or bb_kwargs.get("learn_sub_trajectories"):
        taw = importlib.import_module("fancy_gym.utils.wrappers").TimeAwareObservation
        env = taw(env)
    wrap = {"HoleReacher-v0": ns.MPWrapper_HoleReacher, "ViaPointReacher-v0": ns.MPWrapper_ViaPoint,
            "SimpleReacher-v0": ns.MPWrapper_SimpleReacher, "LongSimpleReacher-v0": ns.MPWrapper_SimpleReacher}[name]
    env = wrap(env)
    orc = make_oracle(env_id, mode=mode, mp_overrides=dict(mp_over or {}, env=env_over or {}), **bb_kwargs)   # only for an identically configured traj_gen
    ctrl = ns.get_controller(**cfg["ctrl"])
    return ns.BlackBoxWrapper(env, TorchTrajGen(orc.traj_gen), ctrl, duration=2.0, verbose=2, **bb_kwargs)


def n_params_of(env_id, mp_over=None):
    cfg = resolved_cfg(env_id, mp_over)
    per_dof = cfg["basis"]["num_basis"] + (0 if cfg["traj"]["trajectory_generator_type"] == "promp" else 1)
    return cfg["env"]["n_links"] * per_dof


def params_for(env_id, seed, n_plans, sub_traj=False, mp_over=None):
    rng = np.random.default_rng(1234 + seed)
    th = (0.5 * rng.standard_normal((n_plans, n_params_of(env_id, mp_over)))).astype(np.float32)
    if sub_traj:      # learn_sub_trajectories: the learned tau leads the parameter vector; a different one per env and plan
        tau = rng.choice(np.array([0.17, 0.25, 0.37, 0.5, 0.8], dtype=np.float32), size=(n_plans, 1))
        th = np.concatenate([tau, th], axis=1)
    else:             # learned tau / delay lead the parameter vector (tau first): a different phase per env
        ph = (mp_over or {}).get("phase", {})
        lead = []
        if ph.get("learn_tau"):
            lead.append(rng.uniform(0.9, 2.0, size=(n_plans, 1)).astype(np.float32))
        if ph.get("learn_delay"):
            lead.append(rng.uniform(0.0, 0.3, size=(n_plans, 1)).astype(np.float32))
        th = np.concatenate(lead + [th], axis=1)
    return th


_POS = dict(ctrl=dict(controller_type="position"))
_REPLAN25 = dict(replanning_schedule=lambda p, v, o, a, t: t % 25 == 0, max_planning_times=4)
# file name, env id, seeds, black-box kwargs[, env constructor overrides]
BB_CASES = [
    ("bb_holereacher_promp", "fancy_ProMP/HoleReacher-v0", list(range(32)), {}),
    ("bb_viapoint_dmp", "fancy_DMP/ViaPointReacher-v0", list(range(8)), {}),
    ("bb_simplereacher_prodmp", "fancy_ProDMP/SimpleReacher-v0", list(range(8)), {}),
    ("bb_simplereacher_prodmp_replan", "fancy_ProDMP/SimpleReacher-v0", list(range(8)), dict(_REPLAN25, condition_on_desired=False)),
    ("bb_simplereacher_prodmp_replan_cod", "fancy_ProDMP/SimpleReacher-v0", list(range(8)), dict(_REPLAN25, condition_on_desired=True)),
    ("bb_holereacher_promp_replan", "fancy_ProMP/HoleReacher-v0", list(range(8)),
     dict(replanning_schedule=lambda p, v, o, a, t: t % 50 == 0)),
    # the rest of the classic_control x MP matrix registered at envs/__init__.py:38-87
    ("bb_holereacher_dmp", "fancy_DMP/HoleReacher-v0", list(range(6)), {}),
    ("bb_holereacher_prodmp", "fancy_ProDMP/HoleReacher-v0", list(range(6)), {}),
    ("bb_viapoint_promp", "fancy_ProMP/ViaPointReacher-v0", list(range(6)), {}),
    ("bb_viapoint_prodmp", "fancy_ProDMP/ViaPointReacher-v0", list(range(6)), {}),
    ("bb_simplereacher_promp", "fancy_ProMP/SimpleReacher-v0", list(range(6)), {}),
    ("bb_simplereacher_dmp", "fancy_DMP/SimpleReacher-v0", list(range(6)), {}),
    ("bb_longsimplereacher_promp", "fancy_ProMP/LongSimpleReacher-v0", list(range(4)), {}),
    ("bb_longsimplereacher_dmp", "fancy_DMP/LongSimpleReacher-v0", list(range(4)), {}),
    ("bb_longsimplereacher_prodmp", "fancy_ProDMP/LongSimpleReacher-v0", list(range(4)), {}),
    # HoleReacher reward functions through the black-box loop (the "unbounded" latch has to survive re-planning)
    ("bb_holereacher_promp_vel_acc", "fancy_ProMP/HoleReacher-v0", list(range(8)), {}, dict(rew_fct="vel_acc")),
    ("bb_holereacher_promp_unbounded", "fancy_ProMP/HoleReacher-v0", list(range(8)), {}, dict(rew_fct="unbounded")),
    ("bb_holereacher_prodmp_unbounded_replan", "fancy_ProDMP/HoleReacher-v0", list(range(6)),
     dict(replanning_schedule=lambda p, v, o, a, t: t % 60 == 0, condition_on_desired=True), dict(rew_fct="unbounded")),
    # sequencing (learn_sub_trajectories): every plan is round(tau / dt) points long, tau differs per env and plan, so the
    # batched paths see RAGGED plans; condition_on_desired only takes effect on a break (black_box_wrapper.py:196-201)
    ("bb_simplereacher_prodmp_subtraj", "fancy_ProDMP/SimpleReacher-v0", list(range(6)), dict(learn_sub_trajectories=True)),
    ("bb_simplereacher_prodmp_subtraj_cod", "fancy_ProDMP/SimpleReacher-v0", list(range(6)),
     dict(learn_sub_trajectories=True, condition_on_desired=True)),
    ("bb_viapoint_dmp_subtraj", "fancy_DMP/ViaPointReacher-v0", list(range(6)), dict(learn_sub_trajectories=True)),
    ("bb_holereacher_promp_subtraj", "fancy_ProMP/HoleReacher-v0", list(range(6)), dict(learn_sub_trajectories=True)),
    # position controller (pos_controller.py:8-9: the action is the desired POSITION) through every kernel variant that can
    # carry it: run-time K with the registry's 5 basis functions and with 7, the trajectory-from-HBM variant (a learned tau
    # per env), DMP, and the torque-controlled env
    ("bb_holereacher_promp_posctrl", "fancy_ProMP/HoleReacher-v0", list(range(8)), {}, {}, _POS),
    ("bb_holereacher_promp_posctrl_k7", "fancy_ProMP/HoleReacher-v0", list(range(6)), {}, {}, dict(_POS, basis=dict(num_basis=7))),
    ("bb_holereacher_promp_posctrl_tau", "fancy_ProMP/HoleReacher-v0", list(range(6)), {}, {},
     dict(_POS, phase=dict(learn_tau=True, learn_delay=True))),
    ("bb_viapoint_dmp_posctrl", "fancy_DMP/ViaPointReacher-v0", list(range(6)), {}, {}, _POS),
    ("bb_simplereacher_prodmp_posctrl", "fancy_ProDMP/SimpleReacher-v0", list(range(6)), {}, {}, _POS),
]


def bb_case(case):
    """(file name, env id, seeds, black-box kwargs, env overrides, MP section overrides)"""
    case = tuple(case)
    return case + ({},) * (6 - len(case))


T_PAD = 200      # rows the planned trajectories of sub-trajectory cases are zero-padded to (max_episode_steps)


def gen_bb(ns, only=None):
    for fname, env_id, seeds, bbk, env_over, mp_over in map(bb_case, BB_CASES):
        if only and only not in fname:
            continue
        sub_traj = bool(bbk.get("learn_sub_trajectories"))
        n_plans = 8 if (bbk.get("replanning_schedule") or sub_traj) else 1
        rec = {k: [] for k in ("obs0", "params", "positions", "velocities", "step_obs", "step_rewards",
                               "ret", "length", "terminated", "truncated", "obs", "n_calls", "n_points")}
        for s in seeds:
            bb = build_reference_bb(ns, env_id, "shipped", bbk, env_over, mp_over)
            ob0, _ = bb.reset(seed=s)
            th = params_for(env_id, s, n_plans, sub_traj, mp_over)
            per = {k: [] for k in rec if k not in ("obs0", "params", "n_calls")}
            calls = 0
            for i in range(n_plans):
                ob, ret, te, tr, info = bb.step(th[i])
                calls += 1
                L = info["trajectory_length"]
                n_pts = info["positions"].shape[0]
                T = T_PAD if sub_traj else n_pts
                pos_pad = np.zeros((T, info["positions"].shape[1]), np.float32)
                vel_pad = np.zeros_like(pos_pad)
                pos_pad[:n_pts], vel_pad[:n_pts] = info["positions"], info["velocities"]
                so = np.zeros((T, info["step_observations"].shape[1]), np.float32)
                so[:L] = info["step_observations"]
                sr = np.zeros(T)
                sr[:L] = info["step_rewards"]
                for k, v in (("positions", pos_pad), ("velocities", vel_pad), ("step_obs", so),
                             ("step_rewards", sr), ("ret", ret), ("length", L), ("terminated", te),
                             ("truncated", tr), ("obs", ob), ("n_points", n_pts)):
                    per[k].append(np.asarray(v))
                if te or tr:
                    break
            # pad to n_plans calls
            while len(per["ret"]) < n_plans:
                for k in per:
                    per[k].append(np.zeros_like(per[k][0]))
            rec["obs0"].append(ob0); rec["params"].append(th); rec["n_calls"].append(calls)
            for k in per:
                rec[k].append(np.stack(per[k]))
        out = {k: np.stack(v) for k, v in rec.items()}
        out["seeds"] = np.array(seeds)
        np.savez_compressed(os.path.join(HERE, fname + ".npz"), **out)

        # ---- pin the oracle's loop (same 'shipped' MP) against the reference's BlackBoxWrapper ----
        orc = make_oracle(env_id, mode="shipped", verbose=2, mp_overrides=dict(mp_over, env=env_over), **bbk)
        ob0 = orc.reset(seeds=seeds)
        assert np.array_equal(ob0, out["obs0"]), fname
        alive = np.ones(len(seeds), bool)
        for i in range(n_plans):
            ob, ret, te, tr, info = orc.step(out["params"][:, i])
            for b in range(len(seeds)):
                if i >= out["n_calls"][b]:
                    continue
                L = out["length"][b, i]
                assert info["trajectory_length"][b] == L, (fname, b, i, info["trajectory_length"][b], L)
                n = out["n_points"][b, i]
                assert np.array_equal(info["positions"][b][:n], out["positions"][b, i][:n]), (fname, b, i)
                assert np.array_equal(info["velocities"][b][:n], out["velocities"][b, i][:n]), (fname, b, i)
                assert np.array_equal(info["step_observations"][b, :L], out["step_obs"][b, i, :L]), (fname, b, i)
                assert close64(info["step_rewards"][b, :L], out["step_rewards"][b, i, :L]), (fname, b, i)
                assert close64(ret[b], out["ret"][b, i]), (fname, b, i, ret[b], out["ret"][b, i])
                assert te[b] == out["terminated"][b, i] and tr[b] == out["truncated"][b, i], (fname, b, i)
                assert np.array_equal(ob[b], out["obs"][b, i]), (fname, b, i)
        print(f"{fname}: oracle loop == reference BlackBoxWrapper on {len(seeds)} seeds; calls {out['n_calls']}, "
              f"lengths[0] {out['length'][0]}")


if __name__ == "__main__":
    assert rl.available(), "needs /root/reference"
    ns = rl.load()
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None      # e.g. --only subtraj
    if only is None:
        gen_env_kat(ns)
    gen_bb(ns, only)
    print("golden vectors written to", HERE)
