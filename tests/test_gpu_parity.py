"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C-ABI
library via the fancy_gym_b200 facade, against the CPU oracle and the committed golden vectors.

Tolerances (north_star): trajectories / joint states within 1e-5 relative, where "relative" is
|diff| <= 1e-5 * max(|ref|, scale) with scale = 1 for trajectories and pi for angles (angles cross
zero); collision / termination flags and step counts bit-exact except where the
oracle's decision margin is below 1e-5 (documented boundary ties), whose count is bounded.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.blackbox import make_oracle  # noqa: E402
from tests.golden.make_golden import BB_CASES, bb_case, mp_config_override_of, n_params_of  # noqa: E402

TIE_EPS = 1e-5


def _fg():
    import fancy_gym_b200 as fancy_gym
    return fancy_gym


def rel_err(a, b, scale=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), scale)


P_OF = {"fancy_ProMP/HoleReacher-v0": 25, "fancy_DMP/ViaPointReacher-v0": 30, "fancy_ProDMP/SimpleReacher-v0": 12}


# --------------------------------------------------------------------------------------------
# stand-alone trajectory generation (fg_trajgen, K4)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id", list(P_OF))
def test_trajgen_matches_oracle(env_id):
    fancy_gym = _fg()
    B = 257                      # ragged: not a multiple of the warp / block size
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0")
    env.reset(seed=0)
    rng = np.random.default_rng(3)
    params = (0.7 * rng.standard_normal((B, P_OF[env_id]))).astype(np.float32)
    pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
    pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
    out = {}
    for mode in ("mirror", "shipped", "gold"):
        orc = make_oracle(env_id, mode=mode)
        orc.reset(seeds=range(B))
        out[mode] = orc.get_trajectory(params)
    # mirror mode is the kernel's specification: bit-exact
    assert np.array_equal(pos, out["mirror"][0]), np.abs(pos - out["mirror"][0]).max()
    assert np.array_equal(vel, out["mirror"][1]), np.abs(vel - out["mirror"][1]).max()
    # the library's float32 path and the float64 definition
    for mode in ("shipped", "gold"):
        assert rel_err(pos, out[mode][0]).max() < 1e-5
        # velocities are float32 finite differences (ProMP) / recurrences (DMP) in the reference itself;
        # their own rounding noise is ~|pos| * 2^-24 / dt ~ 1e-5 absolute
        vscale = max(1.0, np.abs(out["gold"][1]).max())
        assert rel_err(vel, out[mode][1], scale=vscale).max() < 3e-5


def test_trajgen_empty_and_single():
    fancy_gym = _fg()
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=1, device="cuda:0")
    env.reset(seed=5)
    params = np.zeros(25, dtype=np.float32)
    pos, vel = env.get_trajectory(params)
    assert pos.shape == (1, 200, 5) and float(pos.abs().max()) == 0.0 and float(vel.abs().max()) == 0.0


# --------------------------------------------------------------------------------------------
# fused rollout against the golden vectors produced from the reference's own files
# --------------------------------------------------------------------------------------------
# Tolerances of the golden comparison = 2 x the maxima measured on the B200 over all cases (profiles/r2_golden_errors.json,
# written by this test; DESIGN.md §2 "measured parity").  The goldens carry the reference's float32 MP in BLAS order
# ('shipped'); the CUDA path evaluates float64-built tables with an FMA chain ('mirror').  Velocity-like entries inherit the
# reference's own float32 finite-difference noise ~ 2^-24 |pos| / dt (the float64 definition is as far from the goldens
# as the CUDA path is), which is why they are the widest.
# measured maxima (B200, 28 cases): obs0 0, ret 1.05e-6, obs 2.4e-5, positions 2.0e-6, velocities 1.4e-5, step_obs 2.3e-6,
# step_obs_vel 1.2e-4, step_rewards 4.3e-6
GOLDEN_TOL = dict(obs0=1e-6, ret=2.2e-6, obs=4.8e-5, positions=4.1e-6, velocities=2.9e-5, step_obs=4.6e-6, step_obs_vel=2.4e-4,
                  step_rewards=8.7e-6)
_golden_err = {}


@pytest.mark.parametrize("case", BB_CASES, ids=[c[0] for c in BB_CASES])
def test_rollout_matches_reference_goldens(case, golden_dir):
    fancy_gym = _fg()
    fname, env_id, seeds, bbk, env_over, mp_over = bb_case(case)
    g = np.load(os.path.join(golden_dir, fname + ".npz"))
    override = mp_config_override_of(env_id, mp_over, dict(bbk, verbose=2))
    B = len(seeds)
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override=override, **env_over)
    obs0, _ = env.reset(seed=np.array(seeds), options={"as_numpy": True})
    err = dict.fromkeys(GOLDEN_TOL, 0.0)

    def track(key, value):
        err[key] = max(err[key], float(np.max(value, initial=0.0)))

    track("obs0", np.abs(obs0 - g["obs0"]))
    n_plans = g["params"].shape[1]
    last_obs = {}
    nl = env.unwrapped.n_links
    for i in range(n_plans):
        live = i < g["n_calls"]
        if not live.any():
            break
        obs, ret, te, tr, info = env.step(g["params"][:, i])
        for b in np.nonzero(~live)[0]:      # episode over in an earlier call: frozen, reports its last observation, 0 steps
            assert info["trajectory_length"][b] == 0 and np.array_equal(obs[b], last_obs[b]), (fname, b, i)
        for b in np.nonzero(live)[0]:
            last_obs[b] = np.array(obs[b])
            assert info["trajectory_length"][b] == g["length"][b, i], (fname, b, i)
            assert bool(te[b]) == bool(g["terminated"][b, i]) and bool(tr[b]) == bool(g["truncated"][b, i])
            r_ref = g["ret"][b, i]
            if np.isfinite(r_ref):
                track("ret", rel_err(ret[b], r_ref))
            else:
                assert ret[b] == r_ref
            track("obs", np.abs(obs[b] - g["obs"][b, i]) / np.maximum(1.0, np.abs(g["obs"][b, i])))
            # verbose=2 infos: the planned trajectory and the per-step observations / rewards of the reference's loop
            L = g["length"][b, i]
            n = g["n_points"][b, i]         # sub-trajectories: ragged plans
            track("positions", rel_err(info["positions"][b][:n], g["positions"][b, i][:n]))
            vscale = max(1.0, np.abs(g["velocities"][b, i]).max())
            track("velocities", rel_err(info["velocities"][b][:n], g["velocities"][b, i][:n], scale=vscale))
            so, so_ref = info["step_observations"][b, :L], g["step_obs"][b, i, :L]
            e = np.abs(so - so_ref) / np.maximum(1.0, np.abs(so_ref))
            is_vel = np.zeros(so.shape[1], bool)
            is_vel[2 * nl:3 * nl] = True       # joint velocities: float32 finite differences of the trajectory (see above)
            track("step_obs", e[:, ~is_vel])
            track("step_obs_vel", e[:, is_vel])
            sr, sr_ref = info["step_rewards"][b, :L], g["step_rewards"][b, i, :L]
            fin = np.isfinite(sr_ref)
            assert np.array_equal(sr[~fin], sr_ref[~fin])
            # per-step rewards are dominated by 5e-8 * sum(acc^2) with acc = dv/dt of float32 finite differences
            track("step_rewards", np.abs(sr[fin] - sr_ref[fin]) / np.maximum(1.0, np.abs(sr_ref[fin])))
    _golden_err[fname] = err
    os.makedirs("gpurun_out", exist_ok=True)
    import json
    with open(os.path.join("gpurun_out", "golden_errors.json"), "w") as f:
        worst = {k: max(e[k] for e in _golden_err.values()) for k in GOLDEN_TOL}
        json.dump({"tolerance": GOLDEN_TOL, "measured_max_over_cases": worst, "per_case": _golden_err}, f, indent=1)
    for k, tol in GOLDEN_TOL.items():
        assert err[k] <= tol, (fname, k, err[k], tol)


# --------------------------------------------------------------------------------------------
# fused rollout against the oracle on seeded random inputs (thousands of envs)
# --------------------------------------------------------------------------------------------
ALL_IDS = [f"fancy_{mp}/{name}" for name in ("HoleReacher-v0", "ViaPointReacher-v0", "SimpleReacher-v0", "LongSimpleReacher-v0")
           for mp in ("ProMP", "DMP", "ProDMP")]
RANDOM_CASES = [("fancy_ProMP/HoleReacher-v0", 0.25, {}), ("fancy_ProMP/HoleReacher-v0", 1.0, {}),
                ("fancy_DMP/ViaPointReacher-v0", 1.0, {}), ("fancy_ProDMP/SimpleReacher-v0", 1.0, {})]
RANDOM_CASES += [(i, 0.5, {}) for i in ALL_IDS if i not in [c[0] for c in RANDOM_CASES]]
RANDOM_CASES += [("fancy_ProMP/HoleReacher-v0", 0.5, dict(rew_fct="vel_acc")), ("fancy_ProMP/HoleReacher-v0", 0.5, dict(rew_fct="unbounded")),
                 ("fancy_ProMP/HoleReacher-v0", 1.0, dict(allow_self_collision=True)),
                 ("fancy_ProMP/HoleReacher-v0", 1.0, dict(allow_wall_collision=True, hole_x=1.0, hole_width=0.4, hole_depth=0.7)),
                 ("fancy_DMP/ViaPointReacher-v0", 1.0, dict(allow_self_collision=True, random_start=True))]


@pytest.mark.parametrize("env_id,sigma,env_over", RANDOM_CASES,
                         ids=[f"{c[0]}-{c[1]}" + "".join(f"-{k}={v}" for k, v in c[2].items()) for c in RANDOM_CASES])
def test_rollout_matches_oracle_random(env_id, sigma, env_over):
    fancy_gym = _fg()
    B = 2048 + 37 if not env_over and env_id in P_OF else 512 + 5
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", **env_over)
    env.reset(seed=100)
    rng = np.random.default_rng(11)
    params = (sigma * rng.standard_normal((B, n_params_of(env_id)))).astype(np.float32)
    obs, ret, te, tr, info = env.step(torch.as_tensor(params, device="cuda:0"))
    obs, ret, te, tr = obs.cpu().numpy(), ret.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()
    length = info["trajectory_length"].cpu().numpy()

    orc = make_oracle(env_id, mode="mirror", mp_overrides={"env": env_over})
    orc.reset(seeds=100 + np.arange(B))
    o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
    tie = o_info["min_margin"] < TIE_EPS
    agree = (length == o_info["trajectory_length"]) & (te == o_te) & (tr == o_tr)
    assert (agree | tie).all(), f"{(~(agree | tie)).sum()} flag/length mismatches outside boundary ties"
    assert (~agree).sum() <= max(2, B // 500), f"too many boundary ties resolved differently: {(~agree).sum()}"
    m = agree
    fin = m & np.isfinite(o_ret)
    assert not fin.any() or rel_err(ret[fin], o_ret[fin]).max() < 1e-5
    assert np.array_equal(ret[m & ~np.isfinite(o_ret)], o_ret[m & ~np.isfinite(o_ret)])
    oscale = np.maximum(1.0, np.abs(o_obs[m]))
    assert (np.abs(obs[m] - o_obs[m]) <= 1e-5 * oscale).all()
    for k in ("is_success", "is_collided"):
        if k in info:
            assert np.array_equal(info[k].cpu().numpy()[m], o_info[k][m])
    if "joints" in o_info:
        assert rel_err(info["joints"].cpu().numpy()[m], o_info["joints"][m], scale=np.pi).max() < 1e-5
    if "end_effector" in info:
        assert rel_err(info["end_effector"].cpu().numpy()[m], o_info["end_effector"][m], scale=5.0).max() < 1e-5
    if "reward_dist" in info:
        assert rel_err(info["reward_dist"].cpu().numpy()[m], o_info["reward_dist"][m]).max() < 1e-5
        assert rel_err(info["reward_ctrl"].cpu().numpy()[m], o_info["reward_ctrl"][m]).max() < 1e-5


# --------------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes
# --------------------------------------------------------------------------------------------
def _run(env, params, seed):
    env.reset(seed=seed)
    obs, ret, te, tr, info = env.step(params)
    return obs.clone(), ret.clone(), te.clone(), tr.clone(), info["trajectory_length"].clone()


def test_wall_modes_agree_at_full_size():
    """The transition-search wall test (mode 0: estimate + fix-up, mode 3: bisection) must give exactly the flags of the
    literal 100-samples-per-link evaluation (mode 1: with the exact skip of links above ground, mode 2: no skipping at all)."""
    fancy_gym = _fg()
    B = 65536
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    params = torch.randn(B, 25, generator=gen, device="cuda:0")
    outs = []
    for mode in (0, 1, 2, 3):
        env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device="cuda:0", context_sampler="device",
                             mp_config_override={"black_box_kwargs": {"wall_mode": mode}})
        outs.append(_run(env, params, 7))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)
    assert int(outs[0][2].sum()) > B // 10      # the comparison saw many collisions


def test_determinism_and_shard_equivalence():
    """Same inputs -> identical outputs; a batch split in two halves == the full batch (envs are independent)."""
    fancy_gym = _fg()
    B = 65536
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    params = 0.5 * torch.randn(B, 25, generator=gen, device="cuda:0")
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device="cuda:0", context_sampler="device")
    a = _run(env, params, 3)
    b = _run(env, params, 3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    ctx, q0 = env.unwrapped.ctx.clone(), None
    env.reset(seed=3)
    q0 = env.unwrapped.q[:, 0].clone()
    half = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B // 2, device="cuda:0")
    for lo in (0, B // 2):
        sl = slice(lo, lo + B // 2)
        half.reset(options={"contexts": dict(x=ctx[sl, 0].cpu().numpy(), width=ctx[sl, 1].cpu().numpy(),
                                             depth=ctx[sl, 2].cpu().numpy(), q0=q0[sl].cpu().numpy())})
        obs, ret, te, tr, info = half.step(params[sl])
        assert torch.equal(ret, a[1][sl]) and torch.equal(info["trajectory_length"], a[4][sl])
        assert torch.equal(te, a[2][sl]) and torch.equal(obs, a[0][sl])


def test_zero_params_keep_the_arm_still():
    """All-zero ProMP weights: zero velocity, the straight arm never collides with itself (collinear links,
    exact zeros in the orientation tests), 200 steps, return == -dist^2 of the start pose."""
    fancy_gym = _fg()
    B = 1024
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device="cuda:0")
    env.reset(seed=0)
    ee0 = env.unwrapped.end_effector()
    goal = torch.stack([env.unwrapped.ctx[:, 0], -env.unwrapped.ctx[:, 2]], dim=1)
    obs, ret, te, tr, info = env.step(torch.zeros(B, 25, device="cuda:0"))
    assert bool((info["trajectory_length"] == 200).all()) and not bool(te.any()) and bool(tr.all())
    want = -((ee0 - goal) ** 2).sum(1)
    assert torch.allclose(ret, want, rtol=1e-12, atol=0)


# --------------------------------------------------------------------------------------------
# device-side reset (fg_reset): numpy-exact context streams
# --------------------------------------------------------------------------------------------
RESET_CASES = [("fancy/HoleReacher-v0", {}), ("fancy/HoleReacher-v0", dict(hole_x=1.0, random_start=False)),
               ("fancy/HoleReacher-v0", dict(hole_width=0.3, hole_depth=None)),
               ("fancy/ViaPointReacher-v0", {}), ("fancy/ViaPointReacher-v0", dict(random_start=True, target=(3.0, 2.0))),
               ("fancy/SimpleReacher-v0", {}), ("fancy/SimpleReacher-v0", dict(random_start=False)),
               ("fancy/LongSimpleReacher-v0", dict(target=(1.0, -2.0)))]


@pytest.mark.parametrize("env_id,kw", RESET_CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(RESET_CASES)])
def test_device_reset_reproduces_numpy_streams(env_id, kw):
    """env i reset with seed s on the device == numpy's Generator(PCG64(SeedSequence(s + i))) in the reference's draw order
    (bit-exact contexts and start angles), also for explicit per-env seeds, 64-bit seeds and unseeded follow-up resets
    that continue the per-env streams."""
    fancy_gym = _fg()
    B = 1000
    dev = fancy_gym.make(env_id, num_envs=B, device="cuda:0", context_sampler="device", **kw)
    ref = fancy_gym.make(env_id, num_envs=B, device="cuda:0", context_sampler="numpy", **kw)
    rng = np.random.default_rng(0)
    for seed in (0, 12345, rng.integers(0, 2**62, size=B), None, None, 2**40 + 17, None):
        o_dev, _ = dev.reset(seed=seed)
        o_ref, _ = ref.reset(seed=seed)
        assert torch.equal(dev.ctx, ref.ctx), seed
        assert torch.equal(dev.q, ref.q), seed
        assert float(dev.v.abs().max()) == 0.0 and int(dev.steps.abs().max()) == 0 and int(dev.done.max()) == 0
        assert torch.allclose(o_dev, o_ref, rtol=0, atol=1e-6)


def test_blackbox_reset_fast_path_matches_wrapper_chain():
    fancy_gym = _fg()
    B = 513
    for env_id, bbk in (("fancy_ProMP/HoleReacher-v0", {}), ("fancy_DMP/ViaPointReacher-v0", {}),
                        ("fancy_ProDMP/SimpleReacher-v0", {"replanning_schedule": lambda p, v, o, a, t: t % 25 == 0})):
        over = {"black_box_kwargs": bbk}
        fast = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override=over)
        slow = fancy_gym.make(env_id, num_envs=B, device="cuda:0", context_sampler="numpy", mp_config_override=over)
        assert fast._fast_reset and not slow._fast_reset
        a, _ = fast.reset(seed=5)
        b, _ = slow.reset(seed=5)
        assert a.shape == b.shape == (B, fast.observation_space.shape[0])
        assert torch.allclose(a, b, rtol=0, atol=1e-6)


# --------------------------------------------------------------------------------------------
# per-env learned tau / delay: basis evaluated in the kernel, rollout from the HBM trajectory
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id,phase", [("fancy_ProMP/HoleReacher-v0", dict(learn_tau=True, learn_delay=True)),
                                          ("fancy_ProMP/HoleReacher-v0", dict(learn_delay=True)),
                                          ("fancy_DMP/ViaPointReacher-v0", dict(learn_tau=True)),
                                          ("fancy_DMP/HoleReacher-v0", dict(learn_tau=True, learn_delay=True)),
                                          ("fancy_ProMP/SimpleReacher-v0", dict(learn_tau=True)),
                                          ("fancy_ProDMP/SimpleReacher-v0", dict(learn_tau=True)),
                                          ("fancy_ProDMP/HoleReacher-v0", dict(learn_tau=True, learn_delay=True))],
                         ids=["promp-tau-delay", "promp-delay", "dmp-tau", "dmp-tau-delay", "promp-pd-tau", "prodmp-tau",
                              "prodmp-tau-delay"])
def test_per_env_tau_delay_matches_oracle(env_id, phase):
    fancy_gym = _fg()
    B = 300 + 7
    mp_type = "promp" if "ProMP" in env_id else ("prodmp" if "ProDMP" in env_id else "dmp")
    base_phase = dict(RESOLVED_PHASE(env_id), **phase)
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override={"phase_generator_kwargs": base_phase})
    env.reset(seed=50)
    rng = np.random.default_rng(4)
    n_extra = int(phase.get("learn_tau", False)) + int(phase.get("learn_delay", False))
    params = (0.4 * rng.standard_normal((B, n_params_of(env_id) + n_extra))).astype(np.float32)
    i = 0
    if phase.get("learn_tau"):
        # partly outside tau_bound = [0.02, 2.0]: clipped like the reference (ProDMP pre-computes 6 tau: keep t / tau <= 6)
        params[:, i] = rng.uniform(0.45 if mp_type == "prodmp" else 0.3, 2.5, B)
        i += 1
    if phase.get("learn_delay"):
        params[:, i] = rng.uniform(0.0, 0.6, B)
    assert env.action_space.shape[0] == params.shape[1]

    orc = make_oracle(env_id, mode="mirror", mp_overrides={"phase": phase})
    orc.reset(seeds=50 + np.arange(B))
    o_pos, o_vel = orc.get_trajectory(params)
    pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
    pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
    # same arithmetic as the shared-table path (float64 transcendental part rounded once): equal up to the last-bit
    # differences of exp() between libm and CUDA
    assert np.abs(pos - o_pos).max() <= 2e-6 * max(1.0, np.abs(o_pos).max())
    vscale = max(1.0, np.abs(o_vel).max())
    assert np.abs(vel - o_vel).max() <= 3e-5 * vscale
    if phase.get("learn_tau") and phase.get("learn_delay") and mp_type == "promp":
        # the linear phase saturates: constant trajectory after delay + tau, constant before the delay
        tau = np.clip(params[:, 0], 0.02, 2.0)
        delay = np.clip(params[:, 1], 0, 1.98)
        late = (delay + tau) < 1.8
        assert late.any()
        for b in np.nonzero(late)[0][:20]:
            n_end = int(np.ceil((delay[b] + tau[b]) / 0.01)) + 1
            assert np.all(pos[b, n_end:] == pos[b, -1])

    orc.reset(seeds=50 + np.arange(B))
    o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
    obs, ret, te, tr, info = env.step(torch.as_tensor(params, device="cuda:0"))
    obs, ret, te, tr = obs.cpu().numpy(), ret.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()
    length = info["trajectory_length"].cpu().numpy()
    tie = o_info["min_margin"] < TIE_EPS
    agree = (length == o_info["trajectory_length"]) & (te == o_te) & (tr == o_tr)
    assert (agree | tie).all() and (~agree).sum() <= 3
    fin = agree & np.isfinite(o_ret)
    assert not fin.any() or rel_err(ret[fin], o_ret[fin]).max() < 1e-5
    assert (np.abs(obs[agree] - o_obs[agree]) <= 2e-5 * np.maximum(1.0, np.abs(o_obs[agree]))).all()


def RESOLVED_PHASE(env_id):
    from oracle.blackbox import RESOLVED
    return dict(RESOLVED[env_id]["phase"])


# --------------------------------------------------------------------------------------------
# generic kernel paths: run-time basis count (not the registry default 5), other trajectory lengths, other link counts
# --------------------------------------------------------------------------------------------
GENERIC_CASES = [("fancy_ProMP/HoleReacher-v0", {"basis": dict(num_basis=7)}, {}),
                 ("fancy_ProMP/HoleReacher-v0", {"basis": dict(num_basis=3, num_basis_zero_start=2)}, {}),
                 ("fancy_DMP/ViaPointReacher-v0", {"basis": dict(num_basis=8)}, {}),
                 ("fancy_ProDMP/SimpleReacher-v0", {"basis": dict(num_basis=4)}, {}),
                 ("fancy_ProDMP/HoleReacher-v0", {"basis": dict(num_basis=6), "phase": dict(tau=2.0)}, {}),
                 ("fancy_ProMP/SimpleReacher-v0", {"basis": dict(num_basis=6)}, dict(n_links=3))]


@pytest.mark.parametrize("env_id,over,env_over", GENERIC_CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(GENERIC_CASES)])
def test_generic_kernel_paths_match_oracle(env_id, over, env_over):
    fancy_gym = _fg()
    B = 200 + 3
    from oracle.blackbox import RESOLVED
    names = {"basis": "basis_generator_kwargs", "phase": "phase_generator_kwargs"}
    mp_over = {names[k]: dict(RESOLVED[env_id][k], **v) for k, v in over.items()}
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override=mp_over, **env_over)
    env.reset(seed=9)
    P = env.action_space.shape[0]
    params = (0.5 * np.random.default_rng(2).standard_normal((B, P))).astype(np.float32)
    orc = make_oracle(env_id, mode="mirror", mp_overrides=dict(over, env=env_over))
    orc.reset(seeds=9 + np.arange(B))
    o_pos, o_vel = orc.get_trajectory(params)
    pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
    assert np.array_equal(pos.cpu().numpy(), o_pos) and np.array_equal(vel.cpu().numpy(), o_vel)     # mirror mode: bit-exact
    o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
    obs, ret, te, tr, info = env.step(torch.as_tensor(params, device="cuda:0"))
    length = info["trajectory_length"].cpu().numpy()
    tie = o_info["min_margin"] < TIE_EPS
    agree = (length == o_info["trajectory_length"]) & (te.cpu().numpy() == o_te) & (tr.cpu().numpy() == o_tr)
    assert (agree | tie).all()
    fin = agree & np.isfinite(o_ret)
    assert not fin.any() or rel_err(ret.cpu().numpy()[fin], o_ret[fin]).max() < 1e-5
    assert (np.abs(obs.cpu().numpy()[agree] - o_obs[agree]) <= 1e-5 * np.maximum(1.0, np.abs(o_obs[agree]))).all()


def test_trajgen_ragged_lengths_and_long_trajectories():
    """T not a multiple of 4 (ragged last quad) and T * dof beyond the staging buffer (chunked path)"""
    fancy_gym = _fg()
    from tests.toy import ToyWrapper, register_toy
    register_toy(fancy_gym)
    for mp_type, basis in (("promp", "rbf"), ("prodmp", "prodmp")):
        for dof in (1, 5):
            env = fancy_gym.make_bb("toy-v0", [ToyWrapper], {}, {"trajectory_generator_type": mp_type, "action_dim": dof},
                                    {"controller_type": "motor"}, {"phase_generator_type": "exp"},
                                    {"basis_generator_type": basis, "num_basis": 5}, device="cuda:0", num_envs=33, n_links=dof)
            tg = env.traj_gen
            for duration in (0.98, 1.0, 4.22, 30.0 if mp_type == "promp" else 5.5):   # T = 49, 50, 211, 1500 (275: ProDMP pre-computes 6 tau)
                tg.set_duration(duration, 0.02)
                p = torch.randn(33, tg.num_params, device="cuda:0")
                tg.set_params(p)
                tg.set_initial_conditions(0.0, torch.ones(33, dof, device="cuda:0"), torch.zeros(33, dof, device="cuda:0"))
                pos, vel = tg._run_trajgen()
                T = int(round(duration / 0.02))
                assert pos.shape == (33, T, dof) and bool(torch.isfinite(pos).all()) and bool(torch.isfinite(vel).all())
                tb = tg.tables()
                w = p.reshape(33, dof, -1).cpu().numpy().astype(np.float32)
                if mp_type == "promp":      # pos = fma chain over the table row; vel = finite difference, last row duplicated
                    ref = np.zeros((33, T, dof), np.float32)
                    for k in range(tb.tab_a.shape[1]):
                        ref = (np.float32(tb.tab_a[None, :, None, k]) * w[:, None, :, k]).astype(np.float64) + ref
                        ref = ref.astype(np.float32)
                    assert np.abs(pos.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
                    v = vel.cpu().numpy()
                    assert np.array_equal(v[:, -1], v[:, -2])


def test_step_results_stay_valid_for_one_more_step():
    """step() returns views of one of two alternating result sets: what step i returned must be untouched by step i+1
    (and reset), and is recycled by step i+2."""
    fancy_gym = _fg()
    B = 1024
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device="cuda:0")
    env.reset(seed=0)
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    p1, p2 = (0.5 * torch.randn(B, 25, generator=gen, device="cuda:0") for _ in range(2))
    obs1, ret1, te1, tr1, info1 = env.step(p1)
    snap = [x.clone() for x in (obs1, ret1, te1, tr1, info1["trajectory_length"], info1["is_collided"], info1["end_effector"])]
    env.reset(seed=1)
    obs2, ret2, te2, tr2, info2 = env.step(p2)
    for a, b in zip(snap, (obs1, ret1, te1, tr1, info1["trajectory_length"], info1["is_collided"], info1["end_effector"])):
        assert torch.equal(a, b)
    assert not torch.equal(ret1, ret2)
    assert ret1.data_ptr() != ret2.data_ptr()
    # terminated / truncated are exactly the flag bits
    assert torch.equal(te2, (env._flags & 1) != 0) and torch.equal(tr2, (env._flags & 2) != 0)
    assert torch.equal(info2["is_success"], (env._flags & 4) != 0) and torch.equal(info2["is_collided"], (env._flags & 8) != 0)


# --------------------------------------------------------------------------------------------
# sequencing (learn_sub_trajectories, BASELINE config 4's second variant): every black-box step plans a sub-trajectory
# whose length follows the learned tau; the batch shares tau per step, tau changes from step to step
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id", ["fancy_ProDMP/SimpleReacher-v0", "fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0"])
def test_sub_trajectory_sequencing_matches_oracle(env_id):
    fancy_gym = _fg()
    B = 129
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override={"black_box_kwargs": {"learn_sub_trajectories": True}})
    orc = make_oracle(env_id, mode="mirror", learn_sub_trajectories=True)
    obs0, _ = env.reset(seed=21, options={"as_numpy": True})
    o_obs0 = orc.reset(seeds=21 + np.arange(B))
    assert np.allclose(obs0, o_obs0, atol=1e-6) and obs0.shape[1] == env.observation_space.shape[0]
    rng = np.random.default_rng(8)
    P = env.action_space.shape[0]
    assert P == n_params_of(env_id) + 1                      # tau leads the parameter vector
    total = np.zeros(B, dtype=np.int64)
    was_live = np.ones(B, bool)
    ever_terminated = np.zeros(B, bool)
    for tau in (0.37, 0.8, 0.55, 2.0):                       # 37 + 80 + 55 steps, then the rest of the 200-step episode
        params = (0.4 * rng.standard_normal((B, P))).astype(np.float32)
        params[:, 0] = tau
        o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
        obs, ret, te, tr, info = env.step(params)
        tie = o_info["min_margin"] < TIE_EPS
        agree = (info["trajectory_length"] == o_info["trajectory_length"]) & (te == o_te) & (tr == o_tr)
        assert (agree | tie).all(), (tau, int((~(agree | tie)).sum()))
        if tie.any() and not agree.all():
            pytest.skip("a boundary tie resolved differently: the two sides diverge from here on")
        fin = np.isfinite(o_ret)
        assert rel_err(ret[fin], o_ret[fin]).max() < 1e-5 if fin.any() else True
        assert (np.abs(obs - o_obs) <= 2e-5 * np.maximum(1.0, np.abs(o_obs))).all(), tau
        total += info["trajectory_length"]
        ran_through = was_live & ~(te | tr)                  # planned this sub-trajectory and neither terminated nor ran out of time
        if tau < 2.0:
            assert (info["trajectory_length"][ran_through] == round(tau / 0.01)).all()
        assert (info["trajectory_length"][~was_live] == 0).all()
        was_live &= ~(te | tr)
        ever_terminated |= te
    assert (total <= 200).all() and (total[~ever_terminated] == 200).all()


@pytest.mark.parametrize("env_id", ["fancy_ProDMP/SimpleReacher-v0", "fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0"])
def test_ragged_sub_trajectories_match_oracle(env_id):
    """learn_sub_trajectories where every env of the batch learns its OWN tau: env b plans round(tau_b / dt) points on its
    own time grid (fg_phase_basis.n_steps_env / times_table) and the fused rollout stops it after that many steps
    (fg_rollout_io.seg_steps_env).  The oracle's batched version is pinned to its scalar loop in
    tests/test_oracle_golden.py::test_ragged_sub_trajectories_equal_the_scalar_loop."""
    fancy_gym = _fg()
    B = 193
    env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override={"black_box_kwargs": {"learn_sub_trajectories": True}})
    orc = make_oracle(env_id, mode="mirror", learn_sub_trajectories=True)
    env.reset(seed=5)
    orc.reset(seeds=5 + np.arange(B))
    rng = np.random.default_rng(11)
    P = env.action_space.shape[0]
    total = np.zeros(B, dtype=np.int64)
    live = np.ones(B, bool)
    for k in range(5):
        params = (0.4 * rng.standard_normal((B, P))).astype(np.float32)
        params[:, 0] = rng.uniform(0.05, 0.9, size=B).astype(np.float32)
        if k == 0:      # the stand-alone trajectory of a ragged batch: every env's own rows, zero beyond
            pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
            o_pos, o_vel = orc.get_trajectory(params)
            n_valid = np.asarray(orc.traj_gen.n_valid)
            assert pos.shape[1] == env.traj_gen.n_steps >= n_valid.max()
            pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
            for b in range(B):
                n = n_valid[b]
                # (same arithmetic as the shared-table path; equal up to last-bit differences of exp() between libm and CUDA)
                assert np.abs(pos[b, :n] - o_pos[b, :n]).max() <= 2e-6 * max(1.0, np.abs(o_pos[b, :n]).max()), b
                assert np.abs(vel[b, :n] - o_vel[b, :n]).max() <= 3e-5 * max(1.0, np.abs(o_vel[b, :n]).max()), b
                assert not pos[b, n:].any() and not vel[b, n:].any()
        o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
        obs, ret, te, tr, info = env.step(params)
        tie = o_info["min_margin"] < TIE_EPS
        agree = (info["trajectory_length"] == o_info["trajectory_length"]) & (te == o_te) & (tr == o_tr)
        assert (agree | tie).all(), (k, int((~(agree | tie)).sum()))
        if tie.any() and not agree.all():
            pytest.skip("a boundary tie resolved differently: the two sides diverge from here on")
        fin = np.isfinite(o_ret)
        assert rel_err(ret[fin], o_ret[fin]).max() < 1e-5 if fin.any() else True
        assert (np.abs(obs - o_obs) <= 2e-5 * np.maximum(1.0, np.abs(o_obs))).all(), k
        ran_through = live & ~(te | tr)
        want = np.round(params[:, 0].astype(np.float64) / 0.01).astype(np.int64)
        assert (info["trajectory_length"][ran_through] == want[ran_through]).all()
        if k == 0:
            assert len(set(info["trajectory_length"][ran_through].tolist())) > 10        # the batch really is ragged
        assert (info["trajectory_length"][~live] == 0).all()
        live &= ~(te | tr)
        total += info["trajectory_length"]
    assert (total <= 200).all()


def test_env_on_a_non_current_device():
    """every entry point selects the device of its buffers itself (one process driving cuda:1 while cuda:0 is current)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fancy_gym = _fg()
    B = 300
    torch.cuda.set_device(0)
    params = (0.5 * np.random.default_rng(0).standard_normal((B, 25))).astype(np.float32)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev)
        env.reset(seed=4)
        obs, ret, te, tr, info = env.step(torch.as_tensor(params, device=dev))
        pos, vel = env.get_trajectory(torch.as_tensor(params, device=dev))
        assert obs.device == torch.device(dev) and pos.device == torch.device(dev)
        outs.append([x.cpu() for x in (obs, ret, te, info["trajectory_length"], pos, vel)])
    assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_evaluate_leaves_the_envs_where_they_are():
    """evaluate(params) == what step(params) would return, but the episode does not advance: candidates can be compared on
    one reset state; a following step() with the same parameters returns the identical result."""
    fancy_gym = _fg()
    B = 777
    for env_id in ("fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0", "fancy_ProDMP/SimpleReacher-v0"):
        env = fancy_gym.make(env_id, num_envs=B, device="cuda:0")
        env.reset(seed=2)
        base = env.unwrapped
        q0, steps0 = base.q.clone(), base.steps.clone()
        gen = torch.Generator(device="cuda:0").manual_seed(1)
        P = env.action_space.shape[0]
        p1, p2 = (0.5 * torch.randn(B, P, generator=gen, device="cuda:0") for _ in range(2))
        e1 = [x.clone() for x in env.evaluate(p1)[:4]]
        e2 = [x.clone() for x in env.evaluate(p2)[:4]]
        assert torch.equal(base.q, q0) and torch.equal(base.steps, steps0) and int(base.done.max()) == 0
        e1b = env.evaluate(p1)[:4]
        for a, b in zip(e1, e1b):
            assert torch.equal(a, b, ) if a.dtype != torch.float64 else torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
        s2 = env.step(p2)[:4]
        for a, b in zip(e2, s2):
            assert torch.equal(torch.nan_to_num(a.double()), torch.nan_to_num(b.double()))
        assert not torch.equal(base.q, q0) and int(base.steps.max()) > 0
