"""The switchable readings of mp_pytorch on the CUDA path (fancy_gym_b200/mp/assumptions.py): under every flipped switch
the kernels' trajectories and rollouts equal the oracle built under the same switch — bit-exact where the switch only moves
table entries / parameters, 1e-6 for the DMP start-of-recurrence switch (its extra Euler step is a few float32 torch ops
whose forcing contraction is a plain sum, not the kernels' FMA chain)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from fancy_gym_b200.mp import assumptions as pa  # noqa: E402
from oracle import mp as omp  # noqa: E402
from oracle.blackbox import RESOLVED, make_oracle  # noqa: E402

TIE_EPS = 1e-5
CASES = [
    ("fancy_ProDMP/SimpleReacher-v0", dict(exp_phase_right_clip=False), {}, True),
    ("fancy_ProDMP/SimpleReacher-v0", dict(alpha_phase_default=2.0), {}, True),
    ("fancy_ProDMP/SimpleReacher-v0", dict(prodmp_interpolate=True), {}, True),
    ("fancy_ProDMP/SimpleReacher-v0", dict(scale_on_library_side=False), {"traj": dict(weights_scale=0.7, goal_scale=1.3)}, True),
    ("fancy_ProDMP/HoleReacher-v0", dict(exp_phase_right_clip=False, alpha_phase_default=2.5), {}, True),
    ("fancy_DMP/ViaPointReacher-v0", dict(exp_phase_right_clip=False), {"phase": dict(tau=1.2)}, True),
    ("fancy_DMP/ViaPointReacher-v0", dict(scale_on_library_side=False), {}, True),
    ("fancy_DMP/HoleReacher-v0", dict(dmp_init_on_first_grid_point=False), {}, False),
    ("fancy_DMP/SimpleReacher-v0", dict(dmp_init_on_first_grid_point=False, exp_phase_right_clip=False), {"phase": dict(tau=1.5)}, False),
    ("fancy_ProMP/HoleReacher-v0", dict(scale_on_library_side=False), {"traj": dict(weights_scale=1.7)}, True),
    # per-env phase (learned tau / delay): the switches reach the in-kernel basis evaluation (fg_phase_basis)
    ("fancy_DMP/ViaPointReacher-v0", dict(exp_phase_right_clip=False, scale_on_library_side=False), {"phase": dict(learn_tau=True)}, None),
    ("fancy_ProDMP/SimpleReacher-v0", dict(exp_phase_right_clip=False), {"phase": dict(learn_tau=True)}, None),
    ("fancy_DMP/HoleReacher-v0", dict(dmp_init_on_first_grid_point=False), {"phase": dict(learn_tau=True, learn_delay=True)}, None),
]
_SEC = {"phase": "phase_generator_kwargs", "traj": "trajectory_generator_kwargs", "basis": "basis_generator_kwargs"}


@pytest.mark.parametrize("env_id,flips,over,exact", CASES,
                         ids=[f"{c[0]}-" + "+".join(f"{k}={v}" for k, v in c[1].items()) + ("-per-env" if c[3] is None else "")
                              for c in CASES])
def test_cuda_path_follows_the_switches(env_id, flips, over, exact):
    import fancy_gym_b200 as fancy_gym
    B = 301
    mp_over = {_SEC[k]: dict(RESOLVED[env_id][k], **v) for k, v in over.items()}
    with pa.assume(**flips):
        env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override=mp_over)
    with omp.assume(**flips):
        orc = make_oracle(env_id, mode="mirror", mp_overrides=over)
    env.reset(seed=70)
    orc.reset(seeds=70 + np.arange(B))
    P = env.action_space.shape[0]
    rng = np.random.default_rng(6)
    params = (0.5 * rng.standard_normal((B, P))).astype(np.float32)
    ph = over.get("phase", {})
    i = 0
    if ph.get("learn_tau"):
        params[:, i] = rng.uniform(0.6, 2.2, B)
        i += 1
    if ph.get("learn_delay"):
        params[:, i] = rng.uniform(0.0, 0.4, B)
    o_pos, o_vel = orc.get_trajectory(params)
    pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
    pos, vel = pos.cpu().numpy(), vel.cpu().numpy()
    if exact:
        assert np.array_equal(pos, o_pos) and np.array_equal(vel, o_vel), (np.abs(pos - o_pos).max(), np.abs(vel - o_vel).max())
    else:       # per-env phase: libm vs CUDA exp(); DMP pre-step: plain-sum forcing
        assert np.abs(pos - o_pos).max() <= 2e-6 * max(1.0, np.abs(o_pos).max())
        assert np.abs(vel - o_vel).max() <= 3e-5 * max(1.0, np.abs(o_vel).max())
    # and the switch really is in effect: the default reading gives another trajectory
    base = make_oracle(env_id, mode="mirror", mp_overrides=over)
    base.reset(seeds=70 + np.arange(B))
    b_pos, _ = base.get_trajectory(params)
    assert np.abs(b_pos - o_pos).max() > 0

    o_obs, o_ret, o_te, o_tr, o_info = orc.step(params)
    obs, ret, te, tr, info = env.step(torch.as_tensor(params, device="cuda:0"))
    obs, ret, te, tr = obs.cpu().numpy(), ret.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()
    length = info["trajectory_length"].cpu().numpy()
    tie = o_info["min_margin"] < TIE_EPS
    agree = (length == o_info["trajectory_length"]) & (te == o_te) & (tr == o_tr)
    assert (agree | tie).all() and (~agree).sum() <= 3
    fin = agree & np.isfinite(o_ret)
    rtol = 1e-5 if exact else 2e-5
    assert not fin.any() or (np.abs(ret[fin] - o_ret[fin]) <= rtol * np.maximum(1.0, np.abs(o_ret[fin]))).all()
    assert (np.abs(obs[agree] - o_obs[agree]) <= 2e-5 * np.maximum(1.0, np.abs(o_obs[agree]))).all()


def test_goal_offset_switch():
    """goal_offset (kwarg; unused by classic_control) with both placements, DMP and ProDMP"""
    import fancy_gym_b200 as fancy_gym
    B = 65
    rng = np.random.default_rng(1)
    for env_id in ("fancy_DMP/ViaPointReacher-v0", "fancy_ProDMP/SimpleReacher-v0"):
        for after in (True, False):
            traj = dict(RESOLVED[env_id]["traj"], goal_offset=0.3, goal_scale=1.5)
            with pa.assume(goal_offset_after_scale=after):
                env = fancy_gym.make(env_id, num_envs=B, device="cuda:0", mp_config_override={"trajectory_generator_kwargs": traj})
            with omp.assume(goal_offset_after_scale=after):
                orc = make_oracle(env_id, mode="mirror", mp_overrides={"traj": dict(goal_offset=0.3, goal_scale=1.5)})
            env.reset(seed=3)
            orc.reset(seeds=3 + np.arange(B))
            params = (0.5 * rng.standard_normal((B, env.action_space.shape[0]))).astype(np.float32)
            o_pos, o_vel = orc.get_trajectory(params)
            pos, vel = env.get_trajectory(torch.as_tensor(params, device="cuda:0"))
            assert np.abs(pos.cpu().numpy() - o_pos).max() <= 2e-6 * max(1.0, np.abs(o_pos).max()), (env_id, after)
            assert np.abs(vel.cpu().numpy() - o_vel).max() <= 3e-5 * max(1.0, np.abs(o_vel).max()), (env_id, after)
