"""The switchable readings of mp_pytorch (SURVEY.md App. B.8 + the exponential phase's right clip): the oracle and the
CUDA path carry the SAME table of switches with the same defaults, and under every flipped switch the tables the CUDA
kernels consume (built on the host, no GPU needed) are bit for bit the oracle's 'mirror' tables.  The `-m gpu` half
(tests/test_gpu_assumptions.py) checks the kernels' trajectories and rollouts under the flipped switches."""
import numpy as np
import pytest
import torch

import fancy_gym_b200 as fancy_gym
from fancy_gym_b200.mp import assumptions as pa
from oracle import mp as omp
from oracle.blackbox import make_oracle

IDS = ["fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0", "fancy_ProDMP/SimpleReacher-v0",
       "fancy_DMP/HoleReacher-v0", "fancy_ProDMP/HoleReacher-v0"]
FLIPS = [{}, dict(scale_on_library_side=False), dict(alpha_phase_default=2.0), dict(centres_through_unbounded_phase=False),
         dict(exp_phase_right_clip=False), dict(prodmp_interpolate=True),
         dict(exp_phase_right_clip=False, scale_on_library_side=False)]
# a phase shorter than the episode makes the right clip visible for every exponential-phase MP
PHASE_OVER = {"fancy_DMP/ViaPointReacher-v0": dict(tau=1.2), "fancy_DMP/HoleReacher-v0": dict(tau=1.4, delay=0.1)}


def test_both_sides_carry_the_same_switches():
    assert pa.ASSUMPTIONS == omp.ASSUMPTIONS
    assert list(pa.ASSUMPTIONS) == list(omp.ASSUMPTIONS)
    with pytest.raises(KeyError):
        pa.assume(no_such_switch=1)
    with pytest.raises(KeyError):
        omp.assume(no_such_switch=1)
    with pa.assume(exp_phase_right_clip=False):
        assert pa.ASSUMPTIONS["exp_phase_right_clip"] is False
    assert pa.ASSUMPTIONS["exp_phase_right_clip"] is True


def _pair(env_id, flips, B=3):
    from oracle.blackbox import RESOLVED
    ph = PHASE_OVER.get(env_id, {})
    over = {"phase_generator_kwargs": dict(RESOLVED[env_id]["phase"], **ph)} if ph else {}
    with pa.assume(**flips):
        env = fancy_gym.make(env_id, num_envs=B, device="cpu", mp_config_override=over)
    with omp.assume(**flips):
        orc = make_oracle(env_id, mode="mirror", mp_overrides={"phase": ph})
    return env, orc


@pytest.mark.parametrize("flips", FLIPS, ids=["+".join(f"{k}={v}" for k, v in f.items()) or "defaults" for f in FLIPS])
@pytest.mark.parametrize("env_id", IDS)
def test_kernel_tables_equal_the_oracles_under_every_switch(env_id, flips):
    B = 3
    env, orc = _pair(env_id, flips, B)
    tg, otg = env.traj_gen, orc.traj_gen
    assert tg.phase_gn.assume == otg.phase_gn.assume
    P = env.action_space.shape[0]
    n = tg.num_dof
    for init_time in (0.0, 0.5):
        if init_time and "ProDMP" not in env_id:
            continue
        tg.set_params(torch.zeros(B, P))
        tg.set_initial_conditions(init_time, torch.zeros(B, n), torch.zeros(B, n))
        tg.set_duration(2.0, 0.01)
        otg.set_params(np.zeros((B, P), np.float32))
        otg.set_initial_conditions(np.array(init_time), np.zeros((B, n)), np.zeros((B, n)))
        otg.set_duration(2.0, 0.01)
        tb = tg.tables()
        t32 = otg.times.astype(np.float32)
        assert np.array_equal(tg.times32(), t32)
        if "ProMP" in env_id:
            assert np.array_equal(tb.tab_a, otg._scaled_basis_learnable())
            assert np.array_equal(tb.tab_b, np.diff(t32))
        elif "ProDMP" in env_id:
            pos_h, vel_h, xi = otg.tables()
            assert np.array_equal(tb.tab_a, np.concatenate([xi[..., 0:2], pos_h], axis=-1).astype(np.float32))
            assert np.array_equal(tb.tab_b, np.concatenate([xi[..., 2:4], vel_h], axis=-1).astype(np.float32))
        else:
            lin = otg.phase_gn.phase_argument(otg.times)
            scale = 1.0 if otg.phase_gn.assume["scale_on_library_side"] else otg.weights_scale
            xb = (omp._phase_from_linear_phase(otg.phase_gn, lin)[..., None]
                  * omp._basis_from_linear_phase(otg.basis_gn, lin) * scale).astype(np.float32)
            assert np.array_equal(tb.tab_a, xb)
            assert np.array_equal(tb.tab_b, np.diff(otg.phase_gn.left_bound_linear_phase(otg.times)).astype(np.float32))
            assert tb.weights_scale == (otg.weights_scale if otg.phase_gn.assume["scale_on_library_side"] else 1.0)


def test_flipped_switches_change_what_they_should():
    """the right clip only matters once the phase runs out before the plan does; the scale placement never changes a table
    by more than float32 rounding; the alpha default only reaches configs that give none (every fancy_ProDMP id)"""
    def tab(env_id, **flips):
        env, _ = _pair(env_id, flips)
        tg = env.traj_gen
        tg.set_params(torch.zeros(3, env.action_space.shape[0]))
        tg.set_initial_conditions(0.0, torch.zeros(3, tg.num_dof), torch.zeros(3, tg.num_dof))
        tg.set_duration(2.0, 0.01)
        return tg.tables().tab_a
    a, b = tab("fancy_ProDMP/SimpleReacher-v0"), tab("fancy_ProDMP/SimpleReacher-v0", exp_phase_right_clip=False)
    assert np.array_equal(a[:150], b[:150]) and not np.array_equal(a[150:], b[150:])      # tau = 1.5 s of a 2 s plan
    a, b = tab("fancy_DMP/ViaPointReacher-v0"), tab("fancy_DMP/ViaPointReacher-v0", exp_phase_right_clip=False)
    assert not np.array_equal(a, b)             # (tau overridden to 1.2 s here; the registered tau = 2.0 = duration is unaffected)
    a, b = tab("fancy_ProDMP/SimpleReacher-v0"), tab("fancy_ProDMP/SimpleReacher-v0", alpha_phase_default=2.0)
    assert not np.array_equal(a, b)
    a, b = tab("fancy_ProMP/HoleReacher-v0"), tab("fancy_ProMP/HoleReacher-v0", alpha_phase_default=2.0)
    assert np.array_equal(a, b)
    a, b = tab("fancy_ProMP/HoleReacher-v0"), tab("fancy_ProMP/HoleReacher-v0", scale_on_library_side=False)
    assert np.array_equal(a, b * np.float32(2.0))


def test_parameter_side_transforms():
    """scale / goal offset moved onto the parameters: what the kernels receive equals the oracle's float32 parameters"""
    B = 4
    rng = np.random.default_rng(0)
    for env_id, kw in (("fancy_ProMP/HoleReacher-v0", {}), ("fancy_ProDMP/SimpleReacher-v0", dict(weights_scale=0.7, goal_scale=1.3))):
        with pa.assume(scale_on_library_side=False):
            env = fancy_gym.make(env_id, num_envs=B, device="cpu", mp_config_override={"trajectory_generator_kwargs": kw} if kw else {})
        tg = env.traj_gen
        p = rng.standard_normal((B, env.action_space.shape[0])).astype(np.float32)
        tg.set_params(torch.as_tensor(p))
        got = tg.params.numpy().reshape(B, tg.num_dof, -1)
        if "ProMP" in env_id:
            want = (p.reshape(B, tg.num_dof, -1) * np.float32(2.0)).astype(np.float32)
        else:
            sc = np.array([0.7] * 5 + [1.3], np.float32)
            want = (p.reshape(B, tg.num_dof, -1) * sc).astype(np.float32)
        assert np.array_equal(got, want)
