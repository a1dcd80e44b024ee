"""CPU tests: the oracle restatement against the committed golden vectors that were produced from
the reference's own files (tests/golden/make_golden.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest

from oracle.blackbox import make_oracle
from oracle.reacher import (BatchedReacher, sample_hole_context, sample_simple_context,
                            sample_viapoint_context)

from tests.golden.make_golden import BB_CASES, ENV_CASES, bb_case, close64


@pytest.mark.parametrize("case", ENV_CASES, ids=[c[0] for c in ENV_CASES])
def test_env_restatement_matches_reference_files(case, golden_dir):
    key, name, kind, kw, seeds, amps, over = case
    kw = {**kw, **over}
    g = np.load(os.path.join(golden_dir, "env_kat.npz"))
    key = key.replace("-", "_")
    n = kw["n_links"]
    o = BatchedReacher(kind, **kw)
    ob0 = o.reset(seeds=seeds)
    assert np.array_equal(ob0, g[f"{key}/obs0"])
    length = g[f"{key}/length"]
    for t in range(200):
        a = np.stack([(A * np.sin(0.05 * t + np.arange(n))).astype(np.float32) for A in amps])
        ob, r, te, info = o.step(a)
        for i in range(len(seeds)):
            if t < length[i]:
                assert np.array_equal(ob[i], g[f"{key}/obs"][i, t])
                assert close64(r[i], g[f"{key}/rew"][i, t])
                assert te[i] == g[f"{key}/term"][i, t]


def test_context_samplers_match_survey_kats():
    # SURVEY.md App. C, captured from the reference's reset(seed)
    c = sample_hole_context(0)
    assert (c["x"], c["width"], c["q0"]) == (0.32223536589788804, 0.372936590562509, 0.8113597125762666)
    c = sample_hole_context(1)
    assert (c["x"], c["width"], c["q0"]) == (0.645403256627584, 0.3291375686450898, 2.2755332303766402)
    c = sample_viapoint_context(0)
    assert tuple(c["via"]) == (0.6848084366072715, -1.1510664311806484)
    assert tuple(c["goal"]) == (1.066357757671799, 2.294965609839984)
    c = sample_simple_context(0)
    assert tuple(c["goal"]) == (1.6510223091108869, 0.42654310306871945) and c["q0"] == 1.7859352421510681
    c = sample_simple_context(1)
    assert tuple(c["goal"]) == (1.7945977885489754, -0.7526741919580582) and c["q0"] == 1.5893656914508076


def test_survey_rollout_kats():
    # SURVEY.md App. C deterministic-action rollouts
    o = BatchedReacher("hole", n_links=5, random_start=True, hole_width=None, hole_depth=1, hole_x=None,
                       collision_penalty=100)
    o.reset(seeds=[0])
    tot = 0.0
    for t in range(200):
        a = (0.3 * np.sin(0.05 * t + np.arange(5))).astype(np.float32)[None]
        _, r, te, info = o.step(a)
        tot += r[0]
    assert abs(tot - (-31.268045763832532)) < 1e-12
    assert np.allclose(o.end_effector[0], (3.2836263007409707, 3.743211747110169), rtol=0, atol=1e-14)
    o = BatchedReacher("simple", n_links=2)
    o.reset(seeds=[0])
    tot = 0.0
    for t in range(200):
        a = (5.0 * np.sin(0.05 * t + np.arange(2))).astype(np.float32)[None]
        _, r, te, info = o.step(a)
        tot += r[0]
    # (the survey's sum for this case was accumulated in float32; per-step values are pinned by env_kat.npz)
    assert abs(tot - (-4995.885776395183)) < 5e-3
    assert abs(info["reward_dist"][0] - (-2.582664365153849)) < 1e-13


@pytest.mark.parametrize("case", BB_CASES, ids=[c[0] for c in BB_CASES])
def test_blackbox_loop_matches_reference_wrapper(case, golden_dir):
    fname, env_id, seeds, bbk, env_over, mp_over = bb_case(case)
    g = np.load(os.path.join(golden_dir, fname + ".npz"))
    orc = make_oracle(env_id, mode="shipped", verbose=2, mp_overrides=dict(mp_over, env=env_over), **bbk)
    ob0 = orc.reset(seeds=seeds)
    assert np.array_equal(ob0, g["obs0"])
    n_plans = g["params"].shape[1]
    for i in range(n_plans):
        ob, ret, te, tr, info = orc.step(g["params"][:, i])
        for b in range(len(seeds)):
            if i >= g["n_calls"][b]:
                continue
            L = g["length"][b, i]
            assert info["trajectory_length"][b] == L
            n = g["n_points"][b, i] if "n_points" in g.files else len(g["positions"][b, i])   # sub-trajectories: ragged plans
            assert np.array_equal(info["positions"][b][:n], g["positions"][b, i][:n])
            assert np.array_equal(info["velocities"][b][:n], g["velocities"][b, i][:n])
            assert np.array_equal(info["step_observations"][b, :L], g["step_obs"][b, i, :L])
            assert close64(info["step_rewards"][b, :L], g["step_rewards"][b, i, :L])
            assert close64(ret[b], g["ret"][b, i])
            assert te[b] == g["terminated"][b, i] and tr[b] == g["truncated"][b, i]
            assert np.array_equal(ob[b], g["obs"][b, i])


@pytest.mark.parametrize("env_id", ["fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0", "fancy_ProDMP/SimpleReacher-v0"])
def test_ragged_sub_trajectories_equal_the_scalar_loop(env_id):
    """learn_sub_trajectories with a DIFFERENT learned tau per env of a batch: the batched oracle (per-env time grids,
    per-env plan lengths) must be what the reference does env by env (black_box_wrapper.py:96-120,150-217 on one env each,
    here the oracle at B = 1, which the golden black-box cases pin)."""
    B = 6
    orc = make_oracle(env_id, mode="mirror", learn_sub_trajectories=True)
    orc.reset(seeds=3 + np.arange(B))
    singles = []
    for b in range(B):
        o = make_oracle(env_id, mode="mirror", learn_sub_trajectories=True)
        o.reset(seeds=[3 + b])
        singles.append(o)
    rng = np.random.default_rng(0)
    P = orc.traj_gen.num_params
    total = np.zeros(B, dtype=np.int64)
    for k in range(4):
        params = (0.4 * rng.standard_normal((B, P))).astype(np.float32)
        params[:, 0] = rng.uniform(0.1, 1.2, size=B).astype(np.float32)
        obs, ret, te, tr, info = orc.step(params)
        for b in range(B):
            o_obs, o_ret, o_te, o_tr, o_info = singles[b].step(params[b:b + 1])
            assert info["trajectory_length"][b] == o_info["trajectory_length"][0]
            assert np.array_equal(obs[b], o_obs[0]) and np.array_equal(ret[b:b + 1], o_ret, equal_nan=True)
            assert te[b] == o_te[0] and tr[b] == o_tr[0]
        total += info["trajectory_length"]
    assert (total <= 200).all()
