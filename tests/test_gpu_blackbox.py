"""The reference's structural black-box tests (test/test_black_box.py, test/test_replanning_sequencing.py), run
against the CUDA path with num_envs=1 and numpy actions (the scalar contract), on the batched twin of the tests' ToyEnv.
These are the only tests the reference holds at the mp_pytorch boundary: trajectory lengths, tau /
delay plateaus with exact equality, plan counts, reward aggregation, info keys."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 1
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fg():
    import fancy_gym_b200 as fancy_gym
    from tests.toy import register_toy
    register_toy(fancy_gym)
    return fancy_gym


def _types(mp_type):
    return ("prodmp" if mp_type == "prodmp" else "rbf"), ("exp" if mp_type == "prodmp" else "linear")


def _toy(fg, mp_type, bb_kwargs=None, phase_kwargs=None, phase_type=None, **kw):
    from tests.toy import ToyWrapper
    basis, phase = _types(mp_type)
    return fg.make_bb("toy-v0", [ToyWrapper], dict(bb_kwargs or {}), {"trajectory_generator_type": mp_type},
                      {"controller_type": "motor"}, {"phase_generator_type": phase_type or phase, **(phase_kwargs or {})},
                      {"basis_generator_type": basis}, device=DEV, **kw)


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
def test_missing_local_state(fg, mp_type):
    """test_black_box.py:72-84: a bare RawInterfaceWrapper has no current_pos"""
    from fancy_gym_b200.black_box.raw_interface_wrapper import RawInterfaceWrapper
    basis, _ = _types(mp_type)
    env = fg.make_bb("toy-v0", [RawInterfaceWrapper], {}, {"trajectory_generator_type": mp_type}, {"controller_type": "motor"},
                     {"phase_generator_type": "exp"}, {"basis_generator_type": basis}, device=DEV)
    env.reset(seed=SEED)
    with pytest.raises(NotImplementedError):
        env.step(env.action_space.sample())


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
@pytest.mark.parametrize("verbose", [1, 2])
def test_verbosity(fg, mp_type, verbose):
    env = _toy(fg, mp_type, {"verbose": verbose}, phase_type="exp")
    env.reset(seed=SEED)
    _obs, _r, _te, _tr, info = env.step(env.action_space.sample())
    assert "trajectory_length" in info
    if verbose >= 2:
        for k in ("positions", "velocities", "step_actions", "step_observations", "step_rewards"):
            assert k in info, k
        L = info["trajectory_length"]
        assert info["positions"].shape == (50, 1) and info["step_actions"].shape == (L, 1)
        assert info["step_rewards"].shape == (L,) and info["step_observations"].shape[0] == L
        assert np.all(info["step_rewards"] == 1.0) and np.all(info["step_observations"] == -1.0)


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
def test_length(fg, mp_type):
    env = _toy(fg, mp_type, phase_type="exp")
    for _ in range(5):
        env.reset(seed=SEED)
        _obs, _r, te, tr, info = env.step(env.action_space.sample())
        assert info["trajectory_length"] == env.spec.max_episode_steps == 50
        assert tr and not te


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
@pytest.mark.parametrize("agg", [np.sum, np.mean, np.median, lambda x: np.mean(x[::2])], ids=["sum", "mean", "median", "lambda"])
def test_aggregation(fg, mp_type, agg):
    env = _toy(fg, mp_type, {"reward_aggregation": agg}, phase_type="exp")
    env.reset(seed=SEED)
    _obs, reward, _te, _tr, _info = env.step(env.action_space.sample())
    assert reward == agg(np.ones(50))


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
def test_change_env_kwargs(fg, mp_type):
    c, d, e = [np.ones(3)], {"a": {"a": "b"}}, object()
    env = _toy(fg, mp_type, phase_type="exp", a=1, b=1.0, c=c, d=d, e=e)
    assert env.a == 1 and env.b == 1.0 and d == env.d


@pytest.mark.parametrize("mp_type", ["promp", "prodmp"])
@pytest.mark.parametrize("tau", [0.25, 0.5, 0.75, 1])
def test_learn_tau(fg, mp_type, tau):
    _, phase = _types(mp_type)
    env = _toy(fg, mp_type, {"verbose": 2}, {"learn_tau": True, "learn_delay": False})
    env.reset(seed=SEED)
    done = True
    for _ in range(5):
        if done:
            env.reset(seed=SEED)
        action = env.action_space.sample()
        action[0] = tau
        _obs, _r, te, tr, info = env.step(action)
        done = te or tr
        assert info["trajectory_length"] == env.spec.max_episode_steps
        n = int(np.round(tau / env.dt))
        pos, vel = info["positions"].flatten(), info["velocities"].flatten()
        if phase == "linear":
            assert np.all(pos[n:] == pos[-1]) and np.all(vel[n:] == vel[-1])
        assert np.all(pos[:n - 1] != pos[-1]) and np.all(vel[:n - 2] != vel[-1])


@pytest.mark.parametrize("mp_type", ["promp", "prodmp"])
@pytest.mark.parametrize("delay", [0, 0.25, 0.5, 0.75])
def test_learn_delay(fg, mp_type, delay):
    env = _toy(fg, mp_type, {"verbose": 2}, {"learn_tau": False, "learn_delay": True})
    env.reset(seed=SEED)
    done = True
    for _ in range(5):
        if done:
            env.reset(seed=SEED)
        action = env.action_space.sample()
        action[0] = delay
        _obs, _r, te, tr, info = env.step(action)
        done = te or tr
        assert info["trajectory_length"] == env.spec.max_episode_steps
        n = int(np.round(delay / env.dt))
        pos, vel = info["positions"].flatten(), info["velocities"].flatten()
        assert np.all(pos[:max(1, n - 1)] == pos[0]) and np.all(vel[:max(1, n - 2)] == vel[0])
        assert np.all(pos[max(1, n):] != pos[0]) and np.all(vel[max(1, n)] != vel[0])


@pytest.mark.parametrize("mp_type", ["promp", "prodmp"])
@pytest.mark.parametrize("tau", [0.25, 0.5, 0.75, 1])
@pytest.mark.parametrize("delay", [0.25, 0.5, 0.75, 1])
def test_learn_tau_and_delay(fg, mp_type, tau, delay):
    _, phase = _types(mp_type)
    env = _toy(fg, mp_type, {"verbose": 2}, {"learn_tau": True, "learn_delay": True})
    env.reset(seed=SEED)
    if env.spec.max_episode_steps * env.dt < delay + tau:
        return
    done = True
    for _ in range(5):
        if done:
            env.reset(seed=SEED)
        action = env.action_space.sample()
        action[0], action[1] = tau, delay
        _obs, _r, te, tr, info = env.step(action)
        done = te or tr
        assert info["trajectory_length"] == env.spec.max_episode_steps
        nt, nd = int(np.round(tau / env.dt)), int(np.round(delay / env.dt))
        nj = nt + nd
        pos, vel = info["positions"].flatten(), info["velocities"].flatten()
        if phase == "linear":
            assert np.all(pos[nj:] == pos[-1]) and np.all(vel[nj:] == vel[-1])
        assert np.all(pos[:nd - 1] == pos[0]) and np.all(vel[:nd - 2] == vel[0])
        ap, av = pos[nd:nj - 1], vel[nd:nj - 2]
        assert np.all(ap != pos[-1]) and np.all(ap != pos[0])
        assert np.all(av != vel[-1]) and np.all(av != vel[0])


# ---- test/test_replanning_sequencing.py ------------------------------------------------------------------------
@pytest.mark.parametrize("mp_type", ["promp", "dmp"])
def test_learn_sub_trajectories(fg, mp_type):
    env = _toy(fg, mp_type, {"learn_sub_trajectories": True, "verbose": 2}, phase_type="exp")
    env.reset(seed=SEED)
    assert env.learn_sub_trajectories and env.traj_gen.learn_tau and env.observation_space.shape == (2,)
    done = True
    for _ in range(25):
        if done:
            env.reset(seed=SEED)
        action = env.action_space.sample()
        _obs, _r, te, tr, info = env.step(action)
        done = te or tr
        length = info["trajectory_length"]
        tau = float(np.asarray(env.traj_gen.tau).reshape(-1)[0])
        if not done:
            assert length == np.round(np.clip(action[0], *env.tau_bound) / env.dt) == np.round(tau / env.dt)
        else:
            assert length <= np.round(tau / env.dt)


@pytest.mark.parametrize("mp_type", ["promp", "dmp", "prodmp"])
@pytest.mark.parametrize("replanning_time", [10, 100, 1000])
def test_replanning_time(fg, mp_type, replanning_time):
    def schedule(c_pos, c_vel, obs, c_action, t):
        return t % replanning_time == 0
    phase = "exp" if "dmp" in mp_type else "linear"
    env = _toy(fg, mp_type, {"replanning_schedule": schedule, "verbose": 2}, phase_type=phase)
    env.reset(seed=SEED)
    assert env.do_replanning and callable(env.replanning_schedule) and env.observation_space.shape == (2,)
    episode_steps = max(env.spec.max_episode_steps // replanning_time, 1)
    for i in range(3 * episode_steps):
        _obs, _r, te, tr, info = env.step(env.action_space.sample())
        length = info["trajectory_length"]
        if te or tr:
            assert (i + 1) % episode_steps == 0
            env.reset(seed=SEED)
        if replanning_time <= env.spec.max_episode_steps:
            assert schedule(None, None, None, None, length)


@pytest.mark.parametrize("mp_type", ["promp", "prodmp"])
@pytest.mark.parametrize("max_planning_times", [1, 2, 3, 4])
@pytest.mark.parametrize("sub_segment_steps", [5, 10])
@pytest.mark.parametrize("phase_kwargs", [{}, {"learn_tau": True}, {"learn_delay": True}, {"learn_tau": True, "learn_delay": True}],
                         ids=["fixed", "tau", "delay", "tau_delay"])
def test_max_planning_times(fg, mp_type, max_planning_times, sub_segment_steps, phase_kwargs):
    env = _toy(fg, mp_type, {"max_planning_times": max_planning_times, "verbose": 2,
                             "replanning_schedule": lambda pos, vel, obs, action, t: t % sub_segment_steps == 0}, phase_kwargs)
    env.reset(seed=SEED)
    done, planning_times = False, 0
    delay = 0.25
    while not done:
        action = env.action_space.sample()
        i = 0
        if phase_kwargs.get("learn_tau"):
            action[i] = 1.0; i += 1
        if phase_kwargs.get("learn_delay"):
            action[i] = delay
        _obs, _r, te, tr, info = env.step(action)
        done = te or tr
        if phase_kwargs.get("learn_delay") and planning_times == 0:     # the delay only shapes the first plan
            n = int(np.round(delay / env.dt))
            pos, vel = info["positions"].flatten(), info["velocities"].flatten()
            assert np.all(pos[:max(1, n - 1)] == pos[0]) and np.all(vel[:max(1, n - 2)] == vel[0])
            assert np.all(pos[max(1, n):] != pos[0])
        planning_times += 1
        assert planning_times <= 50
    assert planning_times == max_planning_times


# ---- vector-env adapter ----------------------------------------------------------------------
def test_vector_env_autoreset_continues_each_env_stream(fg):
    """BlackBoxVectorEnv.step auto-resets every finished sub-env; sub-env i then is the reference env after
    reset(seed=s+i) followed by an unseeded reset() — checked against the host-side numpy sampler."""
    import torch
    N = 257
    vec = fg.make_vec("fancy_ProMP/HoleReacher-v0", N, device=DEV)
    ref = fg.make("fancy_ProMP/HoleReacher-v0", num_envs=N, device=DEV, context_sampler="numpy")
    obs, _ = vec.reset(seed=3)
    r_obs, _ = ref.reset(seed=3, options={"as_numpy": True})
    assert obs.shape == (N, 18) and np.allclose(obs, r_obs, atol=1e-6)
    rng = np.random.default_rng(0)
    for it in range(3):
        act = (0.5 * rng.standard_normal((N, 25))).astype(np.float32)
        obs, rew, term, trunc, info = vec.step(act)
        _, r_rew, r_term, r_trunc, r_info = ref.step(act)
        assert np.array_equal(term, r_term) and np.array_equal(trunc, r_trunc) and np.allclose(rew, r_rew, rtol=1e-12)
        assert np.array_equal(info["trajectory_length"], r_info["trajectory_length"])
        assert (term | trunc).all() and info["_final_observation"].all()
        r_obs, _ = ref.reset(seed=None, options={"as_numpy": True})       # the reference's unseeded follow-up reset
        assert np.allclose(obs, r_obs, atol=1e-6), it
        assert torch.equal(vec.env.unwrapped.ctx, ref.unwrapped.ctx)
    vec.close()


def test_partial_reset_only_touches_masked_envs(fg):
    import torch
    env = fg.make("fancy_DMP/ViaPointReacher-v0", num_envs=64, device=DEV)
    env.reset(seed=0)
    base = env.unwrapped
    ctx0, q0 = base.ctx.clone(), base.q.clone()
    mask = torch.zeros(64, dtype=torch.uint8, device=DEV)
    mask[::3] = 1
    env.reset_done(mask)
    keep = mask == 0
    assert torch.equal(base.ctx[keep], ctx0[keep]) and torch.equal(base.q[keep], q0[keep])
    assert not torch.equal(base.ctx[~keep], ctx0[~keep])


def test_graphed_episode_replays_the_eager_episode(fg):
    """GraphedEpisode (reset + H2D + rollout + D2H as one CUDA graph) == the same calls made eagerly, episode after episode"""
    import torch
    B = 2048
    eager = fg.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=DEV)
    graphed = fg.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=DEV)
    eager.reset(seed=11)
    graphed.reset(seed=11)
    runner = fg.GraphedEpisode(graphed, warmup=2, copy_obs=True)          # the 2 warm-up episodes advance every env's context stream
    for _ in range(2):                                      # (the capture itself only records, it does not execute)
        eager.reset(seed=None)
    gen = torch.Generator().manual_seed(0)
    for it in range(3):
        p = 0.5 * torch.randn(B, 25, generator=gen)
        runner.host_params.copy_(p)
        ret, length, term = runner.run()
        obs0, _ = eager.reset(seed=None)
        _, e_ret, e_term, _, e_info = eager.step(p.to(DEV))
        assert torch.equal(runner.host_obs, obs0.cpu())
        assert torch.equal(ret, e_ret.cpu()) and torch.equal(length, e_info["trajectory_length"].cpu())
        assert torch.equal(term, e_term.cpu())


@pytest.mark.parametrize("slots,graphs", [(2, False), (3, False), (4, False), (2, True), (3, True)])
def test_episode_pipeline_equals_sequential_steps(fg, slots, graphs):
    """EpisodePipeline (2 - 4 batches in flight, H2D / rollout / D2H on their own streams) returns, batch for batch, what
    sequential reset() / step() calls return — including each env's context stream advancing once per batch"""
    import torch
    B = 4096
    seq = fg.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=DEV)
    piped = fg.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=DEV, mp_config_override={"black_box_kwargs": {"result_sets": slots}})
    seq.reset(seed=3)
    piped.reset(seed=3)
    if slots > 2:
        with pytest.raises(ValueError):
            fg.EpisodePipeline(seq, slots=slots)          # two result sets only
    pipe = fg.EpisodePipeline(piped, slots=slots, graphs=graphs)
    gen = torch.Generator().manual_seed(1)
    pops = [0.5 * torch.randn(B, 25, generator=gen) for _ in range(7)]
    want = []
    for p in pops:
        seq.reset(seed=None)
        _, ret, term, _, info = seq.step(p.to(DEV))
        want.append((ret.cpu(), info["trajectory_length"].cpu(), term.cpu()))
    got = [None] * len(pops)
    for i, p in enumerate(pops):
        s = i % pipe.SLOTS
        if i >= pipe.SLOTS:
            got[i - pipe.SLOTS] = tuple(x.clone() for x in pipe.wait(s))
        pipe.host_params[s].copy_(p)
        pipe.submit(s)
    for i in range(len(pops) - pipe.SLOTS, len(pops)):
        got[i] = tuple(x.clone() for x in pipe.wait(i % pipe.SLOTS))
    for w, g in zip(want, got):
        assert all(torch.equal(a, b) for a, b in zip(w, g))
    assert len({float(w[0].sum()) for w in want}) == len(pops)          # the batches differ: nothing was compared with itself
    with pytest.raises(ValueError):
        pipe.submit((pipe.next_slot + 1) % pipe.SLOTS)
    with pytest.raises(RuntimeError):
        pipe.wait(0)
