"""GPU: launches ONE named workload a few times so that `ncu -k regex:<kernel> -s <skip> -c 1` can capture its kernel.
    python tools/run_kernel.py <workload>
workloads: rollout_hole_prodmp rollout_hole_dmp rollout_sigma1 rollout_learned_tau rollout_1m_sigma1 rollout_1m rollout_config3 rollout_config4_plans trajgen_promp trajgen_prodmp trajgen_dmp
           trajgen_phase_promp trajgen_phase_dmp reset cov"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

dev = torch.device("cuda", 0)
what = sys.argv[1]
gen = torch.Generator(device=dev).manual_seed(0)
REPS = 6


def rollout(env_id, B, sigma, over=None):
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over or {})
    P = env.action_space.shape[0]
    p = sigma * torch.randn(B, P, generator=gen, device=dev)
    for i in range(REPS):
        env.reset(seed=i)
        env.step(p)
    torch.cuda.synchronize()


def trajgen(env_id, phase=None, B=1 << 18):
    over = {"phase_generator_kwargs": phase} if phase else {}
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over)
    env.reset(seed=0)
    tg = env.traj_gen
    P = env.action_space.shape[0]
    p = 0.3 * torch.randn(B, P, generator=gen, device=dev)
    if phase:
        p[:, 0] = 0.5 + 1.5 * torch.rand(B, generator=gen, device=dev)
    tg.set_params(p); tg.set_initial_conditions(0.0, env.unwrapped.q, env.unwrapped.v); tg.set_duration(2.0, 0.01)
    outs = [(torch.empty(B, tg.n_steps, tg.num_dof, device=dev), torch.empty(B, tg.n_steps, tg.num_dof, device=dev)) for _ in range(2)]
    for i in range(REPS):
        tg._run_trajgen(out=outs[i % 2])
    torch.cuda.synchronize()


if what == "rollout_sigma1":
    rollout("fancy_ProMP/HoleReacher-v0", 65536, 1.0)
elif what == "rollout_learned_tau":
    over = {"phase_generator_kwargs": {"phase_generator_type": "linear", "learn_tau": True, "learn_delay": True}}
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=65536, device=dev, mp_config_override=over)
    p = 0.25 * torch.randn(65536, 27, generator=gen, device=dev)
    p[:, 0] = 0.5 + 1.5 * torch.rand(65536, generator=gen, device=dev)
    p[:, 1] = 0.3 * torch.rand(65536, generator=gen, device=dev)
    for i in range(REPS):
        env.reset(seed=i)
        env.step(p)
    torch.cuda.synchronize()
elif what == "rollout_1m_sigma1":
    rollout("fancy_ProMP/HoleReacher-v0", 1 << 20, 1.0)
elif what == "rollout_1m":
    rollout("fancy_ProMP/HoleReacher-v0", 1 << 20, 0.25)
elif what == "rollout_hole_prodmp":
    rollout("fancy_ProDMP/HoleReacher-v0", 65536, 0.25)
elif what == "rollout_hole_dmp":
    rollout("fancy_DMP/HoleReacher-v0", 65536, 0.25)
elif what == "rollout_config3":
    rollout("fancy_DMP/ViaPointReacher-v0", 1 << 18, 1.0)
elif what == "rollout_config4_plans":
    over = {"black_box_kwargs": {"replanning_schedule": lambda p, v, o, a, t: t % 25 == 0, "max_planning_times": 4, "condition_on_desired": True}}
    env = fancy_gym.make("fancy_ProDMP/SimpleReacher-v0", num_envs=65536, device=dev, mp_config_override=over)
    acts = torch.randn(65536, 4, 12, generator=gen, device=dev)
    for i in range(REPS):
        env.reset(seed=i)
        env.step_plans(acts)
    torch.cuda.synchronize()
elif what == "trajgen_promp":
    trajgen("fancy_ProMP/HoleReacher-v0")
elif what == "trajgen_prodmp":
    trajgen("fancy_ProDMP/HoleReacher-v0")
elif what == "trajgen_dmp":
    trajgen("fancy_DMP/ViaPointReacher-v0")
elif what == "trajgen_phase_promp":
    trajgen("fancy_ProMP/HoleReacher-v0", dict(phase_generator_type="linear", learn_tau=True))
elif what == "trajgen_phase_dmp":
    trajgen("fancy_DMP/ViaPointReacher-v0", dict(phase_generator_type="exp", alpha_phase=2, learn_tau=True))
elif what == "reset":
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=1 << 20, device=dev)
    for i in range(REPS):
        env.reset(seed=i)
    torch.cuda.synchronize()
else:
    raise SystemExit(f"unknown workload {what}")
