"""GPU, under compute-sanitizer: small instances of every scheduling path of the round-2 rollout kernel — re-packing (sigma = 1),
the work queue (more blocks than resident), plans inside one launch, the per-env phase evaluated inside the rollout, the
trajectory-from-HBM variant with ragged plans, and the per-step debug variant.
    compute-sanitizer --tool memcheck python tools/sanitize_r2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)


def run(env_id, B, sigma, over=None, plans=0, calls=1, **kw):
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over or {}, **kw)
    env.reset(seed=1)
    P = env.action_space.shape[0]
    tot = 0
    for _ in range(calls):
        if plans:
            out = env.step_plans(sigma * torch.randn(B, plans, P, generator=gen, device=dev))
        else:
            p = sigma * torch.randn(B, P, generator=gen, device=dev)
            if env.traj_gen.phase_gn.num_params:
                p[:, 0] = 0.1 + 0.8 * torch.rand(B, generator=gen, device=dev)
            out = env.step(p)
        tot += int(out[4]["trajectory_length"].sum())
    torch.cuda.synchronize()
    print(f"{env_id} B={B} sigma={sigma} plans={plans}: {tot} env steps", flush=True)


run("fancy_ProMP/HoleReacher-v0", 5000 + 13, 1.0)                                   # re-packing, ragged last block
run("fancy_ProMP/HoleReacher-v0", 160_000, 1.0)                                     # persistent grid + work queue
run("fancy_DMP/ViaPointReacher-v0", 3001, 1.0)
run("fancy_ProDMP/SimpleReacher-v0", 3001, 1.0,
    {"black_box_kwargs": {"replanning_schedule": lambda p, v, o, a, t: t % 25 == 0, "max_planning_times": 4, "condition_on_desired": True}}, plans=4)
run("fancy_ProMP/HoleReacher-v0", 3001, 0.5, {"black_box_kwargs": {"replanning_schedule": lambda p, v, o, a, t: t % 50 == 0}}, plans=4)
run("fancy_ProMP/HoleReacher-v0", 3001, 0.4, {"black_box_kwargs": {"learn_sub_trajectories": True}}, calls=3)      # fused per-env phase, ragged
run("fancy_DMP/ViaPointReacher-v0", 3001, 0.4, {"phase_generator_kwargs": {"phase_generator_type": "exp", "alpha_phase": 2, "learn_tau": True}})
os.environ["FG_PHASE_FUSED"] = "0"
run("fancy_ProMP/HoleReacher-v0", 3001, 0.4, {"black_box_kwargs": {"learn_sub_trajectories": True}}, calls=2)      # trajgen_phase + FG_MP_TRAJ
os.environ.pop("FG_PHASE_FUSED")
run("fancy_ProMP/HoleReacher-v0", 1001, 1.0, {"black_box_kwargs": {"verbose": 2}})                                 # per-step debug variant
print("sanitize_r2 done")
