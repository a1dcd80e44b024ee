"""GPU stress: plans inside one launch vs one launch per plan, many seeds / batch sizes (HoleReacher terminates early: re-packing
and plan switches interleave)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym
dev = torch.device("cuda", 0)
bad = 0
for env_id, sched, n_plans, sigma in (("fancy_ProMP/HoleReacher-v0", 50, 4, 0.5), ("fancy_DMP/HoleReacher-v0", 40, 5, 0.3),
                                      ("fancy_ProDMP/HoleReacher-v0", 60, 4, 0.5), ("fancy_ProMP/HoleReacher-v0", 20, 10, 0.7)):
    for B in (3001, 128 * 7, 20000):
        over = {"black_box_kwargs": {"replanning_schedule": (lambda k: (lambda p, v, o, a, t: t % k == 0))(sched)}}
        one = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over)
        seq = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over)
        for seed in range(6):
            one.reset(seed=seed); seq.reset(seed=seed)
            g = torch.Generator(device=dev).manual_seed(seed)
            acts = sigma * torch.randn(B, n_plans, one.action_space.shape[0], generator=g, device=dev)
            o = one.step_plans(acts)
            tot = 0
            for j in range(n_plans):
                s = seq.step(acts[:, j])
                ok = (torch.equal(o[4]["trajectory_length"][j], s[4]["trajectory_length"]) and torch.equal(o[0][j], s[0])
                      and torch.equal(torch.nan_to_num(o[1][j]), torch.nan_to_num(s[1])) and torch.equal(o[2][j], s[2]))
                if not ok:
                    bad += 1
                    nb = int((o[4]["trajectory_length"][j] != s[4]["trajectory_length"]).sum())
                    print("MISMATCH", env_id, B, seed, "plan", j, "lengths differing:", nb, flush=True)
                tot += int(s[4]["trajectory_length"].sum())
        print(env_id, sched, B, "ok" if not bad else "BAD", tot, flush=True)
print("stress done, mismatches:", bad)
