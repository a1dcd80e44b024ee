"""GPU: device time of the black-box step() alone (CUDA events around it, reset outside) for a list of workloads.
    python tools/time_rollouts.py [name ...]        FG_LIB_PATH=<variant build> for A/B comparisons
Prints per workload: median / min ms over REPS launches, env steps per launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

dev = torch.device("cuda", 0)
REPS = int(os.environ.get("REPS", "15"))
W = {
    "config2": ("fancy_ProMP/HoleReacher-v0", 65536, 0.25, {}),
    "sigma1": ("fancy_ProMP/HoleReacher-v0", 65536, 1.0, {}),
    "sigma1_256k": ("fancy_ProMP/HoleReacher-v0", 1 << 18, 1.0, {}),
    "sigma1_1m": ("fancy_ProMP/HoleReacher-v0", 1 << 20, 1.0, {}),
    "config3": ("fancy_DMP/ViaPointReacher-v0", 1 << 18, 1.0, {}),
    "config3_s025": ("fancy_DMP/ViaPointReacher-v0", 1 << 18, 0.25, {}),
    "viapoint_64k": ("fancy_DMP/ViaPointReacher-v0", 65536, 1.0, {}),
    "viapoint_promp_64k": ("fancy_ProMP/ViaPointReacher-v0", 65536, 1.0, {}),
    "hole_dmp_64k": ("fancy_DMP/HoleReacher-v0", 65536, 0.25, {}),
    "hole_prodmp_64k": ("fancy_ProDMP/HoleReacher-v0", 65536, 0.25, {}),
    "config5_1gpu": ("fancy_ProMP/HoleReacher-v0", 1 << 20, 0.25, {}),
    "learned_tau": ("fancy_ProMP/HoleReacher-v0", 65536, 0.25,
                    {"phase_generator_kwargs": {"phase_generator_type": "linear", "learn_tau": True, "learn_delay": True}}),
}
names = sys.argv[1:] or list(W)
for name in names:
    env_id, B, sigma, over = W[name]
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over)
    gen = torch.Generator(device=dev).manual_seed(0)
    P = env.action_space.shape[0]
    p = sigma * torch.randn(B, P, generator=gen, device=dev)
    if name == "learned_tau":
        p[:, 0] = 0.5 + 1.5 * torch.rand(B, generator=gen, device=dev)
        p[:, 1] = 0.3 * torch.rand(B, generator=gen, device=dev)
    ts = []
    steps = 0
    for i in range(REPS + 3):
        env.reset(seed=i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = env.step(p)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
            steps = int(out[4]["trajectory_length"].sum())
    ts.sort()
    print(f"{name:20s} median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f} ms  {steps} env steps  {steps / ts[len(ts) // 2] * 1e3:.3e} env-steps/s", flush=True)
    env.close()
