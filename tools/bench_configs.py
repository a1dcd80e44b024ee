"""GPU: device-timed throughput of the other BASELINE.json configurations (parity-test cases, not the bench line):
config 3 fancy_DMP/ViaPointReacher-v0 x 262,144, config 4 fancy_ProDMP/SimpleReacher-v0 with replanning x 65,536,
config 5 fancy_ProMP/HoleReacher-v0 x 1,048,576 per GPU; plus HoleReacher at sigma = 1.0 (many early terminations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fancy_gym_b200 as fancy_gym

dev = torch.device("cuda", 0)

def timed_steps(env, params_list, n_rep=10, seed=0):
    """one episode = reset + len(params_list) black-box steps; returns (ms per episode batch, env steps per batch)"""
    def episode():
        env.reset(seed=None)
        steps = 0
        for p in params_list:
            _, _, _, _, info = env.step(p)
            steps = steps + info["trajectory_length"].sum()
        return steps
    env.reset(seed=seed)
    for _ in range(3): s = episode()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tot = 0
    for _ in range(n_rep): tot = tot + episode()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n_rep, int(tot) / n_rep

out = {}
gen = torch.Generator(device=dev).manual_seed(0)
for name, env_id, B, P, sigma, kw, plans in (
        ("config2 HoleReacher/ProMP 65536 sigma 0.25", "fancy_ProMP/HoleReacher-v0", 65536, 25, 0.25, {}, 1),
        ("config2 HoleReacher/ProMP 65536 sigma 1.0", "fancy_ProMP/HoleReacher-v0", 65536, 25, 1.0, {}, 1),
        ("config3 ViaPointReacher/DMP 262144", "fancy_DMP/ViaPointReacher-v0", 262144, 30, 1.0, {}, 1),
        ("config4 SimpleReacher/ProDMP 65536 replanning t%25, 4 plans, condition_on_desired", "fancy_ProDMP/SimpleReacher-v0", 65536, 12, 1.0,
         {"black_box_kwargs": {"replanning_schedule": lambda p, v, o, a, t: t % 25 == 0, "max_planning_times": 4, "condition_on_desired": True}}, 4),
        ("config5 HoleReacher/ProMP 1048576", "fancy_ProMP/HoleReacher-v0", 1 << 20, 25, 0.25, {}, 1)):
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=kw)
    params = [(sigma * torch.randn(B, P, generator=gen, device=dev)).contiguous() for _ in range(plans)]
    ms, steps = timed_steps(env, params)
    out[name] = dict(ms_per_episode_batch=ms, episodes_per_s=B / ms * 1e3, env_steps_per_s=steps / ms * 1e3, mean_len=steps / B)
    print(f"{name}: {ms:.3f} ms  {B / ms * 1e3:.3e} episodes/s  {steps / ms * 1e3:.3e} env-steps/s  (mean episode length {steps / B:.1f})", flush=True)
    del env, params
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bench_configs.json"), "w"), indent=1)
