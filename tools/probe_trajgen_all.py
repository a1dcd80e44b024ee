"""GPU probe: stand-alone trajectory generation throughput for every MP type (+ the per-env-phase kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym
dev = torch.device("cuda", 0)
B = 1 << 18
def run(env_id, label, phase=None):
    over = {"phase_generator_kwargs": phase} if phase else {}
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=over)
    env.reset(seed=0)
    tg = env.traj_gen
    P = env.action_space.shape[0]
    p = 0.3 * torch.randn(B, P, device=dev)
    if phase:
        p[:, 0] = 0.5 + 1.5 * torch.rand(B, device=dev)
    tg.set_params(p); tg.set_initial_conditions(0.0, env.unwrapped.q, env.unwrapped.v); tg.set_duration(2.0, 0.01)
    N, T = tg.num_dof, tg.n_steps
    outs = [(torch.empty(B, T, N, device=dev), torch.empty(B, T, N, device=dev)) for _ in range(2)]
    for i in range(3): tg._run_trajgen(out=outs[i % 2])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for i, (a, b) in enumerate(ev):
        a.record(); tg._run_trajgen(out=outs[i % 2]); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[5]
    nbytes = B * (2 * T * N * 4 + P * 4)
    print(f"{label:46s} [{T},{N}] x {B}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  {B / ms * 1e3:.3e} traj/s", flush=True)
    del env, outs
    torch.cuda.empty_cache()
run("fancy_ProMP/HoleReacher-v0", "ProMP (k_trajgen_closed)")
run("fancy_ProDMP/HoleReacher-v0", "ProDMP (k_trajgen_closed)")
run("fancy_DMP/ViaPointReacher-v0", "DMP (k_trajgen_dmp)")
run("fancy_ProDMP/SimpleReacher-v0", "ProDMP dof 2 (k_trajgen_closed)")
run("fancy_ProMP/HoleReacher-v0", "ProMP per-env tau (k_trajgen_phase)", dict(phase_generator_type="linear", learn_tau=True))
run("fancy_DMP/ViaPointReacher-v0", "DMP per-env tau (k_trajgen_phase)", dict(phase_generator_type="exp", alpha_phase=2, learn_tau=True))
run("fancy_ProDMP/HoleReacher-v0", "ProDMP per-env tau (k_trajgen_phase)", dict(learn_tau=True))
