"""GPU probe: stand-alone DMP trajectory generation (k_trajgen_dmp) throughput."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym
dev = torch.device("cuda", 0)
B = 1 << 18
env = fancy_gym.make("fancy_DMP/ViaPointReacher-v0", num_envs=B, device=dev)
env.reset(seed=0)
tg = env.traj_gen
P = env.action_space.shape[0]
p = 0.3 * torch.randn(B, P, device=dev)
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
print(f"DMP [{T},{N}] x {B}: {ms:.3f} ms  {B * (2 * T * N * 4 + P * 4) / ms / 1e6:.0f} GB/s  checksum {float(outs[0][0].double().sum()):.6f}")
