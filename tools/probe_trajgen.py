"""GPU probe: per-launch time of fg_trajgen vs the write-only HBM ceiling (torch memset) and the host-side call cost."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym

dev = torch.device("cuda", 0)
Bt = 1 << 18
env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=4, device=dev)
tg = env.traj_gen
tp = (0.25 * torch.randn(Bt, 25, device=dev)).contiguous()
tg.set_params(tp); tg.set_initial_conditions(0.0, None, None); tg.set_duration(2.0, 0.01)
outs = [(torch.empty(Bt, 200, 5, device=dev), torch.empty(Bt, 200, 5, device=dev)) for _ in range(2)]
for i in range(3):
    tg._run_trajgen(out=outs[i % 2])
torch.cuda.synchronize()
n = 20
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
t0 = time.perf_counter()
for i in range(n):
    ev[i][0].record(); tg._run_trajgen(out=outs[i % 2]); ev[i][1].record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
byts = Bt * 8100
print(f"trajgen per-launch ms: min {ts[0]:.4f} med {ts[n//2]:.4f} max {ts[-1]:.4f}; GB/s at median {byts/ts[n//2]/1e6:.0f}; host loop {host/n*1e3:.3f} ms/call")
# write-only ceiling
for i in range(3):
    outs[0][0].zero_(); outs[0][1].zero_()
torch.cuda.synchronize()
for i in range(n):
    ev[i][0].record(); outs[i % 2][0].zero_(); outs[i % 2][1].zero_(); ev[i][1].record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
print(f"memset 2.1 GB ms: min {ts[0]:.4f} med {ts[n//2]:.4f}; GB/s {Bt*8000/ts[n//2]/1e6:.0f}")
# copy (read+write) ceiling as MEASURED_PEAKS does
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
for i in range(3): b.copy_(a)
torch.cuda.synchronize()
for i in range(10):
    ev[i][0].record(); b.copy_(a); ev[i][1].record()
torch.cuda.synchronize()
ts = sorted(x.elapsed_time(y) for x, y in ev[:10])
print(f"copy 2x2 GiB ms: min {ts[0]:.4f}; GB/s {2*a.numel()*2/ts[0]/1e6:.0f}")
