"""GPU probe: host cost of bench.py's device-timed loop, with and without the NVML clock sampler thread."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym
import bench

dev = torch.device("cuda", 0)
B = 65536
env = fancy_gym.make(bench.ENV_ID, num_envs=B, device=dev)
base = env.unwrapped
env.reset(seed=1)
s = dict(params=(0.25 * torch.randn(B, 25, device=dev)).contiguous(), q=base.q.clone(), ctx=base.ctx.clone())
total = torch.zeros((), dtype=torch.int64, device=dev)

def loop(K):
    for i in range(K):
        base.q.copy_(s["q"]); base.ctx.copy_(s["ctx"]); base.v.zero_(); base.steps.zero_(); base.done.zero_()
        env.launch(s["params"])
        total.add_(env._len.sum())

def timed(K=200):
    loop(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(K); th = time.perf_counter() - t0
    torch.cuda.synchronize(); ta = time.perf_counter() - t0
    return th / K * 1e3, ta / K * 1e3

print("no sampler      host %.3f total %.3f ms/iter" % timed())
clk = bench.ClockSampler(0); clk.__enter__()
time.sleep(0.2)
print("with sampler    host %.3f total %.3f ms/iter" % timed())
clk.__exit__()
print(clk.summary())
print("after sampler   host %.3f total %.3f ms/iter" % timed())
pr = cProfile.Profile(); pr.enable(); loop(200); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
