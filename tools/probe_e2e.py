"""GPU probe: where does the end-to-end black-box step spend its time (host vs device)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym

dev = torch.device("cuda", 0)
B = 65536
env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev, context_sampler="device")
hp = (0.25 * torch.randn(B, 25)).pin_memory()
host_ret = torch.empty(B, dtype=torch.float64).pin_memory()
host_len = torch.empty(B, dtype=torch.int32).pin_memory()
host_flags = torch.empty(B, dtype=torch.bool).pin_memory()

def sync():
    torch.cuda.synchronize(dev)

def timed(f, n=30):
    for _ in range(3): f()
    sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    t_host = (time.perf_counter() - t0) / n
    sync(); t_all = (time.perf_counter() - t0) / n
    return t_host * 1e3, t_all * 1e3

env.reset(seed=0)
p_dev = hp.to(dev)
print("reset(seed=None)       host %.3f ms  total %.3f ms" % timed(lambda: env.reset(seed=None)))
print("H2D params             host %.3f ms  total %.3f ms" % timed(lambda: hp.to(dev, non_blocking=True)))
def step_only():
    env.reset(seed=None); env.step(p_dev)
print("reset+step(dev params) host %.3f ms  total %.3f ms" % timed(step_only))
def launch_only():
    env.unwrapped.steps.zero_(); env.unwrapped.done.zero_(); env.launch(p_dev)
print("zero+launch            host %.3f ms  total %.3f ms" % timed(launch_only))
def d2h():
    host_ret.copy_(env._ret, non_blocking=True); host_len.copy_(env._len, non_blocking=True); host_flags.copy_(env._flags != 0, non_blocking=True)
print("D2H results            host %.3f ms  total %.3f ms" % timed(d2h))
def full():
    env.reset(seed=None)
    p = hp.to(dev, non_blocking=True)
    obs, ret, te, tr, info = env.step(p)
    host_ret.copy_(ret, non_blocking=True); host_len.copy_(info["trajectory_length"], non_blocking=True); host_flags.copy_(te, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
print("full e2e step          host %.3f ms  total %.3f ms" % timed(full))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): full()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
