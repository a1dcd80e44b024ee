"""GPU probe: throughput of fg_traj_cov, CUDA-core path vs tcgen05 path (BASELINE config 4 shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import fancy_gym_b200 as fancy_gym
from fancy_gym_b200 import _lib
dev = torch.device("cuda", 0)
for env_id, B in (("fancy_ProMP/HoleReacher-v0", 512), ("fancy_ProDMP/SimpleReacher-v0", 4096)):
    env = fancy_gym.make(env_id, num_envs=4, device=dev); env.reset(seed=0)
    tg = env.traj_gen
    tg.set_initial_conditions(0.0, env.unwrapped.q, env.unwrapped.v); tg.set_duration(2.0, 0.01)
    D, T, N = tg._num_local_params, tg.n_steps, tg.num_dof
    L = (torch.tril(0.1 * torch.randn(B, D, D, device=dev)) + 0.5 * torch.eye(D, device=dev)).contiguous()
    h = tg._trajgen_handle()
    covs = [torch.empty(B, N * T, N * T, device=dev) for _ in range(2)]
    work = torch.empty(int(_lib.lib.fg_traj_cov_work_floats(h, B)), device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nbytes = B * (N * T) ** 2 * 4
    for path in (1, 2):
        def run(i): _lib.check(_lib.lib.fg_traj_cov(h, L.data_ptr(), 1e-4, 0, covs[i % 2].data_ptr(), None, work.data_ptr(), path, B, stream))
        for i in range(3): run(i)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for i, (a, b) in enumerate(ev):
            a.record(); run(i); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        print(f"{env_id} B={B} [{N*T}x{N*T}] path {path}: {ts[5]:.3f} ms  {nbytes/ts[5]/1e6:.0f} GB/s  {B/ts[5]*1e3:.0f} cov/s  ({2*N*T*N*T*(D//N)*B/ts[5]/1e9:.1f} TFLOP/s useful)")
    _lib.check(_lib.lib.fg_traj_cov(h, L.data_ptr(), 1e-4, 0, covs[0].data_ptr(), None, work.data_ptr(), 1, B, stream))
    _lib.check(_lib.lib.fg_traj_cov(h, L.data_ptr(), 1e-4, 0, covs[1].data_ptr(), None, work.data_ptr(), 2, B, stream))
    print("max |path1 - path2| / max|cov| =", float((covs[0] - covs[1]).abs().max() / covs[0].abs().max()))
