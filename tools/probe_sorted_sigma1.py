"""GPU experiment: how much of the sigma = 1.0 slowdown is lane divergence in episode LENGTH?  The same episodes are run in
their natural order and sorted by their (known) length, so that the lanes of a warp end together — the upper bound of what
re-packing live envs can recover.  Usage: python tools/probe_sorted_sigma1.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import fancy_gym_b200 as fancy_gym

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
for sigma in (0.25, 1.0):
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev, context_sampler="device")
    base = env.unwrapped
    env.reset(seed=1)
    gen = torch.Generator(device=dev).manual_seed(0)
    params = (sigma * torch.randn(B, 25, generator=gen, device=dev)).contiguous()
    st = SimpleNamespace(q=base.q.clone(), v=torch.zeros_like(base.v), steps=torch.zeros_like(base.steps),
                         done=torch.zeros_like(base.done), ctx=base.ctx.clone())
    env.launch(params, state=st, keep_state=True)
    length = env._len.clone()
    ret0 = env._ret.clone()
    perm = torch.argsort(length, stable=True)
    st_s = SimpleNamespace(q=st.q[perm].contiguous(), v=st.v[perm].contiguous(), steps=st.steps[perm].contiguous(),
                           done=st.done[perm].contiguous(), ctx=st.ctx[perm].contiguous())
    params_s = params[perm].contiguous()

    def timed(p, s, reps=20):
        for _ in range(3):
            env.launch(p, state=s, keep_state=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            env.launch(p, state=s, keep_state=True)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    t_nat = timed(params, st)
    t_sort = timed(params_s, st_s)
    env.launch(params_s, state=st_s, keep_state=True)
    assert torch.equal(torch.nan_to_num(env._ret), torch.nan_to_num(ret0[perm])), "sorting changed results"
    steps = int(length.sum())
    hist = torch.histc(length.float(), bins=10, min=0, max=200).long().tolist()
    print(f"sigma {sigma} B {B}: mean length {steps / B:.1f}, natural order {t_nat:.4f} ms ({steps / t_nat * 1e3:.3e} env-steps/s), "
          f"sorted by length {t_sort:.4f} ms ({steps / t_sort * 1e3:.3e}); length histogram (20-step bins) {hist}", flush=True)
