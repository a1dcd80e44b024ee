"""GPU: sigma = 1.0 x 65 536 (episodes of very different length): device time per launch with 1 - 6 batches in flight on their own
streams (rotating input sets, keep_state launches like bench.py's `value`)."""
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

dev = torch.device("cuda", 0)
B, SETS = int(os.environ.get("ENVS", "65536")), 8
gen = torch.Generator(device=dev).manual_seed(0)
for n_streams in (1, 2, 3, 4, 6):
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev, context_sampler="device",
                         mp_config_override={"black_box_kwargs": {"result_sets": max(2, n_streams)}})
    base = env.unwrapped
    sets = []
    for i in range(SETS):
        env.reset(seed=50_000 + i)
        sets.append(dict(params=(1.0 * torch.randn(B, 25, generator=gen, device=dev)).contiguous(),
                         state=SimpleNamespace(q=base.q.clone(), v=torch.zeros_like(base.v), steps=torch.zeros_like(base.steps),
                                               done=torch.zeros_like(base.done), ctx=base.ctx.clone())))
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)] if n_streams > 1 else [main]
    tot = [torch.zeros((), dtype=torch.int64, device=dev) for _ in streams]

    def run(n):
        for i in range(n):
            st = streams[i % len(streams)]
            with torch.cuda.stream(st):
                sc = sets[i % SETS]
                env.launch(sc["params"], state=sc["state"], keep_state=True)
                tot[i % len(streams)] += env._len.sum()
    run(2 * n_streams)
    torch.cuda.synchronize()
    for t in tot:
        t.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for st in streams:
        st.wait_stream(main)
    N = 120
    run(N)
    for st in streams:
        main.wait_stream(st)
    b.record(main)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / N
    steps = sum(int(t) for t in tot) / N
    print(f"{n_streams} in flight: {ms:.4f} ms per launch, {steps / ms * 1e3:.3e} env-steps/s", flush=True)
    env.close()
