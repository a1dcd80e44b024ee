"""Population-based search on ONE task context: the cross-entropy method over ProMP weights for fancy_ProMP/HoleReacher-v0.

All envs of the batch are reset with the same seed, i.e. they hold the same hole position / width; every CEM generation is
one env.evaluate() call (a fused rollout of the whole population that leaves the envs at their reset state).

    python examples/cem_holereacher.py [population] [generations]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fancy_gym_b200 as fancy_gym  # noqa: E402


def main(population=16384, generations=30, elite_frac=0.05, seed=3):
    dev = "cuda:0"
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=population, device=dev)
    obs, _ = env.reset(seed=np.full(population, seed))       # the same context in every env
    assert bool((obs == obs[0]).all())
    P = env.action_space.shape[0]
    mu = torch.zeros(P, device=dev)
    sigma = torch.full((P,), 0.5, device=dev)
    n_elite = max(2, int(elite_frac * population))
    gen = torch.Generator(device=dev).manual_seed(0)
    for g in range(generations):
        params = mu + sigma * torch.randn(population, P, generator=gen, device=dev)
        _, ret, terminated, _, info = env.evaluate(params)   # the episode does not advance: same start state next time
        elite = torch.topk(ret, n_elite).indices
        mu = params[elite].mean(0)
        sigma = params[elite].std(0) + 1e-3
        best = int(elite[0])
        print(f"generation {g:2d}: best return {ret[best].item():10.4f}  mean {ret.mean().item():10.3f}  "
              f"collided {info['is_collided'].float().mean().item():.2f}  success {info['is_success'].float().mean().item():.3f}")
    env.close()


if __name__ == "__main__":
    assert torch.cuda.is_available(), "fancy_gym_b200 runs on a CUDA device (there is no CPU fallback)"
    main(*(int(a) for a in sys.argv[1:3]))
