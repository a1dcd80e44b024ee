"""GPU: is the fused multi-plan launch (re-packing + plans looped inside one launch) reproducible, and equal to one step() per plan?

    python tools/race_probe.py                                            # plain
    compute-sanitizer --tool racecheck python tools/race_probe.py         # under the race detector
    FG_LIB_PATH=<build with -DFG_ROLLOUT_CHAOS / -DFG_ROLLOUT_CHECK> python tools/race_probe.py

Per trial: the sequential reference (4 step() calls) first, then FUSED step_plans() calls on fresh envs with the same actions; prints,
per fused call, the number of envs whose per-plan lengths / returns differ from the reference.  See profiles/README.md "Sanitizer".
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

dev = torch.device("cuda", 0)
over = {"black_box_kwargs": {"replanning_schedule": lambda p, v, o, a, t: t % 50 == 0}}
B = int(os.environ.get("ENVS", "3001"))
FUSED = int(os.environ.get("FUSED", "4"))
ORDER = os.environ.get("ORDER", "seq-first")
bad_total = 0


def make():
    e = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev, mp_config_override=over)
    e.reset(seed=1)
    return e


def fused(a):
    o = make().step_plans(a)
    torch.cuda.synchronize()
    return o[4]["trajectory_length"].cpu(), o[1].cpu()


def sequential(a):
    e, Ls, Rs = make(), [], []
    for k in range(4):
        o = e.step(a[:, k].contiguous())
        torch.cuda.synchronize()
        Ls.append(o[4]["trajectory_length"].cpu())
        Rs.append(o[1].cpu())
    return torch.stack(Ls, 0), torch.stack(Rs, 0)


for trial in range(int(os.environ.get("TRIALS", "3"))):
    gen = torch.Generator(device=dev).manual_seed(100 + trial)
    P = make().action_space.shape[0]
    a = 0.5 * torch.randn(B, 4, P, generator=gen, device=dev)
    first = fused(a) if ORDER == "fused-first" else None
    L2, R2 = sequential(a)
    runs = ([first] if first else []) + [fused(a) for _ in range(FUSED)]
    counts = [(int((L != L2).any(0).sum()), int((R != R2).any(0).sum())) for L, R in runs]
    bad_total += sum(c[0] + c[1] for c in counts)
    print(f"trial {trial}: {int(L2.sum())} env steps; envs whose (lengths, returns) differ from one step() per plan, per fused call: {counts}",
          flush=True)
print("probe done, mismatching envs in total:", bad_total)
