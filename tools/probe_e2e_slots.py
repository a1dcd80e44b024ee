"""torchrun, N GPUs: end-to-end step time of EpisodePipeline (reset + H2D of the parameters from pinned host memory + step() +
D2H of the results, per batch) with 2, 3, 4 batches in flight; max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/probe_e2e_slots.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, K = 65536, int(os.environ.get("STEPS", "300"))
GRAPHS = os.environ.get("GRAPHS", "0") == "1"
for slots in (2, 3, 4):
    env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev, context_sampler="device",
                         mp_config_override={"black_box_kwargs": {"result_sets": slots}})
    env.reset(seed=rank)
    pipe = fancy_gym.EpisodePipeline(env, slots=slots, graphs=GRAPHS)
    for hp in pipe.host_params:
        hp.copy_(0.25 * torch.randn(hp.shape))

    def run(n):
        steps, first = 0, pipe.next_slot
        for i in range(n):
            slot = (first + i) % pipe.SLOTS
            if i >= pipe.SLOTS:
                steps += int(pipe.wait(slot)[1].sum())
            pipe.submit(slot)
        for i in range(max(0, n - pipe.SLOTS), n):
            steps += int(pipe.wait((first + i) % pipe.SLOTS)[1].sum())
        return steps

    run(12)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps = run(K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    n = torch.tensor([float(steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n)
    if rank == 0:
        print(f"graphs {int(GRAPHS)} slots {slots}: {1e3 * float(t) / K:.4f} ms per step (max over {world} ranks), {float(n) / float(t):.3e} env-steps/s", flush=True)
    env.close()
if world > 1:
    dist.destroy_process_group()
