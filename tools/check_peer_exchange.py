"""2+ GPUs (torchrun): the fused peer-store gather (fancy_gym_b200.dist.PeerResultExchange) must deliver exactly the blocks
an NCCL all-gather of the same result blocks delivers.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_peer_exchange.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fancy_gym_b200 as fancy_gym  # noqa: E402
from fancy_gym_b200.dist import PeerResultExchange, all_gather_result_blocks, result_block_views  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
B = 4096 + 64 * rank * 0
env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev)
px = PeerResultExchange(env, ring=4)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
works = []
for step in range(9):
    while len(works) > px.ring - 2:
        works.pop(0).wait()
    env.reset(seed=1000 * rank + step)
    obs, ret, te, tr, info = env.step((0.3 + 0.2 * step) * torch.randn(B, 25, generator=gen, device=dev))
    w = px.publish()
    works.append(w)
    ref = all_gather_result_blocks(env._result_block)          # NCCL, blocking on this stream
    w.wait()
    got = px.gathered()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.equal(got, ref), f"rank {rank} step {step}: peer-store gather differs from the NCCL all-gather"
    r, ln, fl = result_block_views(got, B)
    assert torch.equal(r[rank], ret) and torch.equal(ln[rank], info["trajectory_length"])
    assert int(ln.sum()) > 0
dist.barrier()
if rank == 0:
    print(f"PEER OK: {world} ranks x {B} envs, 9 steps, gathered blocks == ncclAllGather, multicast support: {px.hdl.has_multicast_support}")
dist.destroy_process_group()
