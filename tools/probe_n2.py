"""torchrun probe: per-phase device time of the multi-GPU bench step (state reload | rollout | all-gather | bookkeeping)."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fancy_gym_b200 as fancy_gym
from fancy_gym_b200.dist import all_gather_result_blocks
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B = 65536
env = fancy_gym.make("fancy_ProMP/HoleReacher-v0", num_envs=B, device=dev)
base = env.unwrapped
env.reset(seed=1 + rank)
s = dict(params=(0.25 * torch.randn(B, 25, device=dev)).contiguous(), q=base.q.clone(), ctx=base.ctx.clone())
gathered = torch.zeros(world * env._result_block.numel(), dtype=torch.uint8, device=dev)
total = torch.zeros((), dtype=torch.int64, device=dev)
K = 100
if os.environ.get("PRE_WARM"):
    for _ in range(int(os.environ["PRE_WARM"])): all_gather_result_blocks(env._result_block, out=gathered)
    torch.cuda.synchronize()
def E(): return torch.cuda.Event(enable_timing=True)
for mode in ("gather", "no-gather", "gather"):
    ev = [[E() for _ in range(5)] for _ in range(K)]
    for warm in range(5):
        env.launch(s["params"]); all_gather_result_blocks(env._result_block, out=gathered)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        ev[i][0].record()
        base.q.copy_(s["q"]); base.ctx.copy_(s["ctx"]); base.v.zero_(); base.steps.zero_(); base.done.zero_()
        ev[i][1].record()
        env.launch(s["params"])
        ev[i][2].record()
        if mode == "gather": all_gather_result_blocks(env._result_block, out=gathered)
        ev[i][3].record()
        total.add_(env._len.sum())
        ev[i][4].record()
    host = (time.perf_counter() - t0) / K * 1e3
    torch.cuda.synchronize()
    ph = [sum(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(K)) / K for j in range(4)]
    tot = ev[0][0].elapsed_time(ev[K - 1][4]) / K
    print(f"rank {rank} {mode:10s} reload {ph[0]:.3f} rollout {ph[1]:.3f} gather {ph[2]:.3f} sum {ph[3]:.3f} | step {tot:.3f} ms, host issue {host:.3f} ms", flush=True)
dist.destroy_process_group()
