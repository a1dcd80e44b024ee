"""torchrun probe: all_gather_into_tensor latency / bandwidth by message size (device-timed)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for nbytes in (1 << 10, 1 << 16, 852000, 1 << 22, 1 << 26):
    x = torch.zeros(nbytes, dtype=torch.uint8, device=dev); out = torch.zeros(world * nbytes, dtype=torch.uint8, device=dev)
    for _ in range(5): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    a.record()
    for _ in range(n): dist.all_gather_into_tensor(out, x)
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        ms = a.elapsed_time(b) / n
        print(f"all_gather {nbytes:>9d} B/rank: {ms*1e3:8.1f} us  {world*nbytes/ms/1e6:8.1f} GB/s", flush=True)
dist.destroy_process_group()
