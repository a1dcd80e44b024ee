"""torchrun, N GPUs of one box: where do the ranks' CPUs / pinned buffers sit relative to their GPUs, and what host->device
bandwidth does each rank get alone and with all ranks copying at once — before and after binding the rank to the CPUs NVML
reports as local to its GPU (fancy_gym_b200.dist.bind_to_gpu_cpus).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/probe_h2d_numa.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    for r in range(world):
        if r == rank:
            print(f"[rank {rank}]", *a, flush=True)
        dist.barrier()


import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(lr)
bus = pynvml.nvmlDeviceGetPciInfo(h).busId
bus = bus.decode() if isinstance(bus, bytes) else bus
short = bus.lower()[-12:]
try:
    numa = open(f"/sys/bus/pci/devices/{short}/numa_node").read().strip()
except OSError as e:
    numa = f"? ({e})"
words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
gpu_cpus = sorted(c for w_i, w in enumerate(words) for c in range(64) if (w >> c) & 1 for c in [w_i * 64 + c])
allowed = sorted(os.sched_getaffinity(0))
say(f"GPU {lr} bus {bus} numa_node {numa}; NVML local CPUs {gpu_cpus[:4]}..{gpu_cpus[-4:] if gpu_cpus else ''} ({len(gpu_cpus)}); "
    f"allowed CPUs {allowed[:4]}..{allowed[-4:]} ({len(allowed)}); intersection {len(set(gpu_cpus) & set(allowed))}")
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -20; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>&1; "
              "ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done")
dist.barrier()

NB = 64 << 20
d = torch.empty(NB, dtype=torch.uint8, device=dev)


def bw(hbuf, reps=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(hbuf, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return NB * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def measure(tag):
    hbuf = torch.empty(NB, dtype=torch.uint8).pin_memory()
    hbuf.fill_(1)
    bw(hbuf, 3)
    alone = None
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone = bw(hbuf)
        dist.barrier()
    dist.barrier()
    together = bw(hbuf, 40)
    t = torch.tensor([together], device=dev)
    dist.all_reduce(t)
    say(f"{tag}: H2D alone {alone:.1f} GB/s, all ranks at once {together:.1f} GB/s (sum over ranks {float(t):.0f} GB/s)")


measure("unbound")
from fancy_gym_b200.dist import bind_to_gpu_cpus  # noqa: E402

say("bind_to_gpu_cpus ->", bind_to_gpu_cpus(lr))
measure("bound  ")
dist.destroy_process_group()
