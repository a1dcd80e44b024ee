"""world_size-2 `gloo` tests of the multi-GPU host logic: contiguous env shards, no data-path
collective, one all-gather of (return, length, flags) per black-box step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fancy_gym_b200.dist import gather_episode_results, shard_bounds, shard_sizes


def test_shard_bounds_partition_the_batch():
    for total in (0, 1, 7, 8, 65536, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # every rank draws the same GLOBAL results
        ret = torch.as_tensor(rng.standard_normal(total) * 100)
        length = torch.as_tensor(rng.integers(1, 201, total).astype(np.int32))
        flags = torch.as_tensor(rng.integers(0, 16, total).astype(np.uint8))
        lo, hi = shard_bounds(total, world, rank)
        g_ret, g_len, g_flags = gather_episode_results(ret[lo:hi].clone(), length[lo:hi].clone(), flags[lo:hi].clone(),
                                                       total_envs=total)
        ok = bool(torch.equal(g_ret, ret) and torch.equal(g_len, length) and torch.equal(g_flags, flags))
        q.put((rank, ok, int(g_ret.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 65, 1])
def test_gather_episode_results_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok and n == total for _, ok, n in res), res


def test_gather_is_identity_without_process_group():
    ret, ln, fl = torch.ones(3, dtype=torch.float64), torch.ones(3, dtype=torch.int32), torch.ones(3, dtype=torch.uint8)
    a, b, c = gather_episode_results(ret, ln, fl)
    assert a is ret and b is ln and c is fl


def _block_worker(rank, world, port, B, q):
    from fancy_gym_b200.dist import all_gather_result_blocks, result_block_bytes, result_block_views
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(rank)
        block = torch.zeros(result_block_bytes(B), dtype=torch.uint8)
        ret, length, flags = result_block_views(block, B)
        ret.copy_(torch.as_tensor(rng.standard_normal(B))); length.copy_(torch.as_tensor(rng.integers(1, 201, B).astype(np.int32)))
        flags.copy_(torch.as_tensor(rng.integers(0, 16, B).astype(np.uint8)))
        g_ret, g_len, g_flags = result_block_views(all_gather_result_blocks(block), B)
        ok = g_ret.shape == (world, B)
        # the overlapped form: (gathered, work), readable after work.wait()
        out2, work = all_gather_result_blocks(block, async_op=True)
        work.wait()
        ok &= bool(torch.equal(result_block_views(out2, B)[0], g_ret))
        for r in range(world):
            rr = np.random.default_rng(r)
            ok &= bool(np.array_equal(g_ret[r].numpy(), rr.standard_normal(B)))
            ok &= bool(np.array_equal(g_len[r].numpy(), rr.integers(1, 201, B).astype(np.int32)))
            ok &= bool(np.array_equal(g_flags[r].numpy(), rr.integers(0, 16, B).astype(np.uint8)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 37])
def test_result_blocks_gather_world2(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_block_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_bind_to_gpu_cpus_is_safe_without_a_gpu():
    """one process per GPU: the rank runs on the CPUs NVML reports next to its GPU; without NVML / a GPU it leaves the
    affinity alone and says so"""
    import os
    from fancy_gym_b200.dist import bind_to_gpu_cpus, gpu_local_cpus
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_cpus(0)
    assert cpus == gpu_local_cpus(0) or set(cpus) <= before
    after = os.sched_getaffinity(0)
    assert after == (set(cpus) if cpus else before)
    os.sched_setaffinity(0, before)
