"""Host-side multi-GPU logic on CPU: contiguous batch shards and the first-control gather over a
world_size-2 gloo group (the N>1 path of bench.py uses the same helpers over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from nmpc_b200.sharding import gather_first_controls, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 4096, 4097, 131072):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (b0, e0), (b1, e1) in zip(ranges[:-1], ranges[1:]):
                assert e0 == b1 and e0 >= b0
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank solves its own shard (with the CPU oracle standing in for the GPU engine here) ...
        b, e = shard_range(total, world, rank)
        x0 = O.cartpole_x0(total, 5)[b:e]
        cfg = O.ddp_config(max_iter=3, horizon_steps=20)
        r = O.ddp_solve_batch("cartpole", O.default_params("cartpole"), cfg, x0, np.zeros((e - b, 20, 1)), nthreads=1)
        u0_local = torch.from_numpy(r["u"][:, 0, :].copy())
        # ... and only the first-step controls cross ranks
        u0_all = gather_first_controls(u0_local, total)
        ret[rank] = u0_all.numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 9])
def test_gather_first_controls_gloo_world2(total):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, total, ret), nprocs=world, join=True)
        got = [ret[r] for r in range(world)]
    cfg = O.ddp_config(max_iter=3, horizon_steps=20)
    full = O.ddp_solve_batch("cartpole", O.default_params("cartpole"), cfg, O.cartpole_x0(total, 5),
                             np.zeros((total, 20, 1)), nthreads=1)
    for g in got:
        assert g.shape == (total, 1)
        np.testing.assert_array_equal(g, full["u"][:, 0, :])
