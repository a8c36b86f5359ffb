"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Instances never interact (the reference runs one DDPSolver object per problem,
nmpc_ddp/include/nmpc_ddp/DDPSolver.h:329-374), so the batch splits into contiguous chunks with no
data-path collective.  The only exchange is the optional all-gather of first-step controls u_list[0]
(what an MPC loop applies, e.g. nmpc_ddp/tests/src/TestDDPBipedal.cpp:254).
"""
import numpy as np


def shard_range(total, world_size, rank):
    """Contiguous [begin, end) of `total` instances owned by `rank`: floor(total / world) each, the
    remainder spread one per rank over the lowest ranks."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(int(total), int(world_size))
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def shard_sizes(total, world_size):
    return [shard_range(total, world_size, r)[1] - shard_range(total, world_size, r)[0] for r in range(world_size)]


def gather_first_controls(u0_local, total, group=None):
    """All-gather of per-rank first-step controls [B_rank, NU] into [total, NU] on every rank.

    Works for ragged shards (pads to the largest shard).  `u0_local` is a torch tensor on the device
    the process group communicates on (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    nu = u0_local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty((total, nu), dtype=u0_local.dtype, device=u0_local.device)
        dist.all_gather_into_tensor(out, u0_local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, nu), dtype=u0_local.dtype, device=u0_local.device)
    buf[:u0_local.shape[0]] = u0_local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
