"""Multi-GPU plumbing for the MSM: one process per GPU, point-range shards, raw-byte
allgather of the 64-byte affine partial results, identical rank-ordered sum on every rank.

`torch.distributed` is plumbing only (NCCL on the GPUs, gloo in the CPU tests)."""
import numpy as np

from .api import g1_sum


def shard_range(n_total, rank, world):
    """Contiguous point range [lo, hi) of rank `rank` (the last rank takes the remainder)."""
    per = n_total // world
    lo = rank * per
    hi = n_total if rank == world - 1 else lo + per
    return lo, hi


def allgather_points(partial, group=None, device=None):
    """All ranks' 64-byte partials concatenated in rank order (numpy uint8, 64*world bytes)."""
    import torch
    import torch.distributed as dist

    partial = np.ascontiguousarray(partial, dtype=np.uint8).reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial.copy()
    world = dist.get_world_size(group)
    mine = torch.from_numpy(partial.copy())
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return torch.cat(out).cpu().numpy()


def allgather_sum(partial, group=None, device=None):
    """Sum over ranks of per-rank partial MSM results; every rank returns the same 64 bytes."""
    return g1_sum(allgather_points(partial, group, device))


def make_commitment_exchange(world, group=None, device=None):
    """The `exchange(buf: numpy uint8[64*m])` step of column-parallel proving (h2a_circuit_set_distribution): column j
    of a batch is owned by rank j % world; after the allgather every rank holds every owner's column."""
    import torch
    import torch.distributed as dist

    def exchange(arr):
        m = arr.size // 64
        mine = torch.from_numpy(arr.copy())
        if device is not None:
            mine = mine.to(device)
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
        allb = torch.stack(out).cpu().numpy()
        for j in range(m):
            arr[64 * j:64 * j + 64] = allb[j % world, 64 * j:64 * j + 64]

    return exchange
