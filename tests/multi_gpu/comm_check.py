#!/usr/bin/env python
"""Checks the library's own multi-GPU plumbing (csrc/comm.cu) under torchrun, one process per GPU:
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu/comm_check.py
allgather of raw bytes (host and device buffers), broadcast, and one MSM over per-rank point ranges compared with the
oracle's best_multiexp over all the points.  torch.distributed only carries the two unique ids and the final verdict."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import halo2_aggregation_b200 as h2a


def main():
    rank, local, world = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = h2a.Context(local)
    ctx.comm_init_torch()
    ok = {}
    # allgather of host bytes
    mine = np.full(96, rank + 1, np.uint8)
    got = ctx.comm_allgather(mine)
    ok["allgather_host"] = bool(np.array_equal(got, np.repeat(np.arange(1, world + 1, dtype=np.uint8), 96)))
    # allgather of device buffers, in place (send = recv + rank * bytes)
    nb = 1 << 20
    d = ctx.dev_alloc(nb * world)
    ctx.h2d(d + nb * rank, np.full(nb, 7 * rank + 3, np.uint8))
    ctx.comm_allgather_dev(d + nb * rank, d, nb)
    ctx.sync()
    back = ctx.d2h(d, nb * world)
    ok["allgather_dev"] = bool(all((back[nb * r:nb * (r + 1)] == (7 * r + 3) % 256).all() for r in range(world)))
    # broadcast from the last rank
    ctx.h2d(d, np.full(nb, 100 + rank, np.uint8))
    ctx.comm_broadcast_dev(d, nb, world - 1)
    ctx.sync()
    ok["broadcast"] = bool((ctx.d2h(d, nb) == 100 + world - 1).all())
    ctx.dev_free(d)
    # one MSM of `total` points split into per-rank ranges vs the oracle over all of them
    total = 1 << 14
    lo, hi = h2a.shard_range(total, rank, world)
    n = hi - lo
    d_b, d_s = ctx.dev_alloc(64 * n), ctx.dev_alloc(32 * n)
    ctx.gen_bases_dev(11, n, d_b, first=lo)
    ctx.gen_scalars_dev(12, n, d_s, first=lo)
    bases = ctx.bases_from_device(d_b, n)
    res = ctx.msm_sharded(bases, d_s, n)
    if rank == 0:
        from oracle import loader as orc
        orc.load()
        d_ab, d_as = ctx.dev_alloc(64 * total), ctx.dev_alloc(32 * total)
        ctx.gen_bases_dev(11, total, d_ab)
        ctx.gen_scalars_dev(12, total, d_as)
        want = orc.msm(ctx.d2h(d_ab, 64 * total), ctx.d2h(d_as, 32 * total))
        ok["msm_sharded_vs_oracle"] = bytes(res) == bytes(want)
    allres = ctx.comm_allgather(res)
    ok["msm_same_on_all_ranks"] = bool(all(bytes(allres[64 * r:64 * r + 64]) == bytes(res) for r in range(world)))
    flag = torch.tensor([int(all(ok.values()))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "checks": ok, "all_ranks_ok": bool(flag.item())}), flush=True)
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
