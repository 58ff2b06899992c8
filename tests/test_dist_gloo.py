"""world_size-2 gloo test (CPU) of the multi-GPU host logic: point-range sharding, raw-byte allgather of the
64-byte partials, rank-ordered sum.  The per-rank partial MSM is computed by the oracle here (no GPU);
on the GPU box the same plumbing carries `Context.msm_dev` results (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import halo2_aggregation_b200 as h2a
    from oracle import loader as orc

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = h2a.shard_range(n_total, rank, world)
    bases = orc.gen_bases(1, hi - lo, first=lo, threads=2)
    scalars = orc.gen_scalars(2, hi - lo, first=lo, threads=2)
    partial = orc.msm(bases, scalars, threads=2)
    total = h2a.allgather_sum(partial)
    gathered = h2a.allgather_points(partial)
    q.put((rank, bytes(total), bytes(gathered), bytes(partial)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1001), (3, 700)])
def test_sharded_msm_allgather_sum(orc, world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes(orc.msm(orc.gen_bases(1, n_total), orc.gen_scalars(2, n_total)))
    partials = b"".join(r[3] for r in results)
    for rank, total, gathered, _ in results:
        assert total == want            # every rank holds the identical full result
        assert gathered == partials     # rank order preserved


def test_shard_range_covers_everything():
    import halo2_aggregation_b200 as h2a
    for n, w in [(10, 3), (1 << 22, 8), (7, 8), (0, 2)]:
        spans = [h2a.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_single_process_passthrough(orc):
    import halo2_aggregation_b200 as h2a
    p = orc.gen_bases(4, 1)
    assert bytes(h2a.allgather_sum(p)) == bytes(p)


def _exchange_worker(rank, world, port, m, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import halo2_aggregation_b200 as h2a

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    buf = np.zeros(64 * m, np.uint8)
    for j in range(m):
        if j % world == rank:
            buf[64 * j:64 * j + 64] = (j * 7 + 1) % 251        # this rank's columns
    h2a.make_commitment_exchange(world)(buf)
    q.put((rank, bytes(buf)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m", [(2, 5), (3, 4)])
def test_column_parallel_commitment_exchange(world, m):
    """Host logic of h2a_circuit_set_distribution: after the exchange every rank holds every owner's column."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = b"".join(bytes([(j * 7 + 1) % 251]) * 64 for j in range(m))
    assert all(r[1] == want for r in results)
