#!/usr/bin/env python
"""Parity of ONE proof spread over several GPUs (native distribution, csrc/comm.cu + plonk_prove.cu) under torchrun:
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu/dist_prove_check.py
every rank proves the test circuits of tests/circuits.py with the library and the bytes must equal the ORACLE prover's
(oracle/plonk.py) on every rank; a lookup input outside its table must fail on every rank (no rank left in a collective)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import circuits
import halo2_aggregation_b200 as h2a
from oracle import loader as orc
from oracle import plonk as pk
from oracle import pymodel as pm


def frs(vals):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for v in vals), dtype=np.uint8) if vals else np.zeros(0, np.uint8)


def cols(c):
    return np.frombuffer(b"".join(pm.fr_mont_bytes(v) for col in c for v in col), dtype=np.uint8)


def main():
    rank, local, world = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    orc.load()
    ctx = h2a.Context(local)
    ctx.comm_init_torch()
    ok = {}
    cases = [("my_circuit k=6", circuits.my_circuit(k=6, table_bits=4)), ("wide k=6", circuits.wide_circuit(k=6)),
             ("many_rotations k=6", circuits.many_rotations_circuit(k=6)), ("my_circuit k=11", circuits.my_circuit(k=11, table_bits=8))]
    for name, c in cases:
        params, keys = circuits.setup(orc, c)
        want, want_inst = pk.create_proof(orc, params, c["shape"], keys, c["instance"], c["advice"], seed=5)
        shape = c["shape"]
        g, gl = ctx.upload_bases(params.g), ctx.upload_bases(params.g_lagrange)
        circ = h2a.Circuit(ctx, shape, frs(shape.constants))
        circ.set_keys(g, gl, cols(c["fixed"]), cols(keys.sigmas), frs([keys.vk_hash]), frs([shape.coset_shift]))
        circ.set_distribution(rank, world, native=True)
        proof, inst = circ.prove(cols(c["instance"]), cols(c["advice"]), frs(pk.blinds_buffer(shape, 5)))
        ok[name] = proof == want
        again, _ = circ.prove(cols(c["instance"]), cols(c["advice"]), frs(pk.blinds_buffer(shape, 5)))
        ok[name + " (second call)"] = again == want
        if name.startswith("my_circuit k=6"):
            bad = [col[:] for col in c["advice"]]
            bad[0][0] = 12345678
            try:
                circ.prove(cols(c["instance"]), cols(bad), frs(pk.blinds_buffer(shape, 5)))
                ok["bad lookup refused"] = False
            except h2a.H2AError:
                ok["bad lookup refused"] = True
            third, _ = circ.prove(cols(c["instance"]), cols(c["advice"]), frs(pk.blinds_buffer(shape, 5)))
            ok["proves again after the failure"] = third == want
        circ.free(); g.free(); gl.free()
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
