import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import halo2_aggregation_b200 as h2a
import prove_bench as pb
from oracle import loader as orc
ctx = h2a.Context(0)
k = 20; n = 1 << k
secret = 0x0f1e2d3c4b5a69788796a5b4c3d2e1f00112233445566778899aabbccddeeff % pb.R
g, gl = ctx.kzg_setup(k, pb.fr(ctx, [secret]))
shape, inst_b, adv_b, fixed_b, sigmas_b = pb.build(ctx, k)
cols = [sigmas_b[32 * n * j:32 * n * (j + 1)] for j in range(9)]
tag = sys.argv[1] if len(sys.argv) > 1 else ""
rb = ctx.upload_bases(orc.gen_bases(7, n))
uni = [orc.gen_scalars(100 + j, n) for j in range(9)]
for name, bases, cc in (("kzg+sigma", gl, cols), ("rand+sigma", rb, cols), ("kzg+uniform", gl, uni)):
    bases.precompute(0)
    ref = [bytes(ctx.msm(bases, c)) for c in cc]
    bases.precompute(16)
    for rep in range(3):
        bad = [j for j in range(9) if bytes(ctx.msm(bases, cc[j])) != ref[j]]
        print(tag, name, "rep", rep, "bad", bad, flush=True)
