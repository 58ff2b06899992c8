"""Multi-rank GPU tests: run under torchrun with one process per GPU when the box has at least two —
tests/multi_gpu/comm_check.py (the library's own NCCL plumbing: allgather, broadcast, one MSM over per-rank point ranges against
the oracle) and tests/multi_gpu/dist_prove_check.py (ONE proof spread over the ranks, byte-identical with the ORACLE prover's on every
rank, a failing lookup refused on every rank, recovery).  Skipped on a single-GPU box; records of 8-rank runs: profiles/r2_*_n8.json."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ranks():
    import torch
    return min(torch.cuda.device_count(), 8)


@pytest.mark.parametrize("script", ["comm_check.py", "dist_prove_check.py"])
def test_under_torchrun(script):
    n = _ranks()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu", script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.loads(r.stdout.strip().splitlines()[-1])
    assert verdict["world"] == n and verdict["all_ranks_ok"] and all(verdict["checks"].values()), verdict
