"""Checker helpers that need the oracle (test infrastructure): imported by the tests and, lazily, by measurement tools that print
a parity verdict next to a timing (`tools/prove_bench.py --oracle-check`).  Nothing under `halo2-aggregation_b200/` imports this."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in (ROOT, os.path.join(ROOT, "tools")):
    if d not in sys.path:
        sys.path.insert(0, d)


def proof_accepted(check):
    """The oracle's verifier (oracle/plonk.py, restating VerifierChip::_verify_proof) replays the proof described by `check` (the
    `_check` dict of tools/prove_bench.run) and the pairing relation must hold with the known setup secret.  O(proof size), any k."""
    import bench_extras
    from oracle import loader as orc, plonk as pk, pymodel as pm
    orc.load()
    ok, _ = bench_extras._oracle_accepts(dict(pk=pk, pm=pm), check)
    return bool(ok)
