"""halo2-aggregation_b200 — B200 (sm_100a) implementation of the data-parallel hot path of
halo2-aggregation: BN254 G1 MSM, Fr NTT / coset FFT, and the batched verifier accumulation.

The product is `libh2agg.so` (C ABI, include/h2agg.h).  This package is the thin Python host
binding used by the tests and benchmarks; it mirrors the names of the dependency interface the
reference consumes (`best_multiexp`, `best_fft`, `EvaluationDomain`, `Params::commit_lagrange`;
SURVEY.md §8b).  There is no CPU fallback: constructing a `Context` without the built library or
without a B200 raises.
"""
from .api import (  # noqa: F401
    comm_unique_id,
    Context,
    Bases,
    Circuit,
    serialize_shape,
    Transcript,
    PermutationAssembly,
    EvaluationDomain,
    H2AError,
    best_multiexp,
    best_fft,
    library_path,
    load_library,
    declared_symbols,
    g1_sum,
    fr_root_of_unity,
    xorshift_scalar,
    vk_hash,
)
from .dist import allgather_points, allgather_sum, make_commitment_exchange, shard_range  # noqa: F401
