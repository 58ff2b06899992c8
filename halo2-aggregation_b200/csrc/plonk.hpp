// plonk.hpp — the opaque `h2a_circuit` of include/h2agg.h: a circuit shape plus the key material the
// verifier glue (plonk_verify.cu) and the prover pipeline (plonk_prove.cu) need.
#pragma once
#include <vector>

#include "ctx.hpp"
#include "plonk_shape.hpp"

struct ProverState;  // plonk_prove.cu

struct h2a_circuit {
    h2a_plonk::Shape shape;
    // verifying key: commitments to the fixed columns and the permutation (sigma) polynomials, and the
    // transcript scalar derived from the pinned vk (src/verifier.rs:341-358; an input here)
    std::vector<uint8_t> fixed_comms, sigma_comms;
    uint8_t vk_hash[32] = {0};
    bool has_vk = false;
    ProverState* prover = nullptr;
    // column-parallel commitments: this process commits the columns j with j % world == rank of every
    // batch and the 64-byte results are exchanged through the caller's collective (NCCL allgather in practice)
    int dist_rank = 0, dist_world = 1;
    h2a_exchange_fn dist_exchange = nullptr;
    void* dist_user = nullptr;
};

void h2a_prover_state_free(h2a_ctx* ctx, ProverState* p);
