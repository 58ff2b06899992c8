/* h2agg.h — C ABI of libh2agg.so: the B200 (sm_100a) implementation of the data-parallel hot path
 * of halo2-aggregation: BN254 G1 MSM, Fr NTT / coset FFT / domain extension, and the batched
 * verifier (GWC multi-open) accumulation.
 *
 * The reference has no FFI of its own: it is generic Rust calling the `halo2` dependency
 * (Cargo.toml:12).  Each entry point below names the dependency function whose call it replaces
 * and the line of the reference through which that call is reached.  INTEGRATION.md shows the Rust
 * `extern "C"` block and the shim a maintainer adds to `halo2::arithmetic` / `halo2::poly`.
 *
 * Conventions
 *  - Status: 0 = ok, negative = error; h2a_last_error(ctx) returns a message.  No aborts, no
 *    exceptions cross the boundary.  There is no CPU fallback: without a usable CUDA device
 *    h2a_init fails and every other call fails with H2A_ERR_NO_DEVICE.
 *  - Field elements: 32 bytes, 4 x u64 little-endian limbs, Montgomery form (R = 2^256) — the
 *    in-memory form of `bn256::Fr` / `bn256::Fq`, so a Rust `&[Fr]` is passed as `*const u8`.
 *  - G1 affine points: 64 bytes x || y (each as above); the identity is 64 zero bytes.
 *  - Results are canonical: an MSM returns the affine form of the unique group element, so any
 *    two correct implementations agree bit for bit.
 *  - `*_dev` variants take device pointers (inputs already resident in HBM); all others take host
 *    pointers and do their own host<->device copies.
 *  - A ctx is bound to one GPU and serialises its calls on one CUDA stream; use one ctx per GPU
 *    (one process per GPU under torchrun) and distinct ctxs from distinct threads.
 */
#ifndef H2AGG_H
#define H2AGG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2A_OK 0
#define H2A_ERR_INVALID (-1)   /* bad argument (NULL, size mismatch, log_n out of range, ...) */
#define H2A_ERR_CUDA (-2)      /* a CUDA runtime call failed; see h2a_last_error */
#define H2A_ERR_NO_DEVICE (-3) /* no CUDA device / wrong architecture */
#define H2A_ERR_OOM (-4)       /* device or host allocation failed */
#define H2A_ERR_PROOF (-5)     /* malformed proof bytes in the verifier glue */

typedef struct h2a_ctx h2a_ctx;
typedef struct h2a_bases h2a_bases;

/* ---- lifecycle ------------------------------------------------------------------------- */
int h2a_version(void);
int h2a_device_count(void);
/* Create a context on CUDA device `device`. */
int h2a_init(h2a_ctx** out, int device);
int h2a_destroy(h2a_ctx* ctx);
const char* h2a_last_error(const h2a_ctx* ctx);
/* The cudaStream_t all work of this ctx is launched on (for event timing by the caller). */
void* h2a_stream(h2a_ctx* ctx);
/* Block until everything queued on the ctx stream has finished. */
int h2a_sync(h2a_ctx* ctx);
/* Device memory helpers so a host program needs no other CUDA binding. */
int h2a_dev_alloc(h2a_ctx* ctx, size_t bytes, void** out_dev);
int h2a_dev_free(h2a_ctx* ctx, void* dev);
int h2a_copy_h2d(h2a_ctx* ctx, void* dev, const void* host, size_t bytes);
int h2a_copy_d2h(h2a_ctx* ctx, void* host, const void* dev, size_t bytes);

/* ---- G1 MSM ----------------------------------------------------------------------------
 * Replaces halo2 `arithmetic::best_multiexp(coeffs, bases)` as reached through
 * `Params::commit_lagrange` / `Params::commit` (examples/simple-example.rs:638-640 and every
 * commitment inside `create_proof`, :606-613, :702-709).
 * Bases (`Params.g`, `Params.g_lagrange`) are uploaded once and stay resident. */
int h2a_bases_upload(h2a_ctx* ctx, const uint8_t* affine_xy, size_t n, h2a_bases** out);
/* Wrap bases that already live in device memory (no copy; caller keeps ownership of d_affine). */
int h2a_bases_from_device(h2a_ctx* ctx, const void* d_affine_xy, size_t n, h2a_bases** out);
int h2a_bases_free(h2a_ctx* ctx, h2a_bases* bases);
size_t h2a_bases_len(const h2a_bases* bases);
/* Optional, once per bases handle: build window tables T[w][i] = 2^(window_bits*w) * P_i next to the bases
 * (ceil(254/window_bits) * n * 64 bytes of HBM; window_bits in 11..20, -1 = automatic, 0 drops the tables).  Every later MSM
 * over this handle then needs ceil(254/window_bits) additions per point into ONE shared bucket set instead of
 * one bucket set per window.  The result of an MSM is unchanged (same group element, bit for bit). */
int h2a_bases_precompute(h2a_ctx* ctx, h2a_bases* bases, int window_bits);

/* out_affine = sum_{i<n} scalars[i] * bases[offset + i] */
int h2a_msm_g1(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const uint8_t* scalars, size_t n,
               uint8_t out_affine[64]);
int h2a_msm_g1_dev(h2a_ctx* ctx, const h2a_bases* bases, size_t offset, const void* d_scalars, size_t n,
                   uint8_t out_affine[64]);
/* m scalar vectors (columns) over the same bases: out_affine[j] = MSM(scalars[j][0..n[j]), bases[0..n[j])) */
int h2a_msm_g1_batch(h2a_ctx* ctx, const h2a_bases* bases, const uint8_t* const* scalars, const size_t* n, int m,
                     uint8_t* out_affine /* m*64 */);
/* Same with device-resident scalars.  Both batch calls pipeline the columns over two CUDA streams (lanes): the
 * latency-bound tail of one MSM overlaps the bucket accumulation of the next. */
int h2a_msm_g1_batch_dev(h2a_ctx* ctx, const h2a_bases* bases, const void* const* d_scalars, const size_t* n, int m,
                         uint8_t* out_affine /* m*64 */);
/* One-shot MSM over caller-supplied host bases (verifier-side sums; Params not involved). */
int h2a_msm_g1_adhoc(h2a_ctx* ctx, const uint8_t* bases_affine, const uint8_t* scalars, size_t n,
                     uint8_t out_affine[64]);
/* Sum of m affine points on the host (combines per-GPU partial MSM results after the allgather). */
int h2a_g1_sum(const uint8_t* points_affine, size_t m, uint8_t out_affine[64]);
/* Force the Pippenger window width (0 = automatic).  For tuning and tests. */
int h2a_msm_set_window(h2a_ctx* ctx, int c);
/* Bucket accumulation algorithm: 1 (default) pairwise tree of batched affine additions; 0 serial XYZZ mixed
 * additions with one thread per bucket task.  Same result bit for bit; for A/B measurement and tests. */
int h2a_msm_set_algorithm(h2a_ctx* ctx, int algo);
/* h2a_msm_g1 with >= 2^21 host scalars is cut into `pieces` point ranges so that the scalar copy of one range
 * overlaps the computation of the previous one (default 2; 1 = one copy then one MSM).  Same result bit for bit. */
int h2a_msm_set_host_split(h2a_ctx* ctx, int pieces);
/* h2a_msm_g1_batch_dev over bases with precomputed tables and columns of equal length commits up to `cols` columns
 * in ONE pass (one scan, one addition tree and one bucket reduction for the whole group; default 8, 1 = one MSM per
 * column); `cols_host` is the same limit inside the prover when the columns are copied from host memory on the way
 * (default 2).  Same results bit for bit. */
int h2a_msm_set_group(h2a_ctx* ctx, int cols, int cols_host);

/* ---- Fr NTT ------------------------------------------------------------------------------
 * Replaces halo2 `arithmetic::best_fft(a, omega, log_n)` and the `EvaluationDomain` methods
 * built on it (`lagrange_to_coeff`, `coeff_to_extended`, `extended_to_coeff`); the reference
 * touches the domain at src/verifier.rs:252 (get_omega) and :431 (get_quotient_poly_degree).
 * Natural order in, natural order out: a[i] <- sum_j a[j] * omega^(i*j).
 *   inverse != 0 : omega is the FORWARD root; the transform uses omega^-1 and scales by 1/n.
 *   coset_shift  : NULL, or g such that
 *                  forward: a[j] is multiplied by g^j before the transform  (evaluate on g*<omega>)
 *                  inverse: result[j] is multiplied by g^-j after the transform. */
int h2a_ntt(h2a_ctx* ctx, uint8_t* a, uint32_t log_n, const uint8_t omega[32], int inverse,
            const uint8_t* coset_shift);
int h2a_ntt_dev(h2a_ctx* ctx, void* d_a, uint32_t log_n, const uint8_t omega[32], int inverse,
                const uint8_t* coset_shift);
/* coeff_to_extended: 2^k coefficients -> 2^ext_k evaluations on the coset g*<omega_ext>
 * (zero padded, omega_ext = ROOT_OF_UNITY^(2^(28-ext_k))). */
int h2a_coeff_to_extended(h2a_ctx* ctx, const uint8_t* coeffs, uint32_t k, uint32_t ext_k,
                          const uint8_t coset_shift[32], uint8_t* out /* 2^ext_k * 32 */);
/* extended_to_coeff: 2^ext_k coset evaluations -> 2^ext_k coefficients (caller truncates). In place. */
int h2a_extended_to_coeff(h2a_ctx* ctx, uint8_t* ext, uint32_t ext_k, const uint8_t coset_shift[32]);
/* The same two on device-resident polynomials (no copies; d_out must not overlap d_coeffs). */
int h2a_coeff_to_extended_dev(h2a_ctx* ctx, const void* d_coeffs, uint32_t k, uint32_t ext_k,
                              const uint8_t coset_shift[32], void* d_out /* 2^ext_k * 32 */);
int h2a_extended_to_coeff_dev(h2a_ctx* ctx, void* d_ext, uint32_t ext_k, const uint8_t coset_shift[32]);
/* omega_k = ROOT_OF_UNITY^(2^(28-k)), Montgomery form (EvaluationDomain::get_omega). */
int h2a_fr_root_of_unity(uint32_t k, uint8_t out[32]);

/* ---- verifier glue -------------------------------------------------------------------------
 * Replaces the native values computed by `MultiopenChip::calc_witness`
 * (src/multiopen.rs:271-509) with ONE batched device MSM per output point:
 *   queries are grouped by rotation in ascending order, insertion order kept inside a set
 *   (src/multiopen.rs:19-45); with S sets, m_i queries in set i, z_i = x * omega^rot_i:
 *     W  = sum_i u^(S-1-i) W_i          ZW = sum_i u^(S-1-i) z_i W_i
 *     F  = sum_i u^(S-1-i) sum_j v^(m_i-1-j) C_ij
 *     E  = -(sum_i u^(S-1-i) sum_j v^(m_i-1-j) eval_ij) * G1
 * out_efwzw = e || f || w || zw (affine), the order of the `[G1Affine; 4]` the reference packs into
 * public inputs (examples/simple-example.rs:668-671, src/verifier.rs:739-742).
 * Returns H2A_ERR_INVALID when n_ws differs from the number of distinct rotations. */
int h2a_verify_accumulate(h2a_ctx* ctx, const uint8_t* commitments /* nq*64 */, const int32_t* rotations /* nq */,
                          const uint8_t* evals /* nq*32 */, size_t nq, const uint8_t* ws /* n_ws*64 */, size_t n_ws,
                          const uint8_t x[32], const uint8_t u[32], const uint8_t v[32], const uint8_t omega[32],
                          const uint8_t g1[64], uint8_t out_efwzw[256]);
/* Batch of independent proofs (BASELINE config 5): proof p uses queries [q_off[p], q_off[p+1]) and
 * witnesses [w_off[p], w_off[p+1]); xuv = n_proofs * (x||u||v).  All 4*n_proofs sums share one launch. */
int h2a_verify_accumulate_batch(h2a_ctx* ctx, size_t n_proofs, const uint8_t* commitments, const int32_t* rotations,
                                const uint8_t* evals, const size_t* q_off, const uint8_t* ws, const size_t* w_off,
                                const uint8_t* xuv, const uint8_t omega[32], const uint8_t g1[64],
                                uint8_t* out_efwzw /* n_proofs*256 */);
/* H = sum_i (x^n)^i h_i  (src/vanishing.rs:177-188) as one device MSM. */
int h2a_fold_h(h2a_ctx* ctx, const uint8_t* h_pieces /* m*64 */, size_t m, const uint8_t xn[32], uint8_t out_affine[64]);

/* ---- whole-proof verifier glue ------------------------------------------------------------------
 * Replaces the native side of `VerifierChip::_verify_proof` (src/verifier.rs:286-762): given the circuit
 * description the reference takes from the verifying key (:286-311), the vk commitments and the proof
 * bytes, replay the Blake2b transcript in the reference's wire order (:341-510), recompute the challenges
 * (theta, beta, gamma, y, x, v, u), l_0/l_last/l_blind (:513-591), the gate / permutation / lookup
 * expressions and the expected h(x) (src/vanishing.rs:145-175), assemble the query list (:654-715) and
 * evaluate (e, f, w, zw) for ALL proofs of a batch in one device launch.
 * `shape_words`: see csrc/plonk_shape.hpp for the word stream.  `vk_hash`: the transcript scalar derived
 * from `format!("{:?}", vk.pinned())` (src/verifier.rs:341-358) — an input, since the dependency's Debug
 * output cannot be reproduced outside it.  TRUST BOUNDARY: the library does not tie `vk_hash` to the shape or to the
 * fixed / sigma commitments; binding the Fiat-Shamir transcript to the circuit is the caller's job (pass the hash
 * of the verifying key these commitments belong to).  Proof encoding: 32-byte compressed points (x little-endian,
 * bit 255 = parity of y, identity = zeros) and 32-byte little-endian canonical scalars.
 * H2A_ERR_PROOF: truncated / trailing bytes, a point not on the curve, a non-canonical scalar. */
typedef struct h2a_circuit h2a_circuit;
int h2a_circuit_create(h2a_ctx* ctx, const uint32_t* shape_words, size_t n_words, const uint8_t* constants /* n*32 */,
                       size_t n_constants, h2a_circuit** out);
int h2a_circuit_free(h2a_ctx* ctx, h2a_circuit* circuit);
int h2a_circuit_set_vk(h2a_ctx* ctx, h2a_circuit* circuit, const uint8_t* fixed_commitments /* n_fixed*64 */,
                       const uint8_t* sigma_commitments /* n_perm*64 */, const uint8_t vk_hash[32]);
int h2a_verify_proof(h2a_ctx* ctx, const h2a_circuit* circuit, const uint8_t* instance_commitments /* n_instance*64 */,
                     const uint8_t* proof, size_t proof_len, uint8_t out_efwzw[256]);
int h2a_verify_proof_batch(h2a_ctx* ctx, const h2a_circuit* circuit, size_t n_proofs,
                           const uint8_t* instance_commitments /* n_proofs*n_instance*64 */, const uint8_t* const* proofs,
                           const size_t* proof_lens, uint8_t* out_efwzw /* n_proofs*256 */);

/* ---- prover pipeline ----------------------------------------------------------------------------
 * Replaces halo2 `plonk::create_proof(&params, &pk, &[circuit], &[&[&public_inputs]], &mut transcript)` as
 * called at examples/simple-example.rs:606-613 and :702-709, for a circuit given by its shape, its proving-key
 * columns and its (already synthesised and blinded) witness columns.  The proof bytes follow the wire order
 * `VerifierChip::_verify_proof` reads (src/verifier.rs:341-510, src/multiopen.rs:392).
 *   h2a_circuit_set_keys: `g` = Params.g, `g_lagrange` = Params.g_lagrange (resident handles, not owned);
 *     fixed_values = n_fixed columns of n elements, sigmas = n_perm permutation columns of n elements
 *     (keygen_pk output); commits them (so the verifying key is set as well) and keeps their coefficient and
 *     extended-coset forms on the device.  coset_shift = the generator of the evaluation coset (halo2: ZETA).
 *   h2a_create_proof: instance_cols / advice_cols = column-major, n elements each, advice already blinded.
 *     blinds = the prover's randomness, drawn by the caller, h2a_blinds_len(circuit) elements:
 *       per lookup: A' tail (bf+1), S' tail (bf+1), Z tail (bf); per permutation chunk: Z tail (bf);
 *       then the n coefficients of the random polynomial.
 *     Writes h2a_proof_len(circuit) bytes and the instance commitments.
 * H2A_ERR_INVALID when a lookup input is absent from its table (the only witness error the pipeline detects). */
int h2a_circuit_set_keys(h2a_ctx* ctx, h2a_circuit* circuit, const h2a_bases* g, const h2a_bases* g_lagrange,
                         const uint8_t* fixed_values, const uint8_t* sigmas, const uint8_t vk_hash[32],
                         const uint8_t coset_shift[32]);
int h2a_circuit_get_vk(h2a_ctx* ctx, const h2a_circuit* circuit, uint8_t* fixed_commitments, uint8_t* sigma_commitments);
size_t h2a_blinds_len(const h2a_circuit* circuit);
size_t h2a_proof_len(const h2a_circuit* circuit);
int h2a_create_proof(h2a_ctx* ctx, h2a_circuit* circuit, const uint8_t* instance_cols, const uint8_t* advice_cols,
                     const uint8_t* blinds, uint8_t* proof_out, size_t proof_cap, size_t* proof_len,
                     uint8_t* instance_commitments_out /* n_instance*64, may be NULL */);
/* Column-parallel proving over several GPUs (one process per GPU; every process calls h2a_create_proof with the SAME
 * inputs): each commitment batch of the prover is split round-robin — this process computes the MSMs of the
 * columns j with j % world == rank — and `exchange` must then fill in the others.  `exchange(user, buf, m)` receives
 * m * 64 bytes in which only this rank's columns are set (the rest zero) and must return with every column set by
 * its owner (rank j % world); an NCCL allgather of the buffers is the intended implementation.  All processes write
 * byte-identical proofs.  world = 1 or exchange = NULL switches the distribution off. */
typedef int (*h2a_exchange_fn)(void* user, uint8_t* commitments, size_t m);
int h2a_circuit_set_distribution(h2a_ctx* ctx, h2a_circuit* circuit, int rank, int world, h2a_exchange_fn exchange,
                                 void* user);

/* ---- several GPUs --------------------------------------------------------------------------------
 * One process per GPU, one ctx per process; the library carries its own NCCL plumbing (bound at run time from
 * libnccl.so.2), so a Rust host needs no other collective library.  Rank 0 draws two unique ids (h2a_comm_unique_id, twice)
 * and hands them to the other processes over any channel it has; every process then calls h2a_comm_init(ctx, rank, world,
 * id, id_bulk).  The first communicator carries the small exchanges on the ctx stream (64-byte MSM partials, commitments),
 * the second the bulk polynomial broadcasts of a proof spread over several GPUs.
 *   h2a_comm_allgather[_dev]: raw bytes, rank order (`recv` holds world * bytes_per_rank).
 *   h2a_msm_g1_sharded: one MSM whose points are split into per-rank ranges (each rank passes ITS bases and scalars): the
 *     64-byte affine partials are allgathered and summed in rank order, so every rank returns the same bytes.
 *   h2a_circuit_set_distribution(ctx, circuit, rank, world, NULL, NULL) on a ctx with a communicator spreads ONE proof over
 *     the ranks: commitments column-parallel, the transforms of a column on the rank that owns it (coefficient and extended
 *     forms broadcast from there), the quotient row-parallel (slices allgathered).  All ranks must be given the same inputs
 *     and all write the same proof bytes.  A witness error every rank can see (a lookup input outside its table) fails on every
 *     rank; a rank that fails for a reason of its own (a CUDA error, an allocation) leaves the others waiting in their next
 *     collective, as in any NCCL program: abort the job. */
int h2a_comm_unique_id(uint8_t out_id[128]);
int h2a_comm_init(h2a_ctx* ctx, int rank, int world, const uint8_t id[128], const uint8_t id_bulk[128]);
int h2a_comm_destroy(h2a_ctx* ctx);
/* Payload bytes this rank has contributed to / received from the library's collectives since h2a_comm_init (a broadcast
 * counts once at its root and once at every receiver). */
int h2a_comm_traffic(const h2a_ctx* ctx, uint64_t out_sent_received[2]);
int h2a_comm_rank(const h2a_ctx* ctx);
int h2a_comm_world(const h2a_ctx* ctx);
int h2a_comm_allgather(h2a_ctx* ctx, const uint8_t* send, uint8_t* recv, size_t bytes_per_rank);
int h2a_comm_allgather_dev(h2a_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);
int h2a_comm_broadcast_dev(h2a_ctx* ctx, void* d_buf, size_t bytes, int root);
int h2a_msm_g1_sharded(h2a_ctx* ctx, const h2a_bases* local_bases, size_t offset, const void* d_local_scalars, size_t n_local,
                       uint8_t out_affine[64]);

/* ---- Key generation: copy constraints -> sigma columns -------------------------------------
 * The permutation part of `keygen_vk` / `keygen_pk` (examples/simple-example.rs:593-594, :696-697).  An assembly
 * holds one permutation over the cells (col, row) of the n_cols permutation columns (`vk.permutation`,
 * src/verifier.rs:244-259), initially the identity; h2a_assembly_copy joins the cycles of two cells, as the layouter
 * does for every copy constraint (host bookkeeping).  h2a_assembly_sigmas evaluates on the device
 *   sigma[col][row] = delta^{col'} * omega^{row'},  (col', row') = the cell that follows (col, row) in its cycle,
 * (delta: `Fr::DELTA`, src/permutation.rs:259) and writes n_cols columns of 2^k elements — the `sigmas` argument of
 * h2a_circuit_set_keys.  h2a_assembly_mapping returns the permutation itself (next cell = col' * 2^k + row'). */
typedef struct h2a_assembly h2a_assembly;
int h2a_assembly_new(uint32_t n_cols, uint32_t k, h2a_assembly** out);
void h2a_assembly_free(h2a_assembly* assembly);
int h2a_assembly_copy(h2a_assembly* assembly, uint32_t col_a, uint32_t row_a, uint32_t col_b, uint32_t row_b);
int h2a_assembly_mapping(const h2a_assembly* assembly, uint32_t* out_next_cell /* n_cols * 2^k */);
int h2a_assembly_sigmas(h2a_ctx* ctx, const h2a_assembly* assembly, const uint8_t omega[32], const uint8_t delta[32],
                        uint8_t* out_sigmas /* n_cols * 2^k * 32 */);

/* KZG parameters on the device: g[i] = [s^i] G and g_lagrange[i] = [L_i(s)] G for i < 2^k, what
 * `Setup::<Bn256>::new(k, rng)` builds (examples/simple-example.rs:589, :687) once `rng` has produced the
 * secret `s` (an input here: how the dependency draws it from XorShiftRng is not visible from the reference).
 * The Lagrange basis uses the closed form L_i(s) = omega^i (s^n - 1) / (n (s - omega^i)) — fixed-base
 * multiplications only, no group FFT.  Returns two resident handles (free with h2a_bases_free). */
int h2a_kzg_setup(h2a_ctx* ctx, uint32_t k, const uint8_t s[32], h2a_bases** out_g, h2a_bases** out_g_lagrange);
/* The scalar that `XorShiftRng::from_seed(seed)` yields for the KZG secret (examples/simple-example.rs:584-589):
 * rand_xorshift 0.3 stream, 64 bytes, Fr::from_bytes_wide (the last step is upstream-inferred). Host only. */
int h2a_xorshift_scalar(const uint8_t seed[16], uint8_t out_scalar[32]);
/* The transcript scalar of the verifying key (src/verifier.rs:341-358): Blake2b-512 with personal "Halo2-Verify-Key" over the length
 * of `pinned_debug` as a little-endian u64 followed by its bytes, reduced with Fr::from_bytes_wide.  `pinned_debug` is the string the
 * Rust side formats with `format!("{:?}", vk.pinned())`: the dependency's Debug output cannot be produced outside it, the hashing can.
 * The result is the `vk_hash` argument of h2a_circuit_set_vk / h2a_circuit_set_keys.  Host only; len = 0 hashes the empty string. */
int h2a_vk_hash(const uint8_t* pinned_debug, size_t len, uint8_t out_scalar[32]);
/* Copy resident bases back to the host (n * 64 bytes), e.g. to write a params file. */
int h2a_bases_download(h2a_ctx* ctx, const h2a_bases* bases, uint8_t* out_affine_xy);
/* Parameter files — `Params::write(&mut file)` / `Params::read(file)` as used at examples/simple-example.rs:679-691 to cache
 * the k = 23 parameters, and `Setup::<Bn256>::verifier_params(&params, public_inputs_size)` (:590, :693).  The dependency's
 * byte format is not visible from the reference, so the format is this library's (csrc/params.cu has the layout): a 64-byte
 * header (magic "H2APARAM", version, k, flags), g[0..2^k), g_lagrange[0..2^k) as 64-byte in-memory points or — `compressed`
 * != 0 — as the proof's 32-byte encoding, an optional opaque 128-byte trailer (the caller's [s]G2), and a Blake2b-512 digest
 * of everything before it.  Points stream between the file and HBM through pinned staging buffers.  h2a_params_read checks the
 * digest and, on the device, that every point is on the curve (H2A_ERR_INVALID otherwise, nothing returned); it returns two
 * resident handles (free with h2a_bases_free).  `trailer128` may be NULL on both sides; *has_trailer says whether the file
 * carried one. */
int h2a_params_write(h2a_ctx* ctx, const char* path, uint32_t k, const h2a_bases* g, const h2a_bases* g_lagrange, int compressed,
                     const uint8_t* trailer128);
int h2a_params_read(h2a_ctx* ctx, const char* path, uint32_t* out_k, h2a_bases** out_g, h2a_bases** out_g_lagrange,
                    uint8_t* trailer128, int* has_trailer);
/* The verifier's parameters: a handle over the first `public_inputs_size` Lagrange bases, against which
 * `params_verifier.commit_lagrange(public_inputs)` (examples/simple-example.rs:638-640) is one h2a_msm_g1.  Shares the memory of
 * `g_lagrange`, which must outlive it. */
int h2a_params_verifier_view(h2a_ctx* ctx, const h2a_bases* g_lagrange, size_t public_inputs_size, h2a_bases** out);
/* Per-phase device times (ms) of the last h2a_create_proof; returns the number of phases written. */
int h2a_prove_phase_ms(h2a_ctx* ctx, const h2a_circuit* circuit, float* ms, int cap);
const char* h2a_prove_phase_name(const h2a_ctx* ctx, int index);

/* ---- aggregation-circuit witness generation: the non-native `mul_var` (row f4) ----------------------------
 * `ecc_chip.mul_var(region, point, scalar, offset)` is how the in-circuit verifier multiplies G1 points by transcript
 * scalars: src/multiopen.rs:393 (z_i W_i), :443 (Horner chain of commitments), :474,480,486 (W, ZW, F), :492 (E),
 * src/vanishing.rs:181-187 (quotient pieces) — about 37 per aggregated proof.  The chip (halo2wrong, not in the tree) holds an Fq
 * coordinate as 4 limbs of 68 bits in Fr cells (examples/simple-example.rs:396-397, packing as :535-548) and witnesses every Fq
 * product a*b = q*p + r with limb products.  h2a_mulvar_witness fills those cells for m independent (point, scalar) pairs at once
 * (one thread per pair walks the ladder, one thread per pair and step writes the records): out_results[i] = scalars[i] * points[i] (affine) and h2a_mulvar_witness_len() Fr elements per
 * pair in the layout documented in csrc/mulvar.cu (bits of the scalar, then per bit the records of one doubling and one
 * addition, limbs of the intermediate points, a final correction by -(2^254 aux)).  `aux` is the auxiliary point the incomplete
 * affine additions start from.  An entry whose ladder meets equal x coordinates (scalar 0, the identity as input, ...) cannot be
 * witnessed: the call returns H2A_ERR_INVALID after filling status_out[i] (0 = ok) for every entry; status_out may be NULL.
 * PARITY UNPINNED at the dependency boundary: the cell layout is this library's statement of the published algorithm; the
 * values are checked against the big-integer oracle (oracle/mulvar.py). */
size_t h2a_mulvar_witness_len(void);
int h2a_mulvar_witness(h2a_ctx* ctx, const uint8_t* points /* m*64 */, const uint8_t* scalars /* m*32 */, size_t m, const uint8_t aux[64],
                       uint8_t* out_results /* m*64 */, uint8_t* out_witness /* m*len*32, may be NULL */, uint32_t* status_out);
int h2a_mulvar_witness_dev(h2a_ctx* ctx, const void* d_points, const void* d_scalars, size_t m, const uint8_t aux[64], void* d_results,
                           void* d_witness, uint32_t* status_out);

/* Blake2b transcript with Challenge255 (src/transcript.rs:58,72,105-107,122-124). */
typedef struct h2a_transcript h2a_transcript;
h2a_transcript* h2a_transcript_new(void);
void h2a_transcript_free(h2a_transcript* t);
int h2a_transcript_common_point(h2a_transcript* t, const uint8_t point_affine[64]);
int h2a_transcript_common_scalar(h2a_transcript* t, const uint8_t scalar[32]);
int h2a_transcript_squeeze_challenge(h2a_transcript* t, uint8_t out_scalar[32]);

/* ---- synthetic inputs for benchmarks (written straight into device memory) -----------------
 * Counter-based and position-addressable: element i depends only on (seed, first + i).
 * Scalars are uniform in [0, r); bases are uniform curve points found by try-and-increment
 * (x from the stream, y = sqrt(x^3 + 3), sign from the stream). */
int h2a_gen_scalars_dev(h2a_ctx* ctx, uint64_t seed, size_t first, size_t n, void* d_out /* n*32 */);
int h2a_gen_bases_dev(h2a_ctx* ctx, uint64_t seed, size_t first, size_t n, void* d_out /* n*64 */);

/* ---- test / measurement hooks ---------------------------------------------------------- */
/* Element-wise device field arithmetic: field 0 = Fq, 1 = Fr; op 0 add 1 sub 2 mul 3 sqr 4 inverse by Fermat (cross-check) 5 neg
 * 6 canonical integer -> Montgomery form, 7 Montgomery form -> canonical integer, 8 inverse by division steps
 * (the one every kernel uses). */
int h2a_field_op(h2a_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* Test hook, no device needed: the spans of points half `half` (0/1) of the batched-affine addition tree reads and writes in
 * round `round` of `rounds`, for a tree of `total_padded` slots (csrc/tree_layout.hpp).  out6 = {in array, in first, in count,
 * out array, out first, out count}; arrays 0/1/2 are the three point arrays, -1 the sorted entries.  The halves run on two
 * unordered streams, so their spans must never meet before the join: tests/test_abi_cpu.py checks that. */
int h2a_tree_layout(uint64_t total_padded, int rounds, int half, int round, int64_t out6[6]);
/* Test hook, no device needed: the rows of an extended-domain column (m rows) that rank `rank` of `world` holds when one proof is
 * spread over several GPUs — its m / world rows of the quotient plus `halo` rows on either side, modulo m — as one or two
 * contiguous pieces out4 = {first, count, first, count} (csrc/dist_layout.hpp).  Returns the number of pieces, or H2A_ERR_INVALID
 * (world < 2, m not a multiple of world, 2 * halo > m / world: whole columns are broadcast then). */
int h2a_dist_window(uint32_t m, int world, int rank, uint32_t halo, uint32_t out4[4]);
/* Element-wise device point arithmetic: op 0: out[i] = a[i] + b[i]; op 1: out[i] = 2*a[i]. Affine in/out. */
int h2a_g1_op(h2a_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* Per-phase device times of the most recent MSM / NTT on this ctx, measured with CUDA events on the
 * ctx stream.  Enable with h2a_set_profiling(ctx, 1).  Returns the number of phases written (<= cap). */
int h2a_set_profiling(h2a_ctx* ctx, int on);
int h2a_last_phase_ms(h2a_ctx* ctx, float* ms, int cap);
const char* h2a_phase_name(int phase_kind /* 0 msm, 1 ntt */, int index);
/* Number of kernels this library has launched on the ctx since creation. */
uint64_t h2a_launch_count(const h2a_ctx* ctx);
/* Integer-pipe micro-benchmark: dependent-free IMAD chains; returns achieved 10^12 IMAD thread-instructions/s. */
int h2a_bench_imad(h2a_ctx* ctx, double* out_tera_imad_per_s);
/* Field-multiply micro-benchmark: returns 10^9 Fq Montgomery products per second. */
int h2a_bench_modmul(h2a_ctx* ctx, double* out_giga_mul_per_s);

#ifdef __cplusplus
}
#endif
#endif /* H2AGG_H */
