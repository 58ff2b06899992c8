//! Safe wrappers over `h2agg-sys`, named after the functions of the `halo2` dependency that the reference
//! (Trapdoor-Tech/halo2-aggregation) reaches its hot path through:
//!
//! | dependency item                                   | reached from the reference at                  | here |
//! |---------------------------------------------------|------------------------------------------------|------|
//! | `arithmetic::best_multiexp(coeffs, bases)`        | examples/simple-example.rs:638-640, :606-613   | [`best_multiexp`], [`Bases::msm`] |
//! | `arithmetic::best_fft(a, omega, log_n)`           | domain used at src/verifier.rs:252,431         | [`best_fft`] |
//! | `EvaluationDomain::{lagrange_to_coeff, ..}`       | inside `create_proof`                          | [`EvaluationDomain`] |
//! | `Setup::new`, `Params::{read,write}`, `verifier_params` | examples/simple-example.rs:589,679-693   | [`Params`] |
//! | `create_proof`, `verify_proof`                    | examples/simple-example.rs:606-626,702-728     | [`Circuit`] |
//! | Blake2b transcript                                | src/transcript.rs:66-129                       | [`Transcript`] |
//!
//! Field elements and points cross as raw bytes in their in-memory form (`bn256::Fr` = 4 x u64 Montgomery limbs,
//! `G1Affine` = x || y), so the fork passes `slice.as_ptr() as *const u8`; the generic parameters below are only
//! checked for size.  No Rust toolchain exists in the image this library is built in: this crate is source for the
//! maintainer, kept in step with the header by `tests/test_abi_cpu.py` (every `h2a_*` symbol used here must exist).
#![allow(clippy::missing_safety_doc)]
use h2agg_sys as sys;
use std::ffi::{CStr, CString};
use std::mem::size_of;
use std::os::raw::c_int;
use std::ptr;

#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
pub type Result<T> = std::result::Result<T, Error>;

/// One GPU.  Calls on one context are serialised; use one context per GPU (one process per GPU).
pub struct Context {
    raw: *mut sys::h2a_ctx,
}
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::h2a_init(&mut raw, device as c_int) };
        if rc != sys::H2A_OK {
            return Err(Error { code: rc, message: "h2a_init failed (no B200 / sm_100a device?)".into() });
        }
        Ok(Context { raw })
    }
    fn check(&self, rc: c_int) -> Result<()> {
        if rc == sys::H2A_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::h2a_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    }
    pub fn as_ptr(&self) -> *mut sys::h2a_ctx {
        self.raw
    }

    /// Several GPUs: rank 0 draws the ids with [`comm_unique_ids`] and hands them to every process.
    pub fn comm_init(&self, rank: i32, world: i32, ids: &[u8; 256]) -> Result<()> {
        self.check(unsafe { sys::h2a_comm_init(self.raw, rank, world, ids.as_ptr(), ids[128..].as_ptr()) })
    }
    /// Raw bytes of every rank, in rank order.
    pub fn allgather(&self, send: &[u8]) -> Result<Vec<u8>> {
        let world = unsafe { sys::h2a_comm_world(self.raw) } as usize;
        let mut out = vec![0u8; send.len() * world];
        self.check(unsafe { sys::h2a_comm_allgather(self.raw, send.as_ptr(), out.as_mut_ptr(), send.len()) })?;
        Ok(out)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::h2a_destroy(self.raw) };
    }
}

pub fn comm_unique_ids() -> Result<[u8; 256]> {
    let mut ids = [0u8; 256];
    for half in 0..2 {
        let rc = unsafe { sys::h2a_comm_unique_id(ids[128 * half..].as_mut_ptr()) };
        if rc != sys::H2A_OK {
            return Err(Error { code: rc, message: "h2a_comm_unique_id failed (libnccl.so.2 not found?)".into() });
        }
    }
    Ok(ids)
}

/// The transcript scalar of a verifying key (src/verifier.rs:341-358) from `format!("{:?}", vk.pinned())`: the `vk_hash`
/// argument of [`Circuit::set_keys`] / [`Circuit::set_vk`].  Host only.
pub fn vk_hash(pinned_debug: &str) -> [u8; 32] {
    let mut out = [0u8; 32];
    let rc = unsafe { sys::h2a_vk_hash(pinned_debug.as_ptr(), pinned_debug.len(), out.as_mut_ptr()) };
    assert_eq!(rc, sys::H2A_OK);
    out
}

fn bytes_of<T>(s: &[T], elem: usize) -> *const u8 {
    assert_eq!(size_of::<T>(), elem, "element is not {} bytes", elem);
    s.as_ptr() as *const u8
}

/// `Params.g` / `Params.g_lagrange` resident in HBM.
pub struct Bases<'c> {
    ctx: &'c Context,
    raw: *mut sys::h2a_bases,
}
impl<'c> Bases<'c> {
    /// `points`: `&[G1Affine]` laid out as 64-byte x || y (identity = zeros).
    pub fn upload<P>(ctx: &'c Context, points: &[P]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { sys::h2a_bases_upload(ctx.raw, bytes_of(points, 64), points.len(), &mut raw) })?;
        Ok(Bases { ctx, raw })
    }
    pub fn len(&self) -> usize {
        unsafe { sys::h2a_bases_len(self.raw) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    /// Window tables 2^(bits*w) * P_i (once per parameters; -1 = the library's choice for single large MSMs, 17 for provers).
    pub fn precompute(&mut self, window_bits: i32) -> Result<()> {
        self.ctx.check(unsafe { sys::h2a_bases_precompute(self.ctx.raw, self.raw, window_bits) })
    }
    /// `best_multiexp(coeffs, &bases[..coeffs.len()])` -> 64-byte affine point.
    pub fn msm<S>(&self, coeffs: &[S]) -> Result<[u8; 64]> {
        let mut out = [0u8; 64];
        self.ctx.check(unsafe { sys::h2a_msm_g1(self.ctx.raw, self.raw, 0, bytes_of(coeffs, 32), coeffs.len(), out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// Several polynomials over the same parameters in one call (the rounds of `create_proof`).
    pub fn msm_batch<S>(&self, columns: &[&[S]]) -> Result<Vec<[u8; 64]>> {
        let ptrs: Vec<*const u8> = columns.iter().map(|c| bytes_of(c, 32)).collect();
        let lens: Vec<usize> = columns.iter().map(|c| c.len()).collect();
        let mut out = vec![[0u8; 64]; columns.len()];
        self.ctx.check(unsafe {
            sys::h2a_msm_g1_batch(self.ctx.raw, self.raw, ptrs.as_ptr(), lens.as_ptr(), columns.len() as c_int, out.as_mut_ptr() as *mut u8)
        })?;
        Ok(out)
    }
    pub fn as_ptr(&self) -> *const sys::h2a_bases {
        self.raw
    }
}
impl Drop for Bases<'_> {
    fn drop(&mut self) {
        unsafe { sys::h2a_bases_free(self.ctx.raw, self.raw) };
    }
}

/// `halo2::arithmetic::best_multiexp` for bases that are not `Params` vectors (the verifier's sums).
pub fn best_multiexp<S, P>(ctx: &Context, coeffs: &[S], bases: &[P]) -> Result<[u8; 64]> {
    assert_eq!(coeffs.len(), bases.len());
    let mut out = [0u8; 64];
    ctx.check(unsafe { sys::h2a_msm_g1_adhoc(ctx.raw, bytes_of(bases, 64), bytes_of(coeffs, 32), coeffs.len(), out.as_mut_ptr()) })?;
    Ok(out)
}

/// `halo2::arithmetic::best_fft(a, omega, log_n)`, in place, natural order.
pub fn best_fft<S>(ctx: &Context, a: &mut [S], omega: &S, log_n: u32) -> Result<()> {
    assert_eq!(a.len(), 1usize << log_n);
    assert_eq!(size_of::<S>(), 32);
    ctx.check(unsafe { sys::h2a_ntt(ctx.raw, a.as_mut_ptr() as *mut u8, log_n, omega as *const S as *const u8, 0, ptr::null()) })
}

/// The transforms of `halo2::poly::EvaluationDomain` (k rows, extended domain 2^ext_k, coset generator `zeta`).
pub struct EvaluationDomain<'c> {
    ctx: &'c Context,
    pub k: u32,
    pub ext_k: u32,
    quotient_poly_degree: usize,
    omega: [u8; 32],
    zeta: [u8; 32],
}
impl<'c> EvaluationDomain<'c> {
    /// `j` = the circuit's degree, as in `EvaluationDomain::new(j, k)`; `zeta` = the dependency's coset generator.
    pub fn new(ctx: &'c Context, j: u32, k: u32, zeta: [u8; 32]) -> Result<Self> {
        assert!(j >= 2);
        let mut ext_k = k;
        while (1u64 << ext_k) < ((1u64 << k) * (j as u64 - 1)) {
            ext_k += 1;
        }
        let mut omega = [0u8; 32];
        ctx.check(unsafe { sys::h2a_fr_root_of_unity(k, omega.as_mut_ptr()) })?;
        Ok(EvaluationDomain { ctx, k, ext_k, quotient_poly_degree: (j - 1) as usize, omega, zeta })
    }
    pub fn get_omega(&self) -> [u8; 32] {
        self.omega
    }
    /// `j - 1`, as the dependency returns it (src/verifier.rs:431 reads it for the number of quotient pieces).
    pub fn get_quotient_poly_degree(&self) -> usize {
        self.quotient_poly_degree
    }
    pub fn lagrange_to_coeff<S>(&self, a: &mut [S]) -> Result<()> {
        assert_eq!(a.len(), 1usize << self.k);
        self.ctx.check(unsafe { sys::h2a_ntt(self.ctx.raw, a.as_mut_ptr() as *mut u8, self.k, self.omega.as_ptr(), 1, ptr::null()) })
    }
    pub fn coeff_to_extended<S: Clone + Default>(&self, coeffs: &[S]) -> Result<Vec<S>> {
        assert_eq!(size_of::<S>(), 32);
        let mut out = vec![S::default(); 1usize << self.ext_k];
        self.ctx.check(unsafe {
            sys::h2a_coeff_to_extended(self.ctx.raw, bytes_of(coeffs, 32), self.k, self.ext_k, self.zeta.as_ptr(), out.as_mut_ptr() as *mut u8)
        })?;
        Ok(out)
    }
    pub fn extended_to_coeff<S>(&self, ext: &mut [S]) -> Result<()> {
        assert_eq!(ext.len(), 1usize << self.ext_k);
        self.ctx.check(unsafe { sys::h2a_extended_to_coeff(self.ctx.raw, ext.as_mut_ptr() as *mut u8, self.ext_k, self.zeta.as_ptr()) })
    }
}

/// `Params<G1Affine>`: `g`, `g_lagrange` resident on the GPU.
pub struct Params<'c> {
    pub k: u32,
    pub g: Bases<'c>,
    pub g_lagrange: Bases<'c>,
    /// The opaque 128-byte trailer of a parameter file ([s]G2 for the pairing check, owned by the caller).
    pub trailer: Option<[u8; 128]>,
}
impl<'c> Params<'c> {
    /// `Setup::<Bn256>::new(k, rng)` once `rng` has produced the secret `s` (examples/simple-example.rs:589,687).
    pub fn setup(ctx: &'c Context, k: u32, s: &[u8; 32]) -> Result<Self> {
        let (mut g, mut gl) = (ptr::null_mut(), ptr::null_mut());
        ctx.check(unsafe { sys::h2a_kzg_setup(ctx.raw, k, s.as_ptr(), &mut g, &mut gl) })?;
        Ok(Params { k, g: Bases { ctx, raw: g }, g_lagrange: Bases { ctx, raw: gl }, trailer: None })
    }
    /// `Params::read` (examples/simple-example.rs:681-684); the file format is the library's (csrc/params.cu).
    pub fn read(ctx: &'c Context, path: &str) -> Result<Self> {
        let cpath = CString::new(path).unwrap();
        let (mut k, mut g, mut gl, mut has) = (0u32, ptr::null_mut(), ptr::null_mut(), 0 as c_int);
        let mut trailer = [0u8; 128];
        ctx.check(unsafe { sys::h2a_params_read(ctx.raw, cpath.as_ptr(), &mut k, &mut g, &mut gl, trailer.as_mut_ptr(), &mut has) })?;
        Ok(Params { k, g: Bases { ctx, raw: g }, g_lagrange: Bases { ctx, raw: gl }, trailer: if has != 0 { Some(trailer) } else { None } })
    }
    /// `Params::write` (examples/simple-example.rs:686-690).
    pub fn write(&self, path: &str, compressed: bool) -> Result<()> {
        let cpath = CString::new(path).unwrap();
        let tr = self.trailer.as_ref().map_or(ptr::null(), |t| t.as_ptr());
        self.g.ctx.check(unsafe { sys::h2a_params_write(self.g.ctx.raw, cpath.as_ptr(), self.k, self.g.raw, self.g_lagrange.raw, compressed as c_int, tr) })
    }
    /// `Setup::verifier_params(&params, public_inputs_size)` (:590, :693): the bases `commit_lagrange(public_inputs)` needs.
    pub fn verifier_params(&self, public_inputs_size: usize) -> Result<Bases<'c>> {
        let mut raw = ptr::null_mut();
        self.g.ctx.check(unsafe { sys::h2a_params_verifier_view(self.g.ctx.raw, self.g_lagrange.raw, public_inputs_size, &mut raw) })?;
        Ok(Bases { ctx: self.g.ctx, raw })
    }
    pub fn commit<S>(&self, poly_coeffs: &[S]) -> Result<[u8; 64]> {
        self.g.msm(poly_coeffs)
    }
    pub fn commit_lagrange<S>(&self, poly_evals: &[S]) -> Result<[u8; 64]> {
        self.g_lagrange.msm(poly_evals)
    }
}

/// A circuit as `VerifierChip::_verify_proof` sees it through the verifying key (src/verifier.rs:233-283), serialised as the
/// word stream of csrc/plonk_shape.hpp; `create_proof` / `verify_proof` for it.
pub struct Circuit<'c> {
    ctx: &'c Context,
    raw: *mut sys::h2a_circuit,
}
impl<'c> Circuit<'c> {
    pub fn new(ctx: &'c Context, shape_words: &[u32], constants: &[[u8; 32]]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { sys::h2a_circuit_create(ctx.raw, shape_words.as_ptr(), shape_words.len(), constants.as_ptr() as *const u8, constants.len(), &mut raw) })?;
        Ok(Circuit { ctx, raw })
    }
    /// `keygen_pk`'s output: fixed columns and permutation columns (n elements each), committed and kept on the device.
    pub fn set_keys<S>(&mut self, params: &Params<'c>, fixed: &[S], sigmas: &[S], vk_hash: &[u8; 32], zeta: &[u8; 32]) -> Result<()> {
        self.ctx.check(unsafe {
            sys::h2a_circuit_set_keys(self.ctx.raw, self.raw, params.g.raw, params.g_lagrange.raw, bytes_of(fixed, 32), bytes_of(sigmas, 32), vk_hash.as_ptr(), zeta.as_ptr())
        })
    }
    pub fn set_vk(&mut self, fixed_commitments: &[[u8; 64]], sigma_commitments: &[[u8; 64]], vk_hash: &[u8; 32]) -> Result<()> {
        self.ctx.check(unsafe {
            sys::h2a_circuit_set_vk(self.ctx.raw, self.raw, fixed_commitments.as_ptr() as *const u8, sigma_commitments.as_ptr() as *const u8, vk_hash.as_ptr())
        })
    }
    /// One proof over the ranks of the context's communicator ([`Context::comm_init`]).
    pub fn distribute(&mut self, rank: i32, world: i32) -> Result<()> {
        self.ctx.check(unsafe { sys::h2a_circuit_set_distribution(self.ctx.raw, self.raw, rank, world, None, ptr::null_mut()) })
    }
    pub fn blinds_len(&self) -> usize {
        unsafe { sys::h2a_blinds_len(self.raw) }
    }
    /// `create_proof(&params, &pk, &[circuit], &[&[&public_inputs]], &mut transcript)`: returns `transcript.finalize()`.
    pub fn create_proof<S>(&mut self, instance_cols: &[S], advice_cols: &[S], blinds: &[S]) -> Result<Vec<u8>> {
        assert_eq!(blinds.len(), self.blinds_len());
        let cap = unsafe { sys::h2a_proof_len(self.raw) };
        let (mut proof, mut len) = (vec![0u8; cap], 0usize);
        self.ctx.check(unsafe {
            sys::h2a_create_proof(self.ctx.raw, self.raw, bytes_of(instance_cols, 32), bytes_of(advice_cols, 32), bytes_of(blinds, 32), proof.as_mut_ptr(), cap, &mut len, ptr::null_mut())
        })?;
        proof.truncate(len);
        Ok(proof)
    }
    /// `verify_proof(&params_verifier, vk, instances, &mut transcript)` up to the pairing: `[e, f, w, zw]`
    /// (examples/simple-example.rs:620, :668-671; src/verifier.rs:739-742).
    pub fn verify_proof(&self, instance_commitments: &[[u8; 64]], proof: &[u8]) -> Result<[[u8; 64]; 4]> {
        let mut out = [[0u8; 64]; 4];
        self.ctx.check(unsafe { sys::h2a_verify_proof(self.ctx.raw, self.raw, instance_commitments.as_ptr() as *const u8, proof.as_ptr(), proof.len(), out.as_mut_ptr() as *mut u8) })?;
        Ok(out)
    }
    /// A batch of independent proofs of this circuit in one launch (BASELINE config 5).
    pub fn verify_proof_batch(&self, instance_commitments: &[[u8; 64]], proofs: &[&[u8]]) -> Result<Vec<[[u8; 64]; 4]>> {
        let ptrs: Vec<*const u8> = proofs.iter().map(|p| p.as_ptr()).collect();
        let lens: Vec<usize> = proofs.iter().map(|p| p.len()).collect();
        let mut out = vec![[[0u8; 64]; 4]; proofs.len()];
        self.ctx.check(unsafe {
            sys::h2a_verify_proof_batch(self.ctx.raw, self.raw, proofs.len(), instance_commitments.as_ptr() as *const u8, ptrs.as_ptr(), lens.as_ptr(), out.as_mut_ptr() as *mut u8)
        })?;
        Ok(out)
    }
}
impl Drop for Circuit<'_> {
    fn drop(&mut self) {
        unsafe { sys::h2a_circuit_free(self.ctx.raw, self.raw) };
    }
}

/// The multi-open accumulation alone (`MultiopenChip::calc_witness`, src/multiopen.rs:271-509) for callers that keep their own
/// transcript replay: returns `[e, f, w, zw]`.
#[allow(clippy::too_many_arguments)]
pub fn verify_accumulate(ctx: &Context, commitments: &[[u8; 64]], rotations: &[i32], evals: &[[u8; 32]], ws: &[[u8; 64]], x: &[u8; 32],
                         u: &[u8; 32], v: &[u8; 32], omega: &[u8; 32], g1: &[u8; 64]) -> Result<[[u8; 64]; 4]> {
    assert!(commitments.len() == rotations.len() && rotations.len() == evals.len());
    let mut out = [[0u8; 64]; 4];
    ctx.check(unsafe {
        sys::h2a_verify_accumulate(ctx.raw, commitments.as_ptr() as *const u8, rotations.as_ptr(), evals.as_ptr() as *const u8, rotations.len(),
                                   ws.as_ptr() as *const u8, ws.len(), x.as_ptr(), u.as_ptr(), v.as_ptr(), omega.as_ptr(), g1.as_ptr(), out.as_mut_ptr() as *mut u8)
    })?;
    Ok(out)
}

/// `H = sum_i (x^n)^i h_i` (src/vanishing.rs:177-188).
pub fn fold_h(ctx: &Context, h_pieces: &[[u8; 64]], xn: &[u8; 32]) -> Result<[u8; 64]> {
    let mut out = [0u8; 64];
    ctx.check(unsafe { sys::h2a_fold_h(ctx.raw, h_pieces.as_ptr() as *const u8, h_pieces.len(), xn.as_ptr(), out.as_mut_ptr()) })?;
    Ok(out)
}

/// Witness cells of the non-native `ecc_chip.mul_var(region, point, scalar, offset)` of the aggregation circuit
/// (src/multiopen.rs:393-492, src/vanishing.rs:181-187) for a batch of (point, scalar) pairs.
pub struct MulVarWitness {
    /// `scalars[i] * points[i]`, affine.
    pub results: Vec<[u8; 64]>,
    /// `mul_var_witness_len()` cells per pair (layout: csrc/mulvar.cu); empty when not asked for.
    pub cells: Vec<[u8; 32]>,
    /// 0 = witnessed; anything else: the incomplete additions met equal x coordinates (scalar 0, the identity, ...).
    pub status: Vec<u32>,
}
pub fn mul_var_witness_len() -> usize {
    unsafe { sys::h2a_mulvar_witness_len() }
}
/// `aux`: the auxiliary point the incomplete additions start from.  `Err` carries the library's message; when single
/// entries cannot be witnessed the error code is `H2A_ERR_INVALID` and `status` of the returned error context names them.
pub fn mul_var_witness(ctx: &Context, points: &[[u8; 64]], scalars: &[[u8; 32]], aux: &[u8; 64], want_cells: bool) -> std::result::Result<MulVarWitness, (Error, Vec<u32>)> {
    assert_eq!(points.len(), scalars.len());
    let m = points.len();
    let mut w = MulVarWitness { results: vec![[0u8; 64]; m], cells: vec![[0u8; 32]; if want_cells { m * mul_var_witness_len() } else { 0 }], status: vec![0u32; m] };
    let cells_ptr = if want_cells { w.cells.as_mut_ptr() as *mut u8 } else { ptr::null_mut() };
    let rc = unsafe {
        sys::h2a_mulvar_witness(ctx.raw, points.as_ptr() as *const u8, scalars.as_ptr() as *const u8, m, aux.as_ptr(), w.results.as_mut_ptr() as *mut u8, cells_ptr, w.status.as_mut_ptr())
    };
    match ctx.check(rc) {
        Ok(()) => Ok(w),
        Err(e) => Err((e, w.status)),
    }
}

/// `Blake2bWrite<_, _, Challenge255<_>>` / `Blake2bRead` (src/transcript.rs:58,72,105-107,122-124).
pub struct Transcript {
    raw: *mut sys::h2a_transcript,
}
impl Transcript {
    pub fn new() -> Self {
        Transcript { raw: unsafe { sys::h2a_transcript_new() } }
    }
    pub fn common_point(&mut self, p: &[u8; 64]) -> bool {
        unsafe { sys::h2a_transcript_common_point(self.raw, p.as_ptr()) == sys::H2A_OK }
    }
    pub fn common_scalar(&mut self, s: &[u8; 32]) {
        unsafe { sys::h2a_transcript_common_scalar(self.raw, s.as_ptr()) };
    }
    pub fn squeeze_challenge(&mut self) -> [u8; 32] {
        let mut out = [0u8; 32];
        unsafe { sys::h2a_transcript_squeeze_challenge(self.raw, out.as_mut_ptr()) };
        out
    }
}
impl Default for Transcript {
    fn default() -> Self {
        Self::new()
    }
}
impl Drop for Transcript {
    fn drop(&mut self) {
        unsafe { sys::h2a_transcript_free(self.raw) };
    }
}
