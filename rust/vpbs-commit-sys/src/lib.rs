//! `extern "C"` declarations for libvpbs_commit.so — one-to-one with include/vpbs_commit.h —
//! plus a small safe wrapper (`Ctx`, `commit`) shaped for plonky2's `PolynomialBatch::from_values`
//! / `from_coeffs` (plonky2 0.2.0 `src/fri/oracle.rs`), which the reference reaches from
//! `prove()` at src/vtfhe/ivc_based_vpbs.rs:302, :333, :364 and `build()` at :275.
//!
//! NOT compiled in this repository's environment (no Rust toolchain there).
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_uint, c_void, CStr};

#[repr(C)]
pub struct vpbs_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct vpbs_batch {
    _private: [u8; 0],
}

#[repr(C)]
pub struct vpbs_sigmas {
    _private: [u8; 0],
}

#[repr(C)]
pub struct vpbs_fri {
    _private: [u8; 0],
}

#[repr(C)]
pub struct vpbs_gate_program {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct vpbs_stats {
    pub h2d_ms: f32,
    pub ifft_ms: f32,
    pub fft_ms: f32,
    pub merkle_ms: f32,
    pub leaf_hash_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
    pub kernel_launches: u64,
}

pub const VPBS_OK: c_int = 0;
pub const VPBS_ERR_ARG: c_int = -1;
pub const VPBS_SALT_SIZE: usize = 4;

extern "C" {
    pub fn vpbs_abi_version() -> c_int;
    pub fn vpbs_device_count() -> c_int;
    pub fn vpbs_ctx_create(device: c_int, out: *mut *mut vpbs_ctx) -> c_int;
    pub fn vpbs_ctx_destroy(ctx: *mut vpbs_ctx);
    pub fn vpbs_ctx_set_stream(ctx: *mut vpbs_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn vpbs_ctx_sync(ctx: *mut vpbs_ctx) -> c_int;
    pub fn vpbs_last_error(ctx: *mut vpbs_ctx) -> *const c_char;
    pub fn vpbs_ctx_kernel_launches(ctx: *mut vpbs_ctx) -> u64;
    pub fn vpbs_ctx_set_host_threads(ctx: *mut vpbs_ctx, threads: c_uint) -> c_int;
    pub fn vpbs_ctx_set_shard(ctx: *mut vpbs_ctx, index: u32, count: u32) -> c_int;
    pub fn vpbs_host_alloc(bytes: usize) -> *mut c_void;
    pub fn vpbs_host_free(p: *mut c_void);
    pub fn vpbs_fft(ctx: *mut vpbs_ctx, inout: *mut u64, log_n: u32) -> c_int;
    pub fn vpbs_ifft(ctx: *mut vpbs_ctx, inout: *mut u64, log_n: u32) -> c_int;
    pub fn vpbs_coset_fft(ctx: *mut vpbs_ctx, inout: *mut u64, log_n: u32, shift: u64) -> c_int;
    pub fn vpbs_poseidon_permute(ctx: *mut vpbs_ctx, states: *mut u64, count: u64) -> c_int;
    pub fn vpbs_hash_or_noop_batch(ctx: *mut vpbs_ctx, rows: *const u64, count: u64, len: u32,
                                   hashes_out: *mut u64) -> c_int;
    pub fn vpbs_two_to_one_batch(ctx: *mut vpbs_ctx, left: *const u64, right: *const u64,
                                 count: u64, hashes_out: *mut u64) -> c_int;
    pub fn vpbs_merkle_new(ctx: *mut vpbs_ctx, leaves_rowmajor: *const u64, nleaves: u64,
                           leaf_len: u32, cap_height: u32, digests_out: *mut u64,
                           cap_out: *mut u64) -> c_int;
    pub fn vpbs_merkle_new_dev(ctx: *mut vpbs_ctx, d_leaves: *const u64, nleaves: u64, leaf_len: u32,
                               cap_height: u32, d_digests_out: *mut u64, d_cap_out: *mut u64,
                               stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_lde_batch(ctx: *mut vpbs_ctx, cols: *const *const u64, ncols: u32, log_n: u32,
                          rate_bits: u32, inputs_are_coeffs: c_int, coeffs_out: *const *mut u64,
                          lde_cols_out: *mut u64) -> c_int;
    pub fn vpbs_commit(ctx: *mut vpbs_ctx, cols: *const *const u64, ncols: u32, log_n: u32,
                       rate_bits: u32, cap_height: u32, inputs_are_coeffs: c_int,
                       salt_cols: *const *const u64, coeffs_out: *const *mut u64,
                       leaves_out: *mut u64, digests_out: *mut u64, cap_out: *mut u64,
                       stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_commit_multi(ctxs: *const *mut vpbs_ctx, nctx: c_int, cols: *const *const u64,
                             ncols: u32, log_n: u32, rate_bits: u32, cap_height: u32,
                             inputs_are_coeffs: c_int, salt_cols: *const *const u64,
                             coeffs_out: *const *mut u64, leaves_out: *mut u64,
                             digests_out: *mut u64, cap_out: *mut u64,
                             stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_commit_dev(ctx: *mut vpbs_ctx, d_cols: *const u64, ncols: u32, log_n: u32,
                           rate_bits: u32, cap_height: u32, inputs_are_coeffs: c_int,
                           d_salt: *const u64, d_coeffs_out: *mut u64, d_leaves_out: *mut u64,
                           d_digests_out: *mut u64, d_cap_out: *mut u64,
                           stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_commit_shard_dev(ctx: *mut vpbs_ctx, d_cols: *const u64, ncols: u32, log_n: u32,
                                 rate_bits: u32, cap_height: u32, inputs_are_coeffs: c_int,
                                 first_leaf: u64, nleaves_shard: u64, d_coeffs_out: *mut u64,
                                 d_leaves_out: *mut u64, d_digests_out: *mut u64,
                                 d_roots_out: *mut u64, stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_eval_ext2(ctx: *mut vpbs_ctx, coeff_cols: *const *const u64, ncols: u32, log_n: u32,
                          points: *const u64, npoints: u32, out: *mut u64) -> c_int;
    pub fn vpbs_batch_eval_ext2(batch: *mut vpbs_batch, points: *const u64, npoints: u32,
                                out: *mut u64) -> c_int;
    pub fn vpbs_fri_layer_commit(ctx: *mut vpbs_ctx, values_ext: *const u64, len: u64, arity_bits: u32,
                                 cap_height: u32, leaves_out: *mut u64, digests_out: *mut u64,
                                 cap_out: *mut u64) -> c_int;
    pub fn vpbs_fri_fold(ctx: *mut vpbs_ctx, coeffs_ext: *const u64, len: u64, arity_bits: u32,
                         beta: *const u64, shift_next: u64, coeffs_out: *mut u64,
                         values_out: *mut u64) -> c_int;
    // FRI commit phase as one device-resident chain
    pub fn vpbs_fri_begin(ctx: *mut vpbs_ctx, final_poly_coeffs_ext: *const u64, ncoeffs: u64,
                          rate_bits: u32, out: *mut *mut vpbs_fri) -> c_int;
    pub fn vpbs_fri_begin_openings(ctx: *mut vpbs_ctx, oracles: *const *mut vpbs_batch, noracles: u32,
                                   batch_sizes: *const u32, nbatches: u32, poly_refs: *const u32,
                                   points: *const u64, alpha: *const u64, rate_bits: u32,
                                   out: *mut *mut vpbs_fri) -> c_int;
    pub fn vpbs_fri_commit_layer(fri: *mut vpbs_fri, arity_bits: u32, cap_height: u32,
                                 cap_out: *mut u64) -> c_int;
    pub fn vpbs_fri_fold_layer(fri: *mut vpbs_fri, beta: *const u64) -> c_int;
    pub fn vpbs_fri_final_poly(fri: *mut vpbs_fri, rate_bits: u32, coeffs_out: *mut u64) -> c_int;
    pub fn vpbs_fri_query_layer(fri: *mut vpbs_fri, layer: u32, leaf_indices: *const u64, count: u64,
                                rows_out: *mut u64, siblings_out: *mut u64) -> c_int;
    pub fn vpbs_fri_destroy(fri: *mut vpbs_fri);
    pub fn vpbs_pow_grind(ctx: *mut vpbs_ctx, state: *const u64, witness_pos: u32, response_lane: u32,
                          min_leading_zeros: u32, first_candidate: u64, count: u64,
                          witness_out: *mut u64, found: *mut c_int) -> c_int;
    // device-resident batches: only the cap crosses PCIe at commit time; rows / paths on demand
    pub fn vpbs_batch_commit(ctx: *mut vpbs_ctx, cols: *const *const u64, ncols: u32, log_n: u32,
                             rate_bits: u32, cap_height: u32, inputs_are_coeffs: c_int,
                             salt_cols: *const *const u64, cap_out: *mut u64,
                             out: *mut *mut vpbs_batch, stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_batch_destroy(batch: *mut vpbs_batch);
    pub fn vpbs_batch_get_leaves(batch: *mut vpbs_batch, leaf_indices: *const u64, count: u64,
                                 rows_out: *mut u64) -> c_int;
    pub fn vpbs_batch_prove(batch: *mut vpbs_batch, leaf_indices: *const u64, count: u64,
                            siblings_out: *mut u64) -> c_int;
    pub fn vpbs_batch_download(batch: *mut vpbs_batch, coeffs_out: *const *mut u64,
                               leaves_out: *mut u64, digests_out: *mut u64) -> c_int;
    pub fn vpbs_batch_commit_dev(ctx: *mut vpbs_ctx, d_cols: *const u64, ncols: u32, log_n: u32,
                                 rate_bits: u32, cap_height: u32, inputs_are_coeffs: c_int,
                                 cap_out: *mut u64, out: *mut *mut vpbs_batch,
                                 stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_batch_get_lde_rows(batch: *mut vpbs_batch, first_index: u64, step: u64, count: u64,
                                   rows_out: *mut u64) -> c_int;
    // permutation argument (prove() steps 4-5) on the device
    pub fn vpbs_sigmas_upload(ctx: *mut vpbs_ctx, sigma_cols: *const *const u64, k_is: *const u64,
                              num_routed: u32, log_n: u32, out: *mut *mut vpbs_sigmas) -> c_int;
    pub fn vpbs_sigmas_destroy(sigmas: *mut vpbs_sigmas);
    pub fn vpbs_zs_partial_products(ctx: *mut vpbs_ctx, wire_cols: *const *const u64,
                                    sigmas: *const vpbs_sigmas, max_degree: u32, betas: *const u64,
                                    gammas: *const u64, num_challenges: u32,
                                    cols_out: *const *mut u64) -> c_int;
    pub fn vpbs_batch_zs_partial_products(wires: *mut vpbs_batch, sigmas: *const vpbs_sigmas,
                                          max_degree: u32, betas: *const u64, gammas: *const u64,
                                          num_challenges: u32, rate_bits: u32, cap_height: u32,
                                          cap_out: *mut u64, out: *mut *mut vpbs_batch,
                                          stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_batch_shape(batch: *mut vpbs_batch, ncols: *mut u32, log_n: *mut u32,
                            rate_bits: *mut u32, cap_height: *mut u32, width: *mut u32) -> c_int;
    pub fn vpbs_batch_quotient_polys(constants_sigmas: *mut vpbs_batch, sigmas_first_col: u32,
                                     wires: *mut vpbs_batch, zs_pp: *mut vpbs_batch, k_is: *const u64,
                                     num_routed: u32, max_degree: u32, quotient_degree_bits: u32,
                                     betas: *const u64, gammas: *const u64, alphas: *const u64,
                                     num_challenges: u32, gate_terms: *const *const u64,
                                     program: *const vpbs_gate_program, public_inputs_hash: *const u64,
                                     rate_bits: u32, cap_height: u32, cap_out: *mut u64,
                                     out: *mut *mut vpbs_batch, stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_batch_quotient_values(constants_sigmas: *mut vpbs_batch, sigmas_first_col: u32,
                                      wires: *mut vpbs_batch, zs_pp: *mut vpbs_batch, k_is: *const u64,
                                      num_routed: u32, max_degree: u32, quotient_degree_bits: u32,
                                      betas: *const u64, gammas: *const u64, alphas: *const u64,
                                      num_challenges: u32, gate_terms: *const *const u64,
                                      program: *const vpbs_gate_program, public_inputs_hash: *const u64,
                                      d_vals_out: *mut u64) -> c_int;
    pub fn vpbs_quotient_commit_values(ctx: *mut vpbs_ctx, d_vals: *const u64, num_challenges: u32,
                                       log_n: u32, quotient_degree_bits: u32, rate_bits: u32,
                                       cap_height: u32, cap_out: *mut u64, out: *mut *mut vpbs_batch,
                                       stats: *mut vpbs_stats) -> c_int;
    pub fn vpbs_gate_program_upload(ctx: *mut vpbs_ctx, code: *const u64, ncode: u32, imms: *const u64,
                                    nimm: u32, nregs: u32, num_constraints: u32,
                                    out: *mut *mut vpbs_gate_program) -> c_int;
    pub fn vpbs_gate_program_destroy(program: *mut vpbs_gate_program);
    pub fn vpbs_batch_shard(batch: *mut vpbs_batch, first_leaf: *mut u64, nleaves: *mut u64) -> c_int;
    pub fn vpbs_batches_eval_ext2(batches: *const *mut vpbs_batch, nbatches: u32, points: *const u64,
                                  npoints: u32, outs: *const *mut u64) -> c_int;
    pub fn vpbs_batches_open(batches: *const *mut vpbs_batch, nbatches: u32, leaf_indices: *const u64,
                             count: u64, rows_out: *const *mut u64, siblings_out: *const *mut u64) -> c_int;
}

/// One device context (device arena + stream), reused across the 730 step proofs of a PBS.
pub struct Ctx(*mut vpbs_ctx);
unsafe impl Send for Ctx {}

impl Ctx {
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { vpbs_ctx_create(device, &mut h) };
        if rc != VPBS_OK {
            let msg = unsafe { CStr::from_ptr(vpbs_last_error(std::ptr::null_mut())) };
            // plonky2 has no error path here either: prove(...).unwrap() (ivc_based_vpbs.rs:308)
            panic!("vpbs_ctx_create({device}) failed: {}", msg.to_string_lossy());
        }
        Ctx(h)
    }
    fn check(&self, rc: c_int) {
        if rc != VPBS_OK {
            let msg = unsafe { CStr::from_ptr(vpbs_last_error(self.0)) };
            panic!("vpbs error {rc}: {}", msg.to_string_lossy());
        }
    }
    /// Copy threads of the pinned staging ring that pageable columns (plonky2's own
    /// `Vec<PolynomialValues<F>>`) travel through; 0 leaves such copies to the CUDA driver.
    pub fn set_host_threads(&self, threads: u32) {
        self.check(unsafe { vpbs_ctx_set_host_threads(self.0, threads) })
    }
    /// Row-range sharding of one proof over `count` GPUs: resident batches created afterwards hold
    /// leaves [index * m / count, (index + 1) * m / count); their caps carry only the own entries
    /// (the rest zero) and are completed by an all-gather across the shards.
    pub fn set_shard(&self, index: u32, count: u32) {
        self.check(unsafe { vpbs_ctx_set_shard(self.0, index, count) })
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { vpbs_ctx_destroy(self.0) }
    }
}

/// A page-locked host buffer (vpbs_host_alloc): device copies into it run at full PCIe speed and
/// overlap with kernels; an ordinary `Vec` is pageable, so the driver stages every copy through its
/// own bounce buffers and nothing overlaps (bench.py `e2e_pageable` vs `e2e_eager`).  Derefs to
/// `[u64]`, which is what the patched `MerkleTree { leaves: PinnedBuf, .. }` hands out from
/// `get(i) -> &[F]` (flat leaves).
pub struct PinnedBuf {
    ptr: *mut u64,
    len: usize,
}
unsafe impl Send for PinnedBuf {}
unsafe impl Sync for PinnedBuf {}
impl PinnedBuf {
    pub fn new(len: usize) -> Self {
        let ptr = unsafe { vpbs_host_alloc(len.max(1) * 8) } as *mut u64;
        assert!(!ptr.is_null(), "vpbs_host_alloc({}) failed", len * 8);
        PinnedBuf { ptr, len }
    }
}
impl std::ops::Deref for PinnedBuf {
    type Target = [u64];
    fn deref(&self) -> &[u64] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}
impl std::ops::DerefMut for PinnedBuf {
    fn deref_mut(&mut self) -> &mut [u64] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl Drop for PinnedBuf {
    fn drop(&mut self) {
        unsafe { vpbs_host_free(self.ptr as *mut c_void) }
    }
}

/// Outputs of one commit in the flat layout of include/vpbs_commit.h, in pinned host memory.
pub struct Commit {
    pub coeffs: PinnedBuf,  // ncols x n, column c at [c * n, (c + 1) * n): PolynomialBatch.polynomials
    pub n: usize,
    pub leaves: PinnedBuf,  // m x width, row-major, leaf k = natural LDE row bitrev(k)
    pub width: usize,
    pub digests: PinnedBuf, // 2(m - 2^h) x 4, plonky2 layout
    pub cap: Vec<u64>,      // 2^h x 4
    pub stats: vpbs_stats,
}

/// A PolynomialBatch that stays in HBM (vpbs_batch_*): the cap is on the host, rows / Merkle paths /
/// openings / LDE row blocks are fetched on demand.
pub struct ResidentBatch {
    h: *mut vpbs_batch,
    pub cap: Vec<u64>,
    pub ncols: usize,
    pub degree_log: usize,
    pub rate_bits: usize,
    pub cap_height: usize,
}
unsafe impl Send for ResidentBatch {}
impl Drop for ResidentBatch {
    fn drop(&mut self) {
        unsafe { vpbs_batch_destroy(self.h) }
    }
}
impl ResidentBatch {
    /// PolynomialBatch::from_values / from_coeffs with the result left on the device.
    pub fn commit(ctx: &Ctx, cols: &[&[u64]], rate_bits: usize, cap_height: usize,
                  inputs_are_coeffs: bool) -> Self {
        let n = cols[0].len();
        assert!(n.is_power_of_two());
        let ptrs: Vec<*const u64> = cols.iter().map(|c| c.as_ptr()).collect();
        let mut cap = vec![0u64; 4 << cap_height];
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            vpbs_batch_commit(ctx.0, ptrs.as_ptr(), cols.len() as u32, n.trailing_zeros(),
                              rate_bits as u32, cap_height as u32, inputs_are_coeffs as c_int,
                              std::ptr::null(), cap.as_mut_ptr(), &mut h, std::ptr::null_mut())
        };
        ctx.check(rc);
        ResidentBatch { h, cap, ncols: cols.len(), degree_log: n.trailing_zeros() as usize, rate_bits,
                        cap_height }
    }
    /// MerkleTree::get(i) + MerkleTree::prove(i) for a set of leaf indices (one FRI query round).
    pub fn open(&self, ctx: &Ctx, leaf_indices: &[u64]) -> (Vec<u64>, Vec<u64>) {
        let layers = self.degree_log + self.rate_bits - self.cap_height;
        let mut rows = vec![0u64; leaf_indices.len() * self.ncols];
        let mut sibs = vec![0u64; leaf_indices.len() * layers * 4];
        ctx.check(unsafe { vpbs_batch_get_leaves(self.h, leaf_indices.as_ptr(), leaf_indices.len() as u64,
                                                 rows.as_mut_ptr()) });
        ctx.check(unsafe { vpbs_batch_prove(self.h, leaf_indices.as_ptr(), leaf_indices.len() as u64,
                                            sibs.as_mut_ptr()) });
        (rows, sibs)
    }
    /// get_lde_values(first + i * step, 1) for i < count: the block compute_quotient_polys reads.
    pub fn lde_rows(&self, ctx: &Ctx, first: u64, step: u64, count: usize, out: &mut [u64]) {
        assert!(out.len() >= count * self.ncols);
        ctx.check(unsafe { vpbs_batch_get_lde_rows(self.h, first, step, count as u64, out.as_mut_ptr()) });
    }
    /// prove() steps 4-5: Z and partial products computed from this (wires) batch on the device and
    /// committed as a new resident batch.
    pub fn commit_zs_partial_products(&self, ctx: &Ctx, sigmas: *const vpbs_sigmas, max_degree: usize,
                                      betas: &[u64], gammas: &[u64], num_routed: usize) -> ResidentBatch {
        assert_eq!(betas.len(), gammas.len());
        let mut cap = vec![0u64; 4 << self.cap_height];
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            vpbs_batch_zs_partial_products(self.h, sigmas, max_degree as u32, betas.as_ptr(),
                                           gammas.as_ptr(), betas.len() as u32, self.rate_bits as u32,
                                           self.cap_height as u32, cap.as_mut_ptr(), &mut h,
                                           std::ptr::null_mut())
        };
        ctx.check(rc);
        let chunks = (num_routed + max_degree - 1) / max_degree;
        ResidentBatch { h, cap, ncols: betas.len() * chunks, degree_log: self.degree_log,
                        rate_bits: self.rate_bits, cap_height: self.cap_height }
    }
}

impl ResidentBatch {
    /// OpeningSet::new for this batch: every polynomial at each extension point (npoints x ncols x 2).
    pub fn eval_ext2(&self, ctx: &Ctx, points: &[[u64; 2]]) -> Vec<u64> {
        let mut out = vec![0u64; points.len() * self.ncols * 2];
        ctx.check(unsafe { vpbs_batch_eval_ext2(self.h, points.as_ptr() as *const u64, points.len() as u32,
                                                out.as_mut_ptr()) });
        out
    }
}

/// The circuit's gate constraints as a straight-line program on the device (vpbs_gate_program_upload;
/// instruction format in include/vpbs_commit.h).  `CircuitBuilder::build` compiles it once per circuit
/// by running every gate's `eval_unfiltered_base_one` on a recording field type.
pub struct GateProgram(*mut vpbs_gate_program);
unsafe impl Send for GateProgram {}
impl GateProgram {
    pub fn upload(ctx: &Ctx, code: &[u64], imms: &[u64], nregs: u32, num_constraints: u32) -> Self {
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { vpbs_gate_program_upload(ctx.0, code.as_ptr(), code.len() as u32, imms.as_ptr(),
                                                    imms.len() as u32, nregs, num_constraints, &mut h) });
        GateProgram(h)
    }
}
impl Drop for GateProgram {
    fn drop(&mut self) {
        unsafe { vpbs_gate_program_destroy(self.0) }
    }
}

/// prove() steps 6-7 on the device: compute_quotient_polys (permutation terms from the resident
/// batches, gate constraints from `program` or as alpha-reduced `gate_terms`), then the quotient commit.
#[allow(clippy::too_many_arguments)]
pub fn quotient_polys(ctx: &Ctx, constants_sigmas: &ResidentBatch, sigmas_first_col: usize,
                      wires: &ResidentBatch, zs_pp: &ResidentBatch, k_is: &[u64], max_degree: usize,
                      quotient_degree_bits: usize, betas: &[u64], gammas: &[u64], alphas: &[u64],
                      program: Option<&GateProgram>, public_inputs_hash: &[u64; 4],
                      gate_terms: Option<&[&[u64]]>) -> ResidentBatch {
    assert!(betas.len() == gammas.len() && betas.len() == alphas.len());
    let gt: Vec<*const u64> = gate_terms.map(|g| g.iter().map(|v| v.as_ptr()).collect()).unwrap_or_default();
    let mut cap = vec![0u64; 4 << wires.cap_height];
    let mut h = std::ptr::null_mut();
    ctx.check(unsafe {
        vpbs_batch_quotient_polys(constants_sigmas.h, sigmas_first_col as u32, wires.h, zs_pp.h, k_is.as_ptr(),
                                  k_is.len() as u32, max_degree as u32, quotient_degree_bits as u32,
                                  betas.as_ptr(), gammas.as_ptr(), alphas.as_ptr(), betas.len() as u32,
                                  if gt.is_empty() { std::ptr::null() } else { gt.as_ptr() },
                                  program.map_or(std::ptr::null(), |p| p.0 as *const _),
                                  public_inputs_hash.as_ptr(), wires.rate_bits as u32,
                                  wires.cap_height as u32, cap.as_mut_ptr(), &mut h, std::ptr::null_mut())
    });
    ResidentBatch { h, cap, ncols: betas.len() << quotient_degree_bits, degree_log: wires.degree_log,
                    rate_bits: wires.rate_bits, cap_height: wires.cap_height }
}

/// OpeningSet::new over all FRI oracles in one round trip: per batch (npoints x ncols x 2).
pub fn eval_ext2_all(ctx: &Ctx, batches: &[&ResidentBatch], points: &[[u64; 2]]) -> Vec<Vec<u64>> {
    let hs: Vec<*mut vpbs_batch> = batches.iter().map(|b| b.h).collect();
    let mut outs: Vec<Vec<u64>> = batches.iter().map(|b| vec![0u64; points.len() * b.ncols * 2]).collect();
    let ptrs: Vec<*mut u64> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
    ctx.check(unsafe { vpbs_batches_eval_ext2(hs.as_ptr(), hs.len() as u32, points.as_ptr() as *const u64,
                                              points.len() as u32, ptrs.as_ptr()) });
    outs
}
/// fri_prover_query_round's initial_trees_proof over all oracles in one round trip: per batch
/// (rows, siblings) at the same leaf indices.
pub fn open_all(ctx: &Ctx, batches: &[&ResidentBatch], leaf_indices: &[u64]) -> Vec<(Vec<u64>, Vec<u64>)> {
    let hs: Vec<*mut vpbs_batch> = batches.iter().map(|b| b.h).collect();
    let layers = batches[0].degree_log + batches[0].rate_bits - batches[0].cap_height;
    let mut out: Vec<(Vec<u64>, Vec<u64>)> = batches.iter()
        .map(|b| (vec![0u64; leaf_indices.len() * b.ncols], vec![0u64; leaf_indices.len() * layers * 4]))
        .collect();
    let rows: Vec<*mut u64> = out.iter_mut().map(|o| o.0.as_mut_ptr()).collect();
    let sibs: Vec<*mut u64> = out.iter_mut().map(|o| o.1.as_mut_ptr()).collect();
    ctx.check(unsafe { vpbs_batches_open(hs.as_ptr(), hs.len() as u32, leaf_indices.as_ptr(),
                                         leaf_indices.len() as u64, rows.as_ptr(), sibs.as_ptr()) });
    out
}

/// Sigma polynomials' values + coset shifts of one circuit in HBM (uploaded once per circuit).
pub struct Sigmas(*mut vpbs_sigmas);
unsafe impl Send for Sigmas {}
impl Drop for Sigmas {
    fn drop(&mut self) {
        unsafe { vpbs_sigmas_destroy(self.0) }
    }
}
impl Sigmas {
    pub fn upload(ctx: &Ctx, sigma_cols: &[&[u64]], k_is: &[u64]) -> Self {
        assert_eq!(sigma_cols.len(), k_is.len());
        let n = sigma_cols[0].len();
        assert!(n.is_power_of_two());
        let ptrs: Vec<*const u64> = sigma_cols.iter().map(|c| c.as_ptr()).collect();
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { vpbs_sigmas_upload(ctx.0, ptrs.as_ptr(), k_is.as_ptr(), k_is.len() as u32,
                                              n.trailing_zeros(), &mut h) });
        Sigmas(h)
    }
    pub fn as_ptr(&self) -> *const vpbs_sigmas {
        self.0
    }
}

/// fri/prover.rs fri_committed_trees as one device-resident chain, started the way
/// fri/oracle.rs prove_openings starts it: from the committed batches (FRI_ORACLES order), the
/// FriInstanceInfo (per batch: opening point + (oracle_index, polynomial_index) pairs) and alpha.
pub struct Fri {
    h: *mut vpbs_fri,
    len: usize,
    rate_bits: usize,
    pending_arity_bits: usize,
}
unsafe impl Send for Fri {}
impl Drop for Fri {
    fn drop(&mut self) {
        unsafe { vpbs_fri_destroy(self.h) }
    }
}
impl Fri {
    pub fn begin_openings(ctx: &Ctx, oracles: &[&ResidentBatch], batches: &[(&[(u32, u32)], [u64; 2])],
                          alpha: [u64; 2], rate_bits: usize) -> Self {
        let hs: Vec<*mut vpbs_batch> = oracles.iter().map(|o| o.h).collect();
        let sizes: Vec<u32> = batches.iter().map(|(polys, _)| polys.len() as u32).collect();
        let refs: Vec<u32> = batches.iter().flat_map(|(polys, _)| polys.iter().flat_map(|&(o, p)| [o, p])).collect();
        let points: Vec<u64> = batches.iter().flat_map(|(_, z)| *z).collect();
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe {
            vpbs_fri_begin_openings(ctx.0, hs.as_ptr(), hs.len() as u32, sizes.as_ptr(), sizes.len() as u32,
                                    refs.as_ptr(), points.as_ptr(), alpha.as_ptr(), rate_bits as u32, &mut h)
        });
        Fri { h, len: (1usize << oracles[0].degree_log) << rate_bits, rate_bits, pending_arity_bits: 0 }
    }
    /// One reduction layer's tree; the cap goes to the challenger.
    pub fn commit_layer(&mut self, ctx: &Ctx, arity_bits: usize, cap_height: usize) -> Vec<u64> {
        let mut cap = vec![0u64; 4 << cap_height];
        ctx.check(unsafe { vpbs_fri_commit_layer(self.h, arity_bits as u32, cap_height as u32, cap.as_mut_ptr()) });
        self.pending_arity_bits = arity_bits;
        cap
    }
    /// reduce_with_powers with the challenger's beta, then the next coset evaluations.
    pub fn fold(&mut self, ctx: &Ctx, beta: [u64; 2]) {
        ctx.check(unsafe { vpbs_fri_fold_layer(self.h, beta.as_ptr()) });
        self.len >>= self.pending_arity_bits;
    }
    pub fn final_poly(&self, ctx: &Ctx) -> Vec<u64> {
        let mut out = vec![0u64; 2 * (self.len >> self.rate_bits)];
        ctx.check(unsafe { vpbs_fri_final_poly(self.h, self.rate_bits as u32, out.as_mut_ptr()) });
        out
    }
    /// fri_prover_query_round on one layer's tree: (rows, siblings).
    pub fn query(&self, ctx: &Ctx, layer: usize, leaf_indices: &[u64], leaf_len: usize, layers: usize)
                 -> (Vec<u64>, Vec<u64>) {
        let mut rows = vec![0u64; leaf_indices.len() * leaf_len];
        let mut sibs = vec![0u64; leaf_indices.len() * layers * 4];
        ctx.check(unsafe { vpbs_fri_query_layer(self.h, layer as u32, leaf_indices.as_ptr(),
                                                leaf_indices.len() as u64, rows.as_mut_ptr(),
                                                if sibs.is_empty() { std::ptr::null_mut() } else { sibs.as_mut_ptr() }) });
        (rows, sibs)
    }
}

/// `cols[c]` is one polynomial's values (or coefficients): GoldilocksField is
/// `#[repr(transparent)]` over u64, so `&[GoldilocksField]` reinterprets as `&[u64]`.
pub fn commit(ctx: &Ctx, cols: &[&[u64]], rate_bits: usize, cap_height: usize,
              inputs_are_coeffs: bool, salt: Option<[&[u64]; VPBS_SALT_SIZE]>) -> Commit {
    let ncols = cols.len();
    let n = cols[0].len();
    assert!(n.is_power_of_two());
    let log_n = n.trailing_zeros();
    let m = n << rate_bits;
    let width = ncols + if salt.is_some() { VPBS_SALT_SIZE } else { 0 };
    let col_ptrs: Vec<*const u64> = cols.iter().map(|c| c.as_ptr()).collect();
    let salt_ptrs: Option<Vec<*const u64>> = salt.map(|s| s.iter().map(|c| c.as_ptr()).collect());
    // outputs in pinned memory: the call then costs the PCIe time of its outputs and no more
    let mut coeffs = PinnedBuf::new(ncols * n);
    let coeff_ptrs: Vec<*mut u64> = (0..ncols).map(|c| unsafe { coeffs.as_mut_ptr().add(c * n) }).collect();
    let mut leaves = PinnedBuf::new(m * width);
    let mut digests = PinnedBuf::new(8 * (m - (1 << cap_height)));
    let mut cap = vec![0u64; 4 << cap_height];
    let mut stats = vpbs_stats::default();
    let rc = unsafe {
        vpbs_commit(ctx.0, col_ptrs.as_ptr(), ncols as u32, log_n, rate_bits as u32,
                    cap_height as u32, inputs_are_coeffs as c_int,
                    salt_ptrs.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                    coeff_ptrs.as_ptr(), leaves.as_mut_ptr(),
                    if digests.is_empty() { std::ptr::null_mut() } else { digests.as_mut_ptr() },
                    cap.as_mut_ptr(), &mut stats)
    };
    ctx.check(rc);
    Commit { coeffs, n, leaves, width, digests, cap, stats }
}
