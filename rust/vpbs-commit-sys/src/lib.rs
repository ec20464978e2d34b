//! `extern "C"` declarations for libvpbs_commit.so — one-to-one with include/vpbs_commit.h —
//! plus a small safe wrapper (`Ctx`, `commit`) shaped for plonky2's `PolynomialBatch::from_values`
//! / `from_coeffs` (plonky2 0.2.0 `src/fri/oracle.rs`), which the reference reaches from
//! `prove()` at src/vtfhe/ivc_based_vpbs.rs:302, :333, :364 and `build()` at :275.
//!
//! NOT compiled in this repository's environment (no Rust toolchain there).
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct vpbs_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct vpbs_batch {
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
    pub fn vpbs_batch_shape(batch: *mut vpbs_batch, ncols: *mut u32, log_n: *mut u32,
                            rate_bits: *mut u32, cap_height: *mut u32, width: *mut u32) -> c_int;
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
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { vpbs_ctx_destroy(self.0) }
    }
}

/// Outputs of one commit in the flat layout of include/vpbs_commit.h.
pub struct Commit {
    pub coeffs: Vec<Vec<u64>>, // PolynomialBatch.polynomials
    pub leaves: Vec<u64>,      // m x width, row-major, leaf k = natural LDE row bitrev(k)
    pub width: usize,
    pub digests: Vec<u64>,     // 2(m - 2^h) x 4, plonky2 layout
    pub cap: Vec<u64>,         // 2^h x 4
    pub stats: vpbs_stats,
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
    let mut coeffs = vec![vec![0u64; n]; ncols];
    let coeff_ptrs: Vec<*mut u64> = coeffs.iter_mut().map(|c| c.as_mut_ptr()).collect();
    let mut leaves = vec![0u64; m * width];
    let mut digests = vec![0u64; 8 * (m - (1 << cap_height))];
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
    Commit { coeffs, leaves, width, digests, cap, stats }
}
