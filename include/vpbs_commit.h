/*
 * vpbs_commit.h — C ABI of libvpbs_commit.so: the B200 (sm_100a) polynomial-commitment path of
 * vPBS (zama-ai/verifiable-fhe-paper), i.e. what plonky2 0.2.0 does inside
 * PolynomialBatch::from_values / from_coeffs and MerkleTree::new for
 * F = GoldilocksField, C = PoseidonGoldilocksConfig.
 *
 * The reference reaches this path only through plonky2 (crates.io dependency,
 * /root/reference/Cargo.toml:7): `prove()` at /root/reference/src/vtfhe/ivc_based_vpbs.rs:302,
 * :333, :364 and `builder.build::<C>()` at :40, :46, :61, :275.  Each entry point below names the
 * plonky2 0.2.0 item ("[P2] file::item") whose FFI replacement it is; INTEGRATION.md shows the
 * Rust `extern "C"` block and the plonky2 patch that bind them.
 *
 * Conventions
 *  - every field element is a uint64_t (GoldilocksField is #[repr(transparent)] over u64);
 *    inputs may be non-canonical (any u64), outputs are canonical (< p = 2^64 - 2^32 + 1);
 *  - a hash (HashOut) is 4 consecutive uint64_t;
 *  - all functions return VPBS_OK (0) or a negative vpbs_status; nothing throws or aborts across
 *    the boundary; vpbs_last_error() gives the message of the last failure on that context;
 *  - there is NO CPU fallback: without a usable CUDA device every call fails with VPBS_ERR_CUDA;
 *  - a context is bound to one device and one stream and is not re-entrant (one call in flight
 *    per context); distinct contexts are independent and may be used from different threads;
 *  - host entry points take caller-owned host memory and never retain it past return;
 *    `*_dev` entry points take device pointers on the context's device and are asynchronous on
 *    the context's stream (see vpbs_ctx_set_stream / vpbs_ctx_sync).
 */
#ifndef VPBS_COMMIT_H
#define VPBS_COMMIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: additions only (second half of round 2): vpbs_ctx_set_host_threads, vpbs_ctx_set_shard,
 * vpbs_batch_shard, vpbs_batches_eval_ext2, vpbs_batches_open, vpbs_batch_quotient_polys / _values,
 * vpbs_quotient_commit_values, vpbs_gate_program_upload / _destroy; every version-2 entry is unchanged. */
#define VPBS_ABI_VERSION 3
#define VPBS_SALT_SIZE 4 /* [P2] fri/oracle.rs SALT_SIZE */

typedef enum vpbs_status {
  VPBS_OK = 0,
  VPBS_ERR_ARG = -1,   /* what plonky2 assert!s on: non power of two, cap_height > log2(leaves), ... */
  VPBS_ERR_CUDA = -2,  /* no device / kernel or copy failure */
  VPBS_ERR_OOM = -3,   /* device or pinned-host allocation failed */
  VPBS_ERR_STATE = -4  /* context unusable (destroyed / wrong device) */
} vpbs_status;

typedef struct vpbs_ctx vpbs_ctx;

/* Per-call timings in milliseconds, named after the TimingTree scopes plonky2 prints for the same
 * work ([P2] fri/oracle.rs timed!(..) labels; the reference prints them per step,
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:301-309).  "transpose LDEs" is fused into the last
 * NTT pass here, so it is reported inside fft_ms. */
typedef struct vpbs_stats {
  float h2d_ms;    /* host -> device copies of the inputs (0 for *_dev calls); for wide
                      batches they are pipelined with the transforms (see fft_ms)  */
  float ifft_ms;   /* "IFFT" (0 when the host API pipelines by column chunk: then
                      the IFFTs are interleaved with the LDEs and counted there)   */
  float fft_ms;    /* "FFT + blinding" + "transpose LDEs"                          */
  float merkle_ms; /* "build Merkle tree" (leaf hashing + all levels)              */
  float leaf_hash_ms; /* the leaf-hashing kernel alone (part of merkle_ms)        */
  float d2h_ms;    /* output copies NOT hidden behind kernels: last kernel -> last copy
                      (the host API streams coefficients / leaf blocks out on a second
                      stream while later kernels run; 0 for *_dev calls)           */
  float total_ms;  /* first event to last event of the call                        */
  uint64_t kernel_launches; /* kernels of this library launched by the call        */
} vpbs_stats;

/* ---- context ------------------------------------------------------------------------------ */
int vpbs_abi_version(void);
int vpbs_device_count(void); /* >= 0, or a negative vpbs_status */
int vpbs_ctx_create(int device, vpbs_ctx** out);
void vpbs_ctx_destroy(vpbs_ctx* ctx);
/* Use an existing CUDA stream (a cudaStream_t passed as void*; NULL = the context's own). */
int vpbs_ctx_set_stream(vpbs_ctx* ctx, void* cuda_stream);
int vpbs_ctx_sync(vpbs_ctx* ctx);
const char* vpbs_last_error(vpbs_ctx* ctx); /* ctx may be NULL: last create/global error */
/* Total kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t vpbs_ctx_kernel_launches(vpbs_ctx* ctx);

/* Host columns handed to vpbs_commit / vpbs_batch_commit in ORDINARY (pageable) memory — what
 * plonky2's prover passes to PolynomialBatch::from_values: Vec<PolynomialValues<F>> it allocated
 * itself, [P2] plonk/prover.rs — are staged by the library through a pinned ring of the context:
 * `threads` copy threads fill 4 MiB slots in parallel and an uploader thread sends them on the
 * H2D stream while the kernels of the previous column chunk already run.
 * Default 4; 0 leaves such copies to the CUDA driver's own staging (serial, on the calling
 * thread, before any kernel is enqueued).  Page-locked inputs (vpbs_host_alloc) never take
 * this path. */
int vpbs_ctx_set_host_threads(vpbs_ctx* ctx, unsigned threads);

/* Row-range sharding of ONE proof over several GPUs (partitioning B): after
 * vpbs_ctx_set_shard(ctx, index, count) every resident batch this context creates
 * (vpbs_batch_commit / _commit_dev / vpbs_batch_zs_partial_products) keeps ALL coefficient columns
 * but transforms, hashes and holds only leaves [index * m / count, (index + 1) * m / count) and the
 * digests / cap entries of the cap subtrees under them.  count: a power of two; m / count must be
 * whole n-row LDE blocks and whole cap subtrees (count <= 2^rate_bits and <= 2^cap_height), else
 * the commit returns VPBS_ERR_ARG.  cap_out of such a commit has all 2^cap_height entries, the ones
 * of other shards ZERO: the caller fills them by gathering the other shards' entries (NCCL
 * all-gather of 32 B per entry — the only cross-GPU traffic of a sharded commit); the union is
 * bit-identical to the unsharded commit.  vpbs_batch_get_leaves / _prove / _get_lde_rows of a
 * sharded batch serve the rows of its own range (global leaf indices; others: VPBS_ERR_ARG),
 * vpbs_batch_download copies the shard's rows and digests.  Everything that reads coefficients
 * (vpbs_batch_eval_ext2, vpbs_batch_zs_partial_products, vpbs_fri_begin_openings) is unaffected.
 * count == 1 (the default) switches sharding off. */
int vpbs_ctx_set_shard(vpbs_ctx* ctx, uint32_t index, uint32_t count);

/* Pinned host memory for callers that want full-speed PCIe copies. */
void* vpbs_host_alloc(size_t bytes);
void vpbs_host_free(void* p);

/* ---- single-polynomial transforms (host memory, in place, natural order in and out) -------- */
/* [P2] plonky2_field/src/fft.rs fft_with_options(poly, None, None): out[i] = sum_j c_j w^(ij). */
int vpbs_fft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n);
/* [P2] plonky2_field/src/fft.rs ifft_with_options: c_j = n^-1 sum_i v_i w^(-ij). */
int vpbs_ifft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n);
/* [P2] plonky2_field/src/polynomial/mod.rs PolynomialCoeffs::coset_fft_with_options(shift, ..). */
int vpbs_coset_fft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n, uint64_t shift);

/* ---- hashing (host memory) ------------------------------------------------------------------ */
/* [P2] plonky2/src/hash/poseidon.rs Poseidon::poseidon on `count` independent 12-word states. */
int vpbs_poseidon_permute(vpbs_ctx* ctx, uint64_t* states_inout, uint64_t count);
/* [P2] plonk/config.rs Hasher::hash_or_noop over `count` rows of `len` elements (row-major):
 * len <= 4 copies (canonical, zero padded), otherwise hash_n_to_m_no_pad (overwrite-mode sponge). */
int vpbs_hash_or_noop_batch(vpbs_ctx* ctx, const uint64_t* rows, uint64_t count, uint32_t len,
                            uint64_t* hashes_out /* count x 4 */);
/* [P2] hash/hashing.rs compress == PoseidonHash::two_to_one on `count` pairs. */
int vpbs_two_to_one_batch(vpbs_ctx* ctx, const uint64_t* left /* count x 4 */,
                          const uint64_t* right /* count x 4 */, uint64_t count,
                          uint64_t* hashes_out /* count x 4 */);

/* ---- [P2] plonky2/src/hash/merkle_tree.rs MerkleTree::new(leaves, cap_height) ----------------
 * leaves_rowmajor: nleaves x leaf_len.  digests_out: 2*(nleaves - 2^cap_height) hashes in
 * plonky2's layout (per cap subtree: left subtree digests || left child || right child || right
 * subtree digests, so MerkleTree::prove's index formula applies unchanged); may be NULL when that
 * count is 0 (tree is all cap).  cap_out: 2^cap_height hashes.
 * VPBS_ERR_ARG if nleaves is not a power of two or cap_height > log2(nleaves) (plonky2 panics). */
int vpbs_merkle_new(vpbs_ctx* ctx, const uint64_t* leaves_rowmajor, uint64_t nleaves,
                    uint32_t leaf_len, uint32_t cap_height, uint64_t* digests_out,
                    uint64_t* cap_out);

/* MerkleTree::new on device-resident leaves (nleaves x leaf_len row-major on the context's device);
 * digests / cap stay on the device.  Asynchronous on the context's stream; stats forces a sync
 * and reports leaf_hash_ms / merkle_ms. */
int vpbs_merkle_new_dev(vpbs_ctx* ctx, const uint64_t* d_leaves, uint64_t nleaves, uint32_t leaf_len,
                        uint32_t cap_height, uint64_t* d_digests_out, uint64_t* d_cap_out,
                        vpbs_stats* stats);

/* ---- [P2] plonky2/src/fri/oracle.rs PolynomialBatch::lde_values --------------------------------
 * cols: ncols pointers to n = 2^log_n elements each (a Vec<PolynomialValues<F>> is ncols separately
 * allocated Vec<F>).  inputs_are_coeffs = 0: values (ifft first, as from_values), 1: coefficients.
 * coeffs_out: NULL or ncols pointers receiving the n coefficients of each polynomial.
 * lde_cols_out: ncols * (n << rate_bits), column-major, NATURAL order: column c, index i holds
 * poly_c(7 * w_m^i) — exactly lde_values()[c][i]. */
int vpbs_lde_batch(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                   uint32_t rate_bits, int inputs_are_coeffs, uint64_t* const* coeffs_out,
                   uint64_t* lde_cols_out);

/* ---- [P2] plonky2/src/fri/oracle.rs PolynomialBatch::from_values / from_coeffs ------------------
 * The whole commit: (ifft) -> lde + coset fft (shift 7) -> transpose + reverse_index_bits ->
 * MerkleTree::new.  m = n << rate_bits leaves of width ncols (+4 if salt_cols).
 *  salt_cols   NULL (blinding = false, the only case the reference uses:
 *              CircuitConfig::standard_recursion_config(), ivc_based_vpbs.rs:38,41,48,190), or
 *              VPBS_SALT_SIZE pointers to m random elements each, drawn by the caller (the RNG stays
 *              on the host so the device path is deterministic); they are appended to every leaf
 *              the way lde_values() chains them.
 *  coeffs_out  NULL or ncols pointers (n each): PolynomialBatch.polynomials.
 *  leaves_out  m x width row-major, leaf k = natural LDE row bitrev(k): MerkleTree.leaves.
 *              May be NULL (cap/digests only).
 *  digests_out 2*(m - 2^cap_height) hashes, plonky2 layout; NULL allowed only if that is 0 or the
 *              caller does not want them.
 *  cap_out     2^cap_height hashes (required).
 *  stats       NULL or timing breakdown. */
int vpbs_commit(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                const uint64_t* const* salt_cols, uint64_t* const* coeffs_out, uint64_t* leaves_out,
                uint64_t* digests_out, uint64_t* cap_out, vpbs_stats* stats);

/* The same commit spread over nctx GPUs of this process (SURVEY.md §8(b) vpbs_commit_multi, §8(e)
 * partitioning B).  Every GPU needs all input columns: each column chunk crosses PCIe once, into
 * the GPU that owns it (chunks round-robin over the GPUs, all host links in parallel), and reaches
 * the others by peer copies (NVLink where peer access exists); GPU g transforms and hashes only leaf
 * rows [g*m/nctx, (g+1)*m/nctx) — whole LDE blocks, already in leaf order — and copies those rows,
 * the digests and the cap entries of the cap subtrees above them straight into the caller's buffers
 * over its own host link; coefficient column c comes from GPU c*nctx/ncols.  The outputs are
 * bit-identical to vpbs_commit's.
 *  ctxs   nctx contexts on distinct devices; nctx a power of two, <= 2^rate_bits and <= 2^cap_height.
 * The call starts every device before it waits for any.  stats: the slowest device's breakdown,
 * kernel_launches summed.  On error vpbs_last_error(ctxs[0]) explains. */
int vpbs_commit_multi(vpbs_ctx* const* ctxs, int nctx, const uint64_t* const* cols, uint32_t ncols,
                      uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                      const uint64_t* const* salt_cols, uint64_t* const* coeffs_out,
                      uint64_t* leaves_out, uint64_t* digests_out, uint64_t* cap_out,
                      vpbs_stats* stats);

/* Same computation on device-resident data (the prover keeps its batches in HBM; also what
 * bench.py times as the kernel-only number).  d_cols: ncols x n column-major, contiguous.
 * d_salt: NULL or 4 x m.  d_coeffs_out: NULL or ncols x n.  d_leaves_out: m x width (required).
 * d_digests_out: required unless the tree is all cap.  d_cap_out: required.
 * Asynchronous on the context's stream; stats (if non-NULL) forces a sync. */
int vpbs_commit_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                    uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                    const uint64_t* d_salt, uint64_t* d_coeffs_out, uint64_t* d_leaves_out,
                    uint64_t* d_digests_out, uint64_t* d_cap_out, vpbs_stats* stats);

/* Row-range shard of one commit for multi-GPU proving (SURVEY.md §8(e) partitioning B): computes
 * only leaves [first_leaf, first_leaf + nleaves_shard) of the m-leaf tree — already in leaf order —
 * and the digests and subtree roots under it.  A shard is a whole number of n-row LDE blocks
 * (first_leaf and nleaves_shard multiples of n = 2^log_n); unless it is the whole tree it must also
 * be a power of two, aligned to its own size, and cover whole cap subtrees
 * (nleaves_shard >= m >> cap_height).  Anything else returns VPBS_ERR_ARG.
 *  d_cols            all ncols x n inputs (every shard needs every coefficient column)
 *  d_coeffs_out      NULL or ncols x n; with inputs_are_coeffs = 1 it receives a plain copy of the
 *                    inputs (not canonicalised)
 *  d_leaves_out      nleaves_shard x width
 *  d_digests_out     the plonky2-layout digests of the cap subtrees this shard owns
 *                    (2*(nleaves_shard - nroots) hashes)
 *  d_roots_out       nroots = nleaves_shard >> (log2 m - cap_height) hashes: the cap entries
 *                    first_leaf >> (log2 m - cap_height) ... owned by this shard.
 * The only cross-GPU traffic of a sharded commit is the gather of d_roots_out (32 B per entry). */
int vpbs_commit_shard_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                          uint64_t first_leaf, uint64_t nleaves_shard, uint64_t* d_coeffs_out,
                          uint64_t* d_leaves_out, uint64_t* d_digests_out, uint64_t* d_roots_out,
                          vpbs_stats* stats);

/* ---- openings: evaluate polynomials at extension-field points -----------
 * [P2] plonky2_field/src/polynomial/mod.rs PolynomialCoeffs::to_extension().eval(zeta) over
 * QuadraticExtension<GoldilocksField> (F[X]/(X^2 - 7), an element is two consecutive uint64_t), as
 * plonk/proof.rs OpeningSet::new does for every committed polynomial at zeta and g*zeta
 * ("construct the opening set").  out[(p * ncols + c) * 2 + {0,1}] = poly_c(points[p]). */
int vpbs_eval_ext2(vpbs_ctx* ctx, const uint64_t* const* coeff_cols, uint32_t ncols, uint32_t log_n,
                   const uint64_t* points, uint32_t npoints, uint64_t* out);

/* ---- FRI commit phase, one reduction layer -----------------------------
 * [P2] plonky2/src/fri/prover.rs fri_committed_trees over D = 2 extension elements stored as
 * (re, im) pairs; the challenger stays with the caller (it supplies beta between the two calls).
 *  vpbs_fri_layer_commit: reverse_index_bits_in_place(values); leaves = chunks of 2^arity_bits
 *      values, flattened (2 * arity base elements per leaf); MerkleTree::new(leaves, cap_height).
 *      leaves_out: (len >> arity_bits) x (2 << arity_bits), may be NULL; digests_out / cap_out as
 *      vpbs_merkle_new.
 *  vpbs_fri_fold: coeffs' = chunks_exact(arity).map(|c| reduce_with_powers(c, beta)) (len >>
 *      arity_bits elements) and values' = coeffs'.coset_fft(shift_next) in natural order, where
 *      shift_next = shift^arity is maintained by the caller as upstream does. */
int vpbs_fri_layer_commit(vpbs_ctx* ctx, const uint64_t* values_ext, uint64_t len,
                          uint32_t arity_bits, uint32_t cap_height, uint64_t* leaves_out,
                          uint64_t* digests_out, uint64_t* cap_out);
int vpbs_fri_fold(vpbs_ctx* ctx, const uint64_t* coeffs_ext, uint64_t len, uint32_t arity_bits,
                  const uint64_t beta[2], uint64_t shift_next, uint64_t* coeffs_out,
                  uint64_t* values_out);

/* The same commit phase as ONE device-resident chain: the polynomial (coefficients and coset
 * evaluations) and every layer's tree stay in HBM; per layer the host supplies only beta and gets
 * back only the cap, as [P2] fri/prover.rs fri_committed_trees needs for the challenger.
 *   vpbs_fri_begin        final_poly_coeffs_ext: ncoeffs extension coefficients (host).  Does what
 *                         [P2] fri/oracle.rs prove_openings does before fri_proof: lde(rate_bits)
 *                         (zero padding) and coset_fft(F::coset_shift()) ("perform final FFT").
 *   vpbs_fri_commit_layer reverse_index_bits + chunk(2^arity_bits) + MerkleTree::new(cap_height) on the
 *                         current values; cap_out: 2^cap_height hashes.  The tree is kept as layer
 *                         number (calls so far).
 *   vpbs_fri_fold_layer   coeffs <- reduce_with_powers(chunks, beta), shift <- shift^arity,
 *                         values <- coeffs.coset_fft(shift); arity = that of the layer just committed.
 *   vpbs_fri_final_poly   coeffs.truncate(len >> rate_bits) -> coeffs_out ((len >> rate_bits) x 2).
 *   vpbs_fri_query_layer  [P2] fri_prover_query_round: MerkleTree::get / prove of a layer's tree;
 *                         rows_out: count x (2 << arity_bits), siblings_out: count x (log2 leaves -
 *                         cap_height) hashes (may be NULL).
 * Calls must alternate commit_layer / fold_layer (VPBS_ERR_STATE otherwise). */
typedef struct vpbs_fri vpbs_fri;
int vpbs_fri_begin(vpbs_ctx* ctx, const uint64_t* final_poly_coeffs_ext, uint64_t ncoeffs,
                   uint32_t rate_bits, vpbs_fri** out);
/* The chain started from the committed batches themselves: [P2] fri/oracle.rs prove_openings from its
 * first line to lde_final_values, on the device.  oracles: the resident batches (FRI_ORACLES order is
 * the caller's); the FRI instance ([P2] FriInstanceInfo) is nbatches batches of batch_sizes[b]
 * polynomials each, poly_refs holding their (oracle_index, polynomial_index) pairs back to back
 * ([P2] FriPolynomialInfo), opened at the extension points points[2b], points[2b+1].  Per batch:
 *   F_b = sum_j alpha^j f_bj (ReducingFactor::reduce_polys_base),
 *   Q_b = (F_b(X) - F_b(z_b)) / (X - z_b) padded with a zero (divide_by_linear),
 *   final_poly = final_poly * alpha^(batch_sizes[b]) + Q_b (shift_poly, +=);
 * then final_poly.lde(rate_bits).coset_fft(F::coset_shift()) as vpbs_fri_begin.  Only alpha and
 * the points cross PCIe.  vpbs_fri_final_poly(fri, rate_bits, ..) right after this call returns
 * final_poly itself (n coefficients). */
struct vpbs_batch; /* resident batches: declared below */
int vpbs_fri_begin_openings(vpbs_ctx* ctx, struct vpbs_batch* const* oracles, uint32_t noracles,
                            const uint32_t* batch_sizes, uint32_t nbatches, const uint32_t* poly_refs,
                            const uint64_t* points, const uint64_t alpha[2], uint32_t rate_bits,
                            vpbs_fri** out);
int vpbs_fri_commit_layer(vpbs_fri* fri, uint32_t arity_bits, uint32_t cap_height, uint64_t* cap_out);
int vpbs_fri_fold_layer(vpbs_fri* fri, const uint64_t beta[2]);
int vpbs_fri_final_poly(vpbs_fri* fri, uint32_t rate_bits, uint64_t* coeffs_out);
int vpbs_fri_query_layer(vpbs_fri* fri, uint32_t layer, const uint64_t* leaf_indices, uint64_t count,
                         uint64_t* rows_out, uint64_t* siblings_out);
void vpbs_fri_destroy(vpbs_fri* fri);

/* ---- FRI proof-of-work grind -------------------------------------------
 * [P2] plonky2/src/fri/prover.rs fri_proof_of_work: the challenger's duplex state with the
 * candidate witness written at `witness_pos`, one permutation, and the response word
 * (`response_lane`, upstream: the last squeezed element = lane 7) must have at least
 * `min_leading_zeros` leading zero bits as a canonical u64.  Scans candidates
 * first_candidate .. first_candidate + count - 1 and returns the SMALLEST that qualifies in
 * *witness_out (deterministic, unlike upstream's rayon find_any) with *found = 1, or *found = 0. */
int vpbs_pow_grind(vpbs_ctx* ctx, const uint64_t state[12], uint32_t witness_pos,
                   uint32_t response_lane, uint32_t min_leading_zeros, uint64_t first_candidate,
                   uint64_t count, uint64_t* witness_out, int* found);

/* ---- device-resident batches (keep the LDE in HBM, open lazily) ---------
 * vpbs_batch_commit is vpbs_commit that returns only the cap and KEEPS coefficients, leaves and
 * digests in device memory owned by the handle.  The readers below serve what plonky2 reads later
 * from a PolynomialBatch: [P2] hash/merkle_tree.rs MerkleTree::get / prove and fri/oracle.rs
 * get_lde_values (28 FRI queries x 4 batches per proof), so the m x width leaf matrix need not
 * cross PCIe for batches that are only opened (the quotient and FRI commits). */
typedef struct vpbs_batch vpbs_batch;
int vpbs_batch_commit(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                      uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                      const uint64_t* const* salt_cols, uint64_t* cap_out, vpbs_batch** out,
                      vpbs_stats* stats);
/* A batch may outlive its context: vpbs_ctx_destroy releases the device buffers of every live batch
 * (and sigma set) of that context and leaves the handles inert — later reads return VPBS_ERR_STATE,
 * and vpbs_batch_destroy only frees the handle. */
void vpbs_batch_destroy(vpbs_batch* batch);
/* rows_out: count x width, row i = leaf leaf_indices[i] (salt columns included). */
int vpbs_batch_get_leaves(vpbs_batch* batch, const uint64_t* leaf_indices, uint64_t count,
                          uint64_t* rows_out);
/* siblings_out: count x (log2 m - cap_height) hashes, MerkleProof.siblings of each leaf. */
int vpbs_batch_prove(vpbs_batch* batch, const uint64_t* leaf_indices, uint64_t count,
                     uint64_t* siblings_out);
/* Bulk download of what vpbs_commit returns eagerly; any pointer may be NULL. */
int vpbs_batch_download(vpbs_batch* batch, uint64_t* const* coeffs_out, uint64_t* leaves_out,
                        uint64_t* digests_out);
/* vpbs_eval_ext2 on the coefficients the batch already holds in HBM (PolynomialBatch.polynomials). */
int vpbs_batch_eval_ext2(vpbs_batch* batch, const uint64_t* points, uint32_t npoints, uint64_t* out);
/* vpbs_batch_commit on columns that are already in HBM (d_cols: ncols x n column-major on the
 * context's device, e.g. produced by another kernel of the prover); no salt. */
int vpbs_batch_commit_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                          uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats);
/* [P2] fri/oracle.rs PolynomialBatch::get_lde_values(index, 1) for the LDE indices
 * first_index + i * step, i < count — what plonk/prover.rs compute_quotient_polys reads row block
 * by row block (get_lde_values_packed).  rows_out: count x ncols (salt columns dropped), row i =
 * leaf reverse_bits(first_index + i * step).  Lets a CPU quotient pull the LDE lazily, in blocks,
 * instead of receiving the whole m x width matrix at commit time. */
int vpbs_batch_get_lde_rows(vpbs_batch* batch, uint64_t first_index, uint64_t step, uint64_t count,
                            uint64_t* rows_out);
/* [P2] plonk/proof.rs OpeningSet::new in one round trip: outs[k] = (npoints x ncols_k x 2) openings
 * of batch k, all batches at the same points (one upload, one kernel per batch, one wait). */
int vpbs_batches_eval_ext2(vpbs_batch* const* batches, uint32_t nbatches, const uint64_t* points,
                           uint32_t npoints, uint64_t* const* outs);
/* [P2] fri/prover.rs fri_prover_query_round (initial_trees_proof) in one round trip: for every batch
 * the leaf rows (count x width_k) and Merkle paths (count x (log2 m - cap_height) x 4) at the same
 * leaf indices.  The batches must share the tree shape (and the shard, if their context shards). */
int vpbs_batches_open(vpbs_batch* const* batches, uint32_t nbatches, const uint64_t* leaf_indices,
                      uint64_t count, uint64_t* const* rows_out, uint64_t* const* siblings_out);
/* Shape of the batch: m = 2^(log_n + rate_bits) leaves of `width` elements. */
int vpbs_batch_shape(vpbs_batch* batch, uint32_t* ncols, uint32_t* log_n, uint32_t* rate_bits,
                     uint32_t* cap_height, uint32_t* width);
/* Leaves the batch holds: [first_leaf, first_leaf + nleaves) (0 and m unless its context shards). */
int vpbs_batch_shard(vpbs_batch* batch, uint64_t* first_leaf, uint64_t* nleaves);

/* ---- permutation argument: Z and partial products (first device consumer) ---
 * [P2] plonky2/src/plonk/prover.rs wires_permutation_partial_products_and_zs /
 * all_wires_permutation_partial_products with util/partial_products.rs quotient_chunk_products and
 * partial_products_and_z_gx: step 4 of prove() ("compute partial products"), reached from
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364.  For challenge c and row i
 * (x_i = w_n^i):  q_j = (wire_j + beta_c k_j x_i + gamma_c) / (wire_j + beta_c sigma_j(x_i) + gamma_c),
 * chunk products over max_degree (= quotient_degree_factor) routed wires, K = ceil(num_routed /
 * max_degree) chunks; Z_c(x_0) = 1, partial product k = Z_c(x_i) * chunk_0 .. chunk_k (k < K - 1),
 * Z_c(x_{i+1}) = Z_c(x_i) * chunk_0 .. chunk_{K-1}.
 * Output columns are in the order prove() commits them ("Z is expected at the front of our batch"):
 * Z_0 .. Z_{C-1}, then the K - 1 partial products of challenge 0, of challenge 1, ... (C * K columns:
 * 2 + 18 = 20 for CircuitConfig::standard_recursion_config()).
 * VPBS_ERR_ARG if a denominator is zero (upstream's batch_multiplicative_inverse panics). */
typedef struct vpbs_sigmas vpbs_sigmas;
/* The circuit's sigma polynomial VALUES and coset shifts, uploaded once per circuit
 * ([P2] ProverOnlyCircuitData.sigmas transposed: sigma_cols[j][i] = sigmas[i][j]; k_is =
 * CommonCircuitData.k_is).  num_routed pointers to n = 2^log_n values each. */
int vpbs_sigmas_upload(vpbs_ctx* ctx, const uint64_t* const* sigma_cols, const uint64_t* k_is,
                       uint32_t num_routed, uint32_t log_n, vpbs_sigmas** out);
void vpbs_sigmas_destroy(vpbs_sigmas* sigmas);
/* Host in, host out: wire_cols[j][i] = witness.get_wire(i, j) for the num_routed routed wires;
 * cols_out: num_challenges * K pointers (NULL entries are skipped) receiving n values each. */
int vpbs_zs_partial_products(vpbs_ctx* ctx, const uint64_t* const* wire_cols,
                             const vpbs_sigmas* sigmas, uint32_t max_degree, const uint64_t* betas,
                             const uint64_t* gammas, uint32_t num_challenges,
                             uint64_t* const* cols_out);
/* Device-resident form: the routed wire values are recovered from the first num_routed coefficient
 * columns of the resident wires batch (one forward transform in HBM), and the C * K result columns
 * are committed at once as a new resident batch (PolynomialBatch::from_values(zs_partial_products,
 * rate_bits, false, cap_height)): only the cap crosses PCIe. */
int vpbs_batch_zs_partial_products(vpbs_batch* wires, const vpbs_sigmas* sigmas, uint32_t max_degree,
                                   const uint64_t* betas, const uint64_t* gammas,
                                   uint32_t num_challenges, uint32_t rate_bits, uint32_t cap_height,
                                   uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats);

/* ---- quotient polynomials: the gate-independent part + the tail --------------
 * [P2] plonky2/src/plonk/prover.rs compute_quotient_polys with plonk/vanishing_poly.rs
 * eval_vanishing_poly_base_batch restricted to the terms that do not depend on the gate set (step 6
 * of prove(), /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364).  On the quotient domain
 * (n << quotient_degree_bits points of the coset 7<w>, read from the resident batches' LDE rows):
 *   L_0(x) (Z_c(x) - 1)   and   the partial-product checks of every challenge
 *   (check_partial_products with max_degree = quotient_degree_factor), in upstream's order, reduced
 *   with the powers of every alpha (reduce_with_powers_multi), divided by Z_H(x) = x^n - 1
 *   (ZeroPolyOnCoset), then per challenge coset_ifft(7), split into 2^quotient_degree_bits chunks of
 *   n coefficients and committed from coefficients as a new resident batch (prove() step 7:
 *   num_challenges * 2^quotient_degree_bits columns, challenge-major).
 * The gate constraints (every gate's eval_unfiltered_base_batch with its selector filter; the gate
 * definitions are plonky2 source that is not restated here) enter as ALREADY alpha-reduced values:
 *   gate_terms[c][i] = sum_j alpha_c^j gate_constraint_j(7 w_q^i),  i < n << quotient_degree_bits,
 * natural order, host memory, NULL = no gate constraints; they are added times
 * alpha_c^(number of permutation terms), which is where upstream's term order puts them.
 *   constants_sigmas: the circuit's batch; its sigma polynomials are columns
 *                     [sigmas_first_col, sigmas_first_col + num_routed)
 *   wires:            routed wires are its first num_routed columns
 *   zs_pp:            the batch vpbs_batch_zs_partial_products returned (Z_0 .. Z_{nc-1}, then the
 *                     partial products of challenge 0, 1, ..)
 *   k_is:             num_routed coset shifts (host)
 * All three batches must be unsharded and have rate_bits >= quotient_degree_bits; at most 4
 * challenges.
 *
 * Instead of gate_terms the gate constraints can be evaluated ON THE DEVICE from a program
 * ([P2] plonk/vanishing_poly.rs evaluate_gate_constraints_base_batch as data): the host compiles,
 * once per circuit, every gate's eval_unfiltered_base and its selector filter into straight-line
 * code over the local wires, the local constants (columns of the constants/sigmas batch) and
 * public_inputs_hash; the device runs it at every point of the quotient domain, so no LDE row
 * crosses PCIe.  One 64-bit word per instruction:
 *     bits 0-7 op | 8-15 dst register | 16-19 kind(a) | 20-23 kind(b) | 24-39 index(a) | 40-55 index(b)
 *     op   0 ADD, 1 SUB, 2 MUL: dst <- a op b;  5 MAD: dst <- dst + a b;  3 EMIT: constraint number
 *          index(b) of the current gate has value a;  4 ENDGATE: the current gate's constraints,
 *          times the filter value a, are added to the totals (constraint j counts alpha_c^j, as
 *          reduce_with_powers does)
 *     kind 0 register, 1 wire column, 2 column of the constants/sigmas batch, 3 immediate table
 *          entry, 4 public_inputs_hash element
 * nregs <= 224 registers (shared memory), num_constraints = the circuit's num_gate_constraints. */
typedef struct vpbs_gate_program vpbs_gate_program;
int vpbs_gate_program_upload(vpbs_ctx* ctx, const uint64_t* code, uint32_t ncode, const uint64_t* imms,
                             uint32_t nimm, uint32_t nregs, uint32_t num_constraints,
                             vpbs_gate_program** out);
void vpbs_gate_program_destroy(vpbs_gate_program* program);
/* gate_terms and program are mutually exclusive (either or both NULL); public_inputs_hash: 4
 * elements (NULL: zeros), read by programs only. */
int vpbs_batch_quotient_polys(vpbs_batch* constants_sigmas, uint32_t sigmas_first_col, vpbs_batch* wires,
                              vpbs_batch* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                              uint32_t max_degree, uint32_t quotient_degree_bits, const uint64_t* betas,
                              const uint64_t* gammas, const uint64_t* alphas, uint32_t num_challenges,
                              const uint64_t* const* gate_terms, const vpbs_gate_program* program,
                              const uint64_t* public_inputs_hash, uint32_t rate_bits, uint32_t cap_height,
                              uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats);

/* The two halves of vpbs_batch_quotient_polys, for a proof whose resident batches are sharded by row
 * range (vpbs_ctx_set_shard; needs rate_bits == quotient_degree_bits, i.e. the quotient domain is the
 * whole LDE): vpbs_batch_quotient_values writes the quotient VALUES of the rank's own leaves into
 * d_vals_out (DEVICE memory, num_challenges x (n << quotient_degree_bits), natural order; the entries of
 * other shards are zeroed), the ranks add their buffers up (an all-reduce over NVLink: a value and
 * zeros), and vpbs_quotient_commit_values turns the complete values into the rank's shard of the
 * quotient batch (coset_ifft, chunks, from_coeffs).  On unsharded batches the pair equals
 * vpbs_batch_quotient_polys. */
int vpbs_batch_quotient_values(vpbs_batch* constants_sigmas, uint32_t sigmas_first_col, vpbs_batch* wires,
                               vpbs_batch* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                               uint32_t max_degree, uint32_t quotient_degree_bits, const uint64_t* betas,
                               const uint64_t* gammas, const uint64_t* alphas, uint32_t num_challenges,
                               const uint64_t* const* gate_terms, const vpbs_gate_program* program,
                               const uint64_t* public_inputs_hash, uint64_t* d_vals_out);
int vpbs_quotient_commit_values(vpbs_ctx* ctx, const uint64_t* d_vals, uint32_t num_challenges, uint32_t log_n,
                                uint32_t quotient_degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* VPBS_COMMIT_H */
