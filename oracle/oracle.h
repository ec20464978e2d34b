/*
 * oracle.h — CPU restatement of plonky2 0.2.0's polynomial-commitment path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (libvpbs_commit.so) never links, loads or calls this file.
 *
 * What it restates: the algorithm lives in crates that are NOT under /root/reference
 * (plonky2 0.2.0, plonky2_field 0.2.0, plonky2_util 0.2.0 — /root/reference/Cargo.toml:7,
 * Cargo.lock:371-374, 396-399, 421-424).  Each function below names the upstream item it
 * follows ("[P2] path::item") and the reference call site that reaches it.
 *
 * Parity status: the plonky2 crates cannot be built here (no Rust toolchain), so this oracle is
 * pinned by (1) plonky2's published Poseidon known-answer vectors, (2) the reference's own
 * Goldilocks NTT golden vectors /root/reference/src/ntt/params_{8..2048}.rs (TESTG/TESTGHAT,
 * ROOTS, INVROOTS, NINV) and (3) an independent big-integer Python model (oracle/model.py).
 * No reference test pins an LDE value, digest or cap, so at the commit boundary itself the
 * reference is "parity unpinned" (SURVEY.md §8(c)); the conventions (overwrite sponge,
 * hash_or_noop threshold, digests layout, bit-reversed leaves, coset shift 7) follow upstream's
 * definitions as restated in SURVEY.md §8(a).
 */
#ifndef VPBS_ORACLE_H
#define VPBS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_P 0xFFFFFFFF00000001ULL

/* [P2] plonky2_field/src/goldilocks_field.rs — GoldilocksField arithmetic (canonical results). */
uint64_t orc_gl_add(uint64_t a, uint64_t b);
uint64_t orc_gl_sub(uint64_t a, uint64_t b);
uint64_t orc_gl_mul(uint64_t a, uint64_t b);
uint64_t orc_gl_pow(uint64_t a, uint64_t e);
uint64_t orc_gl_inv(uint64_t a);
/* [P2] Field::primitive_root_of_unity(n_log) = POWER_OF_TWO_GENERATOR^(2^(32-n_log)). */
uint64_t orc_primitive_root_of_unity(unsigned n_log);

/* [P2] plonky2_field/src/fft.rs — fft_with_options / ifft_with_options, natural order in & out. */
void orc_fft(uint64_t* v, unsigned log_n);
void orc_ifft(uint64_t* v, unsigned log_n);
/* [P2] polynomial/mod.rs — PolynomialCoeffs::coset_fft_with_options(shift, ..). */
void orc_coset_fft(uint64_t* v, unsigned log_n, uint64_t shift);
/* [P2] PolynomialCoeffs::lde(rate_bits) then coset_fft(F::coset_shift()=7): out has n<<rate_bits. */
void orc_lde(const uint64_t* coeffs, unsigned log_n, unsigned rate_bits, uint64_t* out);

/* [P2] plonky2/src/hash/poseidon.rs + poseidon_goldilocks.rs — width-12 permutation. */
void orc_poseidon_round_constants(uint64_t out[360]);
void orc_poseidon(uint64_t state[12]);
/* [P2] hash/hashing.rs hash_n_to_m_no_pad (m=4), plonk/config.rs Hasher::hash_or_noop, compress. */
void orc_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]);
void orc_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]);
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);

/* [P2] hash/merkle_tree.rs MerkleTree::new: leaves row-major nleaves x leaf_len.
 * digests: 2*(nleaves - 2^cap_height) x 4 in plonky2's layout; cap: 2^cap_height x 4.
 * Returns 0, or -1 on bad arguments (the conditions plonky2 asserts on). */
int orc_merkle_new(const uint64_t* leaves, uint64_t nleaves, uint32_t leaf_len, uint32_t cap_height,
                   uint64_t* digests, uint64_t* cap);
/* [P2] MerkleTree::prove(leaf_index): siblings_out gets (log2 nleaves - cap_height) x 4. */
int orc_merkle_prove(const uint64_t* digests, uint64_t nleaves, uint32_t cap_height,
                     uint64_t leaf_index, uint64_t* siblings_out);
/* [P2] merkle_proofs.rs verify_merkle_proof_to_cap: returns 0 iff the path hashes to cap. */
int orc_merkle_verify(const uint64_t* leaf, uint32_t leaf_len, uint64_t leaf_index,
                      const uint64_t* siblings, uint32_t nsiblings, const uint64_t* cap,
                      uint32_t cap_height);

/* [P2] fri/oracle.rs PolynomialBatch::from_values / from_coeffs.
 * cols: ncols pointers to n=2^log_n values (or coefficients).  salt_cols: NULL or 4 pointers of
 * n<<rate_bits elements (the blinding columns, natural LDE order, as lde_values() chains them).
 * coeffs_out: ncols*n (column-major) or NULL; lde_cols_out: ncols*(n<<r) natural order,
 * column-major, or NULL; leaves_out: m x (ncols+salt) row-major in leaf (bit-reversed) order;
 * digests_out / cap_out as orc_merkle_new. */
int orc_commit(const uint64_t* const* cols, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, int inputs_are_coeffs, const uint64_t* const* salt_cols,
               uint64_t* coeffs_out, uint64_t* lde_cols_out, uint64_t* leaves_out,
               uint64_t* digests_out, uint64_t* cap_out);

/* [P2] plonky2_field/src/extension/quadratic.rs QuadraticExtension<GoldilocksField> (W = 7:
 * F[X]/(X^2 - 7)) and polynomial/mod.rs PolynomialCoeffs::eval / to_extension().eval(zeta), as used
 * by plonk/proof.rs OpeningSet::new ("construct the opening set", reached from prove(),
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:302): out[c] = sum_j cols[c][j] * x^j for the
 * extension point x = (x[0], x[1]); out is ncols x 2. */
void orc_eval_ext2(const uint64_t* const* cols, uint32_t ncols, uint64_t n, const uint64_t x[2],
                   uint64_t* out);

/* [P2] plonky2/src/fri/prover.rs fri_committed_trees, one reduction layer (arity = 2^arity_bits) over
 * D = 2 extension elements stored as (re, im) pairs:
 *   commit: reverse_index_bits_in_place(values); leaves = chunks of `arity` values, flattened;
 *           MerkleTree::new(leaves, cap_height).  leaves_out: (len/arity) x (2*arity).
 *   fold:   coeffs' = chunks_exact(arity).map(reduce_with_powers(chunk, beta)) (= sum_i chunk[i] beta^i);
 *           values' = coeffs'.coset_fft(shift_next) with shift_next = shift^arity (natural order). */
int orc_fri_layer_commit(const uint64_t* values_ext, uint64_t len, uint32_t arity_bits,
                         uint32_t cap_height, uint64_t* leaves_out, uint64_t* digests_out,
                         uint64_t* cap_out);
void orc_fri_fold(const uint64_t* coeffs_ext, uint64_t len, uint32_t arity_bits,
                  const uint64_t beta[2], uint64_t shift_next, uint64_t* coeffs_out,
                  uint64_t* values_out);

/* [P2] plonky2/src/fri/oracle.rs PolynomialBatch::prove_openings from its first line to final_poly
 * (the polynomial FRI is run on), with util/reducing.rs ReducingFactor::{reduce_polys_base,
 * shift_poly} and polynomial/division.rs divide_by_linear; reached from prove(),
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:302.  polys: the coefficient polynomials of all FRI
 * batches back to back (batch_sizes[b] pointers to n base-field coefficients each); points: the
 * opening point of every batch (2 words each); alpha: the extension challenge.  out: n x 2.
 *   F_b = sum_j alpha^j f_bj;  Q_b = (F_b(X) - F_b(z_b)) / (X - z_b), padded with a zero;
 *   final_poly = final_poly * alpha^(batch_sizes[b]) + Q_b. */
int orc_fri_final_poly(const uint64_t* const* polys, const uint32_t* batch_sizes, uint32_t nbatches,
                       uint64_t n, const uint64_t* points, const uint64_t alpha[2], uint64_t* out);

/* [P2] plonky2/src/plonk/prover.rs wires_permutation_partial_products_and_zs with
 * util/partial_products.rs quotient_chunk_products / partial_products_and_z_gx, for ONE challenge
 * pair (beta, gamma) — step 4 of prove() ("compute partial products"), reached from
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:302,333,364:
 *   row i (x_i = w_n^i):  q_j = (wire_j + beta k_j x_i + gamma) / (wire_j + beta sigma_j(x_i) + gamma)
 *   chunk products c_k = prod of q_j over chunks of max_degree routed wires
 *   Z(x_0) = 1,  pp_k(x_i) = Z(x_i) c_0 .. c_k,  Z(x_{i+1}) = Z(x_i) c_0 .. c_{K-1}
 *  wires / sigmas: num_routed pointers to n = 2^log_n values each (wires[j][i] =
 *  witness.get_wire(i, j), sigmas[j][i] = prover_data.sigmas[i][j]); k_is: num_routed coset shifts.
 *  out: K = ceil(num_routed / max_degree) columns of n, column-major: out[0] = Z, out[1 + k] =
 *  partial product k for k < K - 1 (the order prove() commits them in: Zs at the front).
 * Returns 0, -1 on bad arguments, -2 if some denominator is zero (upstream panics there). */
int orc_zs_partial_products(const uint64_t* const* wires, const uint64_t* const* sigmas,
                            const uint64_t* k_is, uint32_t num_routed, uint32_t log_n,
                            uint32_t max_degree, uint64_t beta, uint64_t gamma, uint64_t* out);

/* [P2] plonk/prover.rs compute_quotient_polys with plonk/vanishing_poly.rs
 * eval_vanishing_poly_base_batch restricted to the gate-independent terms (Z(1) = 1 and the
 * partial-product checks of the permutation argument), reduce_with_powers_multi over alphas, division
 * by Z_H on the coset, coset_ifft and the split into chunks of n coefficients; the gate constraints
 * enter as already alpha-reduced values gate_terms[c][i] (NULL: none).  Layouts: see oracle.c. */
int orc_quotient_polys(const uint64_t* const* wires, const uint64_t* const* sigmas,
                       const uint64_t* const* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                       uint32_t log_n, uint32_t max_degree, uint32_t qdb, const uint64_t* betas,
                       const uint64_t* gammas, const uint64_t* alphas, uint32_t nc,
                       const uint64_t* const* gate_terms, uint64_t* out);

/* [P2] evaluate_gate_constraints_base_batch with the gates given as a program (format:
 * include/vpbs_commit.h): out[c][i] = alpha_c-reduced gate constraints at point i of the quotient
 * domain, from the coefficient columns of all wires and of the constants/sigmas batch. */
int orc_gate_program_eval(const uint64_t* code, uint32_t ncode, const uint64_t* imms, uint32_t nimm,
                          uint32_t nregs, uint32_t num_constraints, const uint64_t* const* wires,
                          uint32_t nwires, const uint64_t* const* cs, uint32_t ncs, uint32_t log_n,
                          uint32_t qdb, const uint64_t pih[4], const uint64_t* alphas, uint32_t nc,
                          uint64_t* const* out);

/* SIMD width of the hashing path: 0 = widest the CPU supports (default), 1 = scalar — the naive
 * restatement, which is the checker for the other two — 4 = AVX2, 8 = AVX-512.  The SIMD paths
 * evaluate the SAME permutation on 4 / 8 independent leaves or tree nodes per call
 * (poseidon_simd.inc) so that the CPU arm of bench.py runs at a speed comparable with plonky2's
 * hand-vectorised x86 Poseidon; results are bit-identical (tests/test_oracle_golden.py). */
void orc_set_simd(int width);
int orc_get_simd(void);
void orc_poseidon_batch(uint64_t* states, uint64_t count);

/* Threads used by the parallel regions (mirrors rayon's pool). */
void orc_set_threads(int n);
int orc_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
