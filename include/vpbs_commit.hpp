// vpbs_commit.hpp — header-only C++ host mirror of plonky2 0.2.0's commitment API over the C ABI
// (include/vpbs_commit.h).  The reference's host language is Rust, which this environment cannot
// compile; this mirror keeps plonky2's names, argument meaning and failure behaviour so that a
// call site reads like upstream:
//
//   [P2] plonky2_field/src/fft.rs            fft / ifft                      -> vpbs::fft, vpbs::ifft
//   [P2] plonky2_field/src/polynomial/mod.rs PolynomialCoeffs::coset_fft     -> vpbs::coset_fft
//   [P2] plonky2/src/hash/merkle_tree.rs     MerkleTree::{new,get,prove}     -> vpbs::MerkleTree
//   [P2] plonky2/src/fri/oracle.rs           PolynomialBatch::{from_values,from_coeffs,
//                                            get_lde_values}                 -> vpbs::PolynomialBatch
//
// reached in the reference from prove()/build(), /root/reference/src/vtfhe/ivc_based_vpbs.rs:275,
// :302, :333, :364.  Where plonky2 panics (assert!/unwrap) these throw std::invalid_argument
// (argument errors) or std::runtime_error (device errors).  No CPU fallback exists behind any call.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "vpbs_commit.h"

namespace vpbs {

using F = uint64_t;  // GoldilocksField: #[repr(transparent)] u64
constexpr std::size_t NUM_HASH_OUT_ELTS = 4;
constexpr std::size_t SALT_SIZE = VPBS_SALT_SIZE;
struct HashOut {
  F elements[NUM_HASH_OUT_ELTS];
  bool operator==(const HashOut& o) const {
    for (std::size_t i = 0; i < NUM_HASH_OUT_ELTS; i++)
      if (elements[i] != o.elements[i]) return false;
    return true;
  }
};

inline unsigned log2_strict(std::size_t n) {  // [P2] plonky2_util::log2_strict
  if (n == 0 || (n & (n - 1))) throw std::invalid_argument("Not a power of two: " + std::to_string(n));
  unsigned l = 0;
  while ((std::size_t(1) << l) < n) l++;
  return l;
}
inline std::size_t reverse_bits(std::size_t x, unsigned bits) {  // [P2] plonky2_util::reverse_bits
  std::size_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

class Context {
 public:
  explicit Context(int device = 0) {
    int rc = vpbs_ctx_create(device, &h_);
    if (rc != VPBS_OK) throw std::runtime_error(std::string("vpbs_ctx_create: ") + vpbs_last_error(nullptr));
  }
  ~Context() { vpbs_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  vpbs_ctx* get() const { return h_; }
  // copy threads of the pinned staging ring for pageable host columns (0: driver staging)
  void set_host_threads(unsigned threads) { check(vpbs_ctx_set_host_threads(h_, threads)); }
  // row-range shard of the resident batches created afterwards (one proof over `count` GPUs)
  void set_shard(uint32_t index, uint32_t count) { check(vpbs_ctx_set_shard(h_, index, count)); }
  void check(int rc) const {
    if (rc == VPBS_OK) return;
    std::string msg = vpbs_last_error(h_);
    if (rc == VPBS_ERR_ARG) throw std::invalid_argument(msg);
    throw std::runtime_error("vpbs error " + std::to_string(rc) + ": " + msg);
  }

 private:
  vpbs_ctx* h_ = nullptr;
};

// ---- fft.rs / polynomial/mod.rs ---------------------------------------------------------------
inline std::vector<F> fft(const Context& ctx, std::vector<F> coeffs) {
  ctx.check(vpbs_fft(ctx.get(), coeffs.data(), log2_strict(coeffs.size())));
  return coeffs;
}
inline std::vector<F> ifft(const Context& ctx, std::vector<F> values) {
  ctx.check(vpbs_ifft(ctx.get(), values.data(), log2_strict(values.size())));
  return values;
}
inline std::vector<F> coset_fft(const Context& ctx, std::vector<F> coeffs, F shift = 7) {
  ctx.check(vpbs_coset_fft(ctx.get(), coeffs.data(), log2_strict(coeffs.size()), shift));
  return coeffs;
}

// ---- hash/merkle_tree.rs ------------------------------------------------------------------------
struct MerkleProof {
  std::vector<HashOut> siblings;
};

class MerkleTree {
 public:
  // Row-major leaves (leaf k = leaves[k*leaf_len .. (k+1)*leaf_len)): the flat form of plonky2's
  // Vec<Vec<F>>.
  std::vector<F> leaves;
  std::size_t leaf_len = 0;
  std::vector<HashOut> digests;
  std::vector<HashOut> cap;  // MerkleCap.0

  MerkleTree() = default;
  // MerkleTree::new(leaves, cap_height)
  MerkleTree(const Context& ctx, std::vector<F> leaves_rowmajor, std::size_t leaf_length,
             unsigned cap_height)
      : leaves(std::move(leaves_rowmajor)), leaf_len(leaf_length) {
    const std::size_t nleaves = leaf_len ? leaves.size() / leaf_len : 0;
    const unsigned lg = log2_strict(nleaves);
    if (cap_height > lg)
      throw std::invalid_argument("cap_height=" + std::to_string(cap_height) +
                                  " should be at most log2(leaves.len())=" + std::to_string(lg));
    digests.resize(2 * (nleaves - (std::size_t(1) << cap_height)));
    cap.resize(std::size_t(1) << cap_height);
    ctx.check(vpbs_merkle_new(ctx.get(), leaves.data(), nleaves, (uint32_t)leaf_len, cap_height,
                              digests.empty() ? nullptr : digests[0].elements, cap[0].elements));
  }
  std::size_t num_leaves() const { return leaf_len ? leaves.size() / leaf_len : 0; }
  const F* get(std::size_t i) const { return leaves.data() + i * leaf_len; }
  // MerkleTree::prove: index arithmetic over the unchanged `digests` layout.
  MerkleProof prove(std::size_t leaf_index) const {
    const unsigned cap_height = log2_strict(cap.size());
    const unsigned num_layers = log2_strict(num_leaves()) - cap_height;
    if (leaf_index >> (cap_height + num_layers)) throw std::invalid_argument("leaf_index out of range");
    const std::size_t tree_index = leaf_index >> num_layers;
    const std::size_t tree_len = digests.size() >> cap_height;
    const HashOut* tree = digests.data() + tree_len * tree_index;
    std::size_t pair_index = leaf_index & ((std::size_t(1) << num_layers) - 1);
    MerkleProof p;
    for (unsigned i = 0; i < num_layers; i++) {
      const std::size_t parity = pair_index & 1;
      pair_index >>= 1;
      const std::size_t siblings_index = (pair_index << (i + 1)) + (std::size_t(1) << i) - 1;
      p.siblings.push_back(tree[2 * siblings_index + (1 - parity)]);
    }
    return p;
  }
};

// [P2] hash/merkle_proofs.rs verify_merkle_proof_to_cap (hashing on the device).
inline bool verify_merkle_proof_to_cap(const Context& ctx, const F* leaf, std::size_t leaf_len,
                                       std::size_t leaf_index, const std::vector<HashOut>& cap,
                                       const MerkleProof& proof) {
  HashOut cur;
  ctx.check(vpbs_hash_or_noop_batch(ctx.get(), leaf, 1, (uint32_t)leaf_len, cur.elements));
  std::size_t idx = leaf_index;
  for (const HashOut& sib : proof.siblings) {
    HashOut nxt;
    if (idx & 1) ctx.check(vpbs_two_to_one_batch(ctx.get(), sib.elements, cur.elements, 1, nxt.elements));
    else ctx.check(vpbs_two_to_one_batch(ctx.get(), cur.elements, sib.elements, 1, nxt.elements));
    cur = nxt;
    idx >>= 1;
  }
  return cur == cap[idx];
}

// ---- fri/oracle.rs ------------------------------------------------------------------------------
class PolynomialBatch {
 public:
  std::vector<std::vector<F>> polynomials;  // coefficient vectors
  MerkleTree merkle_tree;
  unsigned degree_log = 0;
  unsigned rate_bits = 0;
  bool blinding = false;
  vpbs_stats stats{};  // plonky2's TimingTree scopes for this commit

  // from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table): `salt` stands in
  // for F::rand_vec (the RNG stays on the host); required iff blinding.
  static PolynomialBatch from_values(const Context& ctx, const std::vector<std::vector<F>>& values,
                                     unsigned rate_bits, bool blinding, unsigned cap_height,
                                     const std::vector<std::vector<F>>* salt = nullptr) {
    return commit(ctx, values, rate_bits, blinding, cap_height, false, salt);
  }
  static PolynomialBatch from_coeffs(const Context& ctx, const std::vector<std::vector<F>>& polys,
                                     unsigned rate_bits, bool blinding, unsigned cap_height,
                                     const std::vector<std::vector<F>>* salt = nullptr) {
    return commit(ctx, polys, rate_bits, blinding, cap_height, true, salt);
  }
  // The same commit spread over several GPUs of this process (vpbs_commit_multi): row ranges per GPU,
  // no GPU-to-GPU traffic, identical outputs.  ctxs: contexts on distinct devices.
  static PolynomialBatch from_values(const std::vector<const Context*>& ctxs,
                                     const std::vector<std::vector<F>>& values, unsigned rate_bits,
                                     bool blinding, unsigned cap_height,
                                     const std::vector<std::vector<F>>* salt = nullptr) {
    if (ctxs.empty() || !ctxs[0]) throw std::invalid_argument("no contexts");
    return commit(*ctxs[0], values, rate_bits, blinding, cap_height, false, salt, &ctxs);
  }
  static PolynomialBatch from_coeffs(const std::vector<const Context*>& ctxs,
                                     const std::vector<std::vector<F>>& polys, unsigned rate_bits,
                                     bool blinding, unsigned cap_height,
                                     const std::vector<std::vector<F>>* salt = nullptr) {
    if (ctxs.empty() || !ctxs[0]) throw std::invalid_argument("no contexts");
    return commit(*ctxs[0], polys, rate_bits, blinding, cap_height, true, salt, &ctxs);
  }
  // get_lde_values(index, step): leaf reverse_bits(index * step) without the salt.
  std::vector<F> get_lde_values(std::size_t index, std::size_t step = 1) const {
    const std::size_t k = reverse_bits(index * step, degree_log + rate_bits);
    const F* row = merkle_tree.get(k);
    return std::vector<F>(row, row + merkle_tree.leaf_len - (blinding ? SALT_SIZE : 0));
  }

 private:
  static PolynomialBatch commit(const Context& ctx, const std::vector<std::vector<F>>& cols,
                                unsigned rate_bits, bool blinding, unsigned cap_height,
                                bool are_coeffs, const std::vector<std::vector<F>>* salt,
                                const std::vector<const Context*>* multi = nullptr) {
    if (cols.empty()) throw std::invalid_argument("empty batch");
    const std::size_t n = cols[0].size();
    const unsigned log_n = log2_strict(n);
    for (auto& c : cols)
      if (c.size() != n) throw std::invalid_argument("Polynomial degrees inconsistent");
    const std::size_t m = n << rate_bits, ncols = cols.size();
    if (blinding && (!salt || salt->size() != SALT_SIZE)) throw std::invalid_argument("blinding needs 4 salt columns");
    PolynomialBatch b;
    b.degree_log = log_n;
    b.rate_bits = rate_bits;
    b.blinding = blinding;
    b.polynomials.assign(ncols, std::vector<F>(n));
    std::vector<const F*> in(ncols), sp(SALT_SIZE);
    std::vector<F*> co(ncols);
    for (std::size_t c = 0; c < ncols; c++) {
      in[c] = cols[c].data();
      co[c] = b.polynomials[c].data();
    }
    if (blinding)
      for (std::size_t s = 0; s < SALT_SIZE; s++) {
        if ((*salt)[s].size() != m) throw std::invalid_argument("salt column length");
        sp[s] = (*salt)[s].data();
      }
    if (cap_height > log_n + rate_bits)
      throw std::invalid_argument("cap_height should be at most log2(leaves.len())");
    MerkleTree& t = b.merkle_tree;
    t.leaf_len = ncols + (blinding ? SALT_SIZE : 0);
    t.leaves.resize(m * t.leaf_len);
    t.digests.resize(2 * (m - (std::size_t(1) << cap_height)));
    t.cap.resize(std::size_t(1) << cap_height);
    if (multi) {
      std::vector<vpbs_ctx*> hs;
      for (const Context* c : *multi) hs.push_back(c ? c->get() : nullptr);
      ctx.check(vpbs_commit_multi(hs.data(), (int)hs.size(), in.data(), (uint32_t)ncols, log_n,
                                  rate_bits, cap_height, are_coeffs ? 1 : 0,
                                  blinding ? sp.data() : nullptr, co.data(), t.leaves.data(),
                                  t.digests.empty() ? nullptr : t.digests[0].elements,
                                  t.cap[0].elements, &b.stats));
      return b;
    }
    ctx.check(vpbs_commit(ctx.get(), in.data(), (uint32_t)ncols, log_n, rate_bits, cap_height,
                          are_coeffs ? 1 : 0, blinding ? sp.data() : nullptr, co.data(),
                          t.leaves.data(), t.digests.empty() ? nullptr : t.digests[0].elements,
                          t.cap[0].elements, &b.stats));
    return b;
  }
};


// ---- PolynomialBatch kept in HBM: MerkleTree::get / prove and openings served on demand ----------
class ResidentBatch {
 public:
  std::vector<HashOut> cap;
  unsigned degree_log = 0, rate_bits = 0;
  std::size_t ncols = 0, width = 0;

  ResidentBatch(const Context& ctx, const std::vector<std::vector<F>>& cols, unsigned rate_bits_,
                unsigned cap_height, bool inputs_are_coeffs)
      : ctx_(ctx) {
    if (cols.empty()) throw std::invalid_argument("empty batch");
    degree_log = log2_strict(cols[0].size());
    rate_bits = rate_bits_;
    ncols = width = cols.size();
    if (cap_height > degree_log + rate_bits)
      throw std::invalid_argument("cap_height should be at most log2(leaves.len())");
    std::vector<const F*> in(ncols);
    for (std::size_t c = 0; c < ncols; c++) in[c] = cols[c].data();
    cap.resize(std::size_t(1) << cap_height);
    ctx.check(vpbs_batch_commit(ctx.get(), in.data(), (uint32_t)ncols, degree_log, rate_bits,
                                cap_height, inputs_are_coeffs ? 1 : 0, nullptr, cap[0].elements, &h_,
                                nullptr));
  }
  ~ResidentBatch() { vpbs_batch_destroy(h_); }
  ResidentBatch(const ResidentBatch&) = delete;
  ResidentBatch& operator=(const ResidentBatch&) = delete;

  std::vector<F> get(std::size_t leaf_index) const {  // MerkleTree::get
    std::vector<F> row(width);
    uint64_t idx = leaf_index;
    ctx_.check(vpbs_batch_get_leaves(h_, &idx, 1, row.data()));
    return row;
  }
  MerkleProof prove(std::size_t leaf_index) const {  // MerkleTree::prove
    MerkleProof p;
    p.siblings.resize(degree_log + rate_bits - log2_strict(cap.size()));
    uint64_t idx = leaf_index;
    ctx_.check(vpbs_batch_prove(h_, &idx, 1, p.siblings.empty() ? nullptr : p.siblings[0].elements));
    return p;
  }
  std::vector<F> get_lde_values(std::size_t index, std::size_t step = 1) const {
    return get(reverse_bits(index * step, degree_log + rate_bits));
  }
  // openings of every polynomial at one point of F[X]/(X^2 - 7): ncols x (re, im)
  std::vector<F> eval_ext2(const F (&point)[2]) const {
    std::vector<F> out(2 * ncols);
    ctx_.check(vpbs_batch_eval_ext2(h_, point, 1, out.data()));
    return out;
  }
  // get_lde_values(first + k * step) for k < count, salt dropped: count x ncols, row-major
  std::vector<F> get_lde_rows(std::size_t first, std::size_t step, std::size_t count) const {
    std::vector<F> out(count * ncols);
    ctx_.check(vpbs_batch_get_lde_rows(h_, first, step, count, out.data()));
    return out;
  }
  // PolynomialBatch.polynomials of the resident batch (coefficient columns), copied to the host
  std::vector<std::vector<F>> coefficients() const {
    std::vector<std::vector<F>> out(ncols, std::vector<F>(std::size_t(1) << degree_log));
    std::vector<F*> ptrs(ncols);
    for (std::size_t c = 0; c < ncols; c++) ptrs[c] = out[c].data();
    ctx_.check(vpbs_batch_download(h_, ptrs.data(), nullptr, nullptr));
    return out;
  }
  vpbs_batch* handle() const { return h_; }
  const Context& context() const { return ctx_; }

 private:
  friend class Sigmas;
  friend class GateProgram;
  ResidentBatch(const Context& ctx, vpbs_batch* h, std::vector<HashOut> cap_, unsigned degree_log_,
                unsigned rate_bits_, std::size_t ncols_)
      : cap(std::move(cap_)), degree_log(degree_log_), rate_bits(rate_bits_), ncols(ncols_), width(ncols_),
        ctx_(ctx), h_(h) {}
  const Context& ctx_;
  vpbs_batch* h_ = nullptr;
};

// plonk/proof.rs OpeningSet::new over all FRI oracles in one round trip: per batch ncols x (re, im)
inline std::vector<std::vector<F>> open_all_at_point(const std::vector<const ResidentBatch*>& batches,
                                                     const F (&point)[2]) {
  std::vector<vpbs_batch*> hs;
  std::vector<std::vector<F>> outs;
  std::vector<F*> ptrs;
  for (const ResidentBatch* b : batches) {
    hs.push_back(b->handle());
    outs.emplace_back(2 * b->ncols);
  }
  for (auto& o : outs) ptrs.push_back(o.data());
  batches.at(0)->context().check(vpbs_batches_eval_ext2(hs.data(), (uint32_t)hs.size(), point, 1, ptrs.data()));
  return outs;
}
// fri/prover.rs fri_prover_query_round (initial_trees_proof): (row, Merkle path) of every oracle at
// one leaf index, one round trip for all of them
inline std::vector<std::pair<std::vector<F>, MerkleProof>> open_all_at_leaf(
    const std::vector<const ResidentBatch*>& batches, std::size_t leaf_index) {
  const ResidentBatch* b0 = batches.at(0);
  const std::size_t layers = b0->degree_log + b0->rate_bits - log2_strict(b0->cap.size());
  std::vector<vpbs_batch*> hs;
  std::vector<std::pair<std::vector<F>, MerkleProof>> out;
  for (const ResidentBatch* b : batches) {
    hs.push_back(b->handle());
    out.emplace_back(std::vector<F>(b->width), MerkleProof{std::vector<HashOut>(layers)});
  }
  std::vector<F*> rows, sibs;
  for (auto& o : out) {
    rows.push_back(o.first.data());
    sibs.push_back(layers ? o.second.siblings[0].elements : nullptr);
  }
  if (!layers) {  // all-cap trees have no paths: rows only
    uint64_t idx = leaf_index;
    for (std::size_t k = 0; k < hs.size(); k++) b0->context().check(vpbs_batch_get_leaves(hs[k], &idx, 1, rows[k]));
    return out;
  }
  uint64_t idx = leaf_index;
  b0->context().check(vpbs_batches_open(hs.data(), (uint32_t)hs.size(), &idx, 1, rows.data(), sibs.data()));
  return out;
}

// ---- plonk/prover.rs, step 4-5: Z and partial products of the permutation argument ----------------
// The sigma polynomials' values and the coset shifts k_is of one circuit, resident in HBM.
class Sigmas {
 public:
  Sigmas(const Context& ctx, const std::vector<std::vector<F>>& sigma_cols, const std::vector<F>& k_is)
      : ctx_(ctx), num_routed_(sigma_cols.size()) {
    if (sigma_cols.empty() || k_is.size() != sigma_cols.size())
      throw std::invalid_argument("one k_i per routed wire");
    std::vector<const F*> in(sigma_cols.size());
    for (std::size_t j = 0; j < in.size(); j++) in[j] = sigma_cols[j].data();
    degree_log_ = log2_strict(sigma_cols[0].size());
    ctx.check(vpbs_sigmas_upload(ctx.get(), in.data(), k_is.data(), (uint32_t)in.size(), degree_log_, &h_));
  }
  ~Sigmas() { vpbs_sigmas_destroy(h_); }
  Sigmas(const Sigmas&) = delete;
  Sigmas& operator=(const Sigmas&) = delete;

  // all_wires_permutation_partial_products on host wire columns: num_challenges * K columns in commit
  // order (the Zs first), K = ceil(num_routed / max_degree)
  std::vector<std::vector<F>> partial_products(const std::vector<std::vector<F>>& wire_cols,
                                               const std::vector<F>& betas, const std::vector<F>& gammas,
                                               unsigned max_degree) const {
    if (wire_cols.size() != num_routed_ || betas.size() != gammas.size() || betas.empty())
      throw std::invalid_argument("wire_cols must be num_routed columns; one gamma per beta");
    const std::size_t n = std::size_t(1) << degree_log_, K = (num_routed_ + max_degree - 1) / max_degree;
    std::vector<std::vector<F>> out(betas.size() * K, std::vector<F>(n));
    std::vector<const F*> in(num_routed_);
    std::vector<F*> op(out.size());
    for (std::size_t j = 0; j < num_routed_; j++) in[j] = wire_cols[j].data();
    for (std::size_t c = 0; c < out.size(); c++) op[c] = out[c].data();
    ctx_.check(vpbs_zs_partial_products(ctx_.get(), in.data(), h_, max_degree, betas.data(), gammas.data(),
                                        (uint32_t)betas.size(), op.data()));
    return out;
  }
  // the same from a resident wires batch, committed at once as a new resident batch (steps 4-5)
  ResidentBatch commit_partial_products(const ResidentBatch& wires, const std::vector<F>& betas,
                                        const std::vector<F>& gammas, unsigned max_degree, unsigned rate_bits,
                                        unsigned cap_height) const {
    if (betas.size() != gammas.size() || betas.empty()) throw std::invalid_argument("one gamma per beta");
    const std::size_t K = (num_routed_ + max_degree - 1) / max_degree;
    std::vector<HashOut> cap(std::size_t(1) << cap_height);
    vpbs_batch* h = nullptr;
    ctx_.check(vpbs_batch_zs_partial_products(wires.handle(), h_, max_degree, betas.data(), gammas.data(),
                                              (uint32_t)betas.size(), rate_bits, cap_height, cap[0].elements,
                                              &h, nullptr));
    return ResidentBatch(ctx_, h, std::move(cap), degree_log_, rate_bits, betas.size() * K);
  }

 private:
  const Context& ctx_;
  vpbs_sigmas* h_ = nullptr;
  std::size_t num_routed_ = 0;
  unsigned degree_log_ = 0;
};

// ---- plonk/prover.rs, steps 6-7: compute_quotient_polys + the quotient commit ----------------------
// The circuit's gate constraints as a straight-line program on the device (format: vpbs_commit.h).
class GateProgram {
 public:
  GateProgram(const Context& ctx, const std::vector<uint64_t>& code, const std::vector<F>& imms,
              unsigned nregs, unsigned num_constraints)
      : ctx_(ctx) {
    ctx.check(vpbs_gate_program_upload(ctx.get(), code.data(), (uint32_t)code.size(), imms.data(),
                                       (uint32_t)imms.size(), nregs, num_constraints, &h_));
  }
  ~GateProgram() { vpbs_gate_program_destroy(h_); }
  GateProgram(const GateProgram&) = delete;
  GateProgram& operator=(const GateProgram&) = delete;
  const vpbs_gate_program* handle() const { return h_; }

  // vanishing terms of the permutation argument + gate constraints (from `program`, or as alpha-reduced
  // values gate_terms[c][i] over the quotient domain, or none), reduced with the alphas, divided by
  // Z_H, coset_ifft, chunks of n, committed: the quotient batch (num_challenges << quotient_degree_bits
  // columns)
  static ResidentBatch quotient_polys(const ResidentBatch& constants_sigmas, unsigned sigmas_first_col,
                                      const ResidentBatch& wires, const ResidentBatch& zs_pp,
                                      const std::vector<F>& k_is, unsigned max_degree,
                                      unsigned quotient_degree_bits, const std::vector<F>& betas,
                                      const std::vector<F>& gammas, const std::vector<F>& alphas,
                                      unsigned rate_bits, unsigned cap_height,
                                      const GateProgram* program = nullptr,
                                      const F* public_inputs_hash = nullptr,
                                      const std::vector<std::vector<F>>* gate_terms = nullptr) {
    if (betas.size() != gammas.size() || betas.size() != alphas.size() || betas.empty())
      throw std::invalid_argument("one beta, gamma and alpha per challenge");
    std::vector<const F*> gt;
    if (gate_terms)
      for (const auto& v : *gate_terms) gt.push_back(v.data());
    std::vector<HashOut> cap(std::size_t(1) << cap_height);
    vpbs_batch* h = nullptr;
    const Context& ctx = wires.context();
    ctx.check(vpbs_batch_quotient_polys(constants_sigmas.handle(), sigmas_first_col, wires.handle(),
                                        zs_pp.handle(), k_is.data(), (uint32_t)k_is.size(), max_degree,
                                        quotient_degree_bits, betas.data(), gammas.data(), alphas.data(),
                                        (uint32_t)betas.size(), gate_terms ? gt.data() : nullptr,
                                        program ? program->handle() : nullptr, public_inputs_hash, rate_bits,
                                        cap_height, cap[0].elements, &h, nullptr));
    return ResidentBatch(ctx, h, std::move(cap), wires.degree_log, rate_bits,
                         betas.size() << quotient_degree_bits);
  }

 private:
  const Context& ctx_;
  vpbs_gate_program* h_ = nullptr;
};

// ---- fri/prover.rs ------------------------------------------------------------------------------
// fri_committed_trees, one layer: values are (re, im) pairs, flat.
inline MerkleTree fri_layer_commit(const Context& ctx, const std::vector<F>& values_ext,
                                   unsigned arity_bits, unsigned cap_height) {
  const std::size_t len = values_ext.size() / 2;
  const unsigned lg = log2_strict(len);
  if (arity_bits > lg || cap_height > lg - arity_bits)
    throw std::invalid_argument("cap_height should be at most log2(leaves.len())");
  MerkleTree t;
  t.leaf_len = std::size_t(2) << arity_bits;
  t.leaves.resize(2 * len);
  t.digests.resize(2 * ((len >> arity_bits) - (std::size_t(1) << cap_height)));
  t.cap.resize(std::size_t(1) << cap_height);
  ctx.check(vpbs_fri_layer_commit(ctx.get(), values_ext.data(), len, arity_bits, cap_height,
                                  t.leaves.data(), t.digests.empty() ? nullptr : t.digests[0].elements,
                                  t.cap[0].elements));
  return t;
}
struct FriFold {
  std::vector<F> coeffs, values;  // (re, im) pairs
};
inline FriFold fri_fold(const Context& ctx, const std::vector<F>& coeffs_ext, unsigned arity_bits,
                        const F (&beta)[2], F shift_next) {
  const std::size_t len = coeffs_ext.size() / 2;
  log2_strict(len);
  FriFold r;
  r.coeffs.resize(2 * (len >> arity_bits));
  r.values.resize(2 * (len >> arity_bits));
  ctx.check(vpbs_fri_fold(ctx.get(), coeffs_ext.data(), len, arity_bits, beta, shift_next,
                          r.coeffs.data(), r.values.data()));
  return r;
}
// fri_proof_of_work: smallest witness in [first, first + count), or -1 if none.
inline long long fri_proof_of_work(const Context& ctx, const F (&state)[12], unsigned witness_pos,
                                   unsigned min_leading_zeros, uint64_t first = 0,
                                   uint64_t count = uint64_t(1) << 32) {
  uint64_t w = 0;
  int found = 0;
  ctx.check(vpbs_pow_grind(ctx.get(), state, witness_pos, 7, min_leading_zeros, first, count, &w, &found));
  return found ? (long long)w : -1;
}

// fri_committed_trees as one device-resident chain (vpbs_fri_*): started either from the final
// polynomial's coefficients or, as fri/oracle.rs prove_openings does, from the committed batches.
struct FriPolynomialInfo {
  uint32_t oracle_index, polynomial_index;
};
class FriChain {
 public:
  FriChain(const Context& ctx, const std::vector<F>& final_poly_coeffs_ext, unsigned rate_bits)
      : ctx_(ctx), rate_bits_(rate_bits) {
    log2_strict(final_poly_coeffs_ext.size() / 2);
    ctx.check(vpbs_fri_begin(ctx.get(), final_poly_coeffs_ext.data(), final_poly_coeffs_ext.size() / 2,
                             rate_bits, &h_));
    len_ = (final_poly_coeffs_ext.size() / 2) << rate_bits;
  }
  // batches[b]: the polynomials opened at points[2b], points[2b+1]  ([P2] FriInstanceInfo)
  FriChain(const std::vector<const ResidentBatch*>& oracles,
           const std::vector<std::vector<FriPolynomialInfo>>& batches, const std::vector<F>& points,
           const F (&alpha)[2], unsigned rate_bits)
      : ctx_(oracles.at(0)->context()), rate_bits_(rate_bits) {
    if (points.size() != 2 * batches.size() || batches.empty())
      throw std::invalid_argument("one opening point per FRI batch");
    std::vector<vpbs_batch*> hs;
    for (const ResidentBatch* o : oracles) hs.push_back(o->handle());
    std::vector<uint32_t> sizes, refs;
    for (const auto& b : batches) {
      sizes.push_back((uint32_t)b.size());
      for (const auto& r : b) {
        refs.push_back(r.oracle_index);
        refs.push_back(r.polynomial_index);
      }
    }
    ctx_.check(vpbs_fri_begin_openings(ctx_.get(), hs.data(), (uint32_t)hs.size(), sizes.data(),
                                       (uint32_t)sizes.size(), refs.data(), points.data(), alpha, rate_bits,
                                       &h_));
    len_ = (std::size_t(1) << oracles[0]->degree_log) << rate_bits;
  }
  ~FriChain() { vpbs_fri_destroy(h_); }
  FriChain(const FriChain&) = delete;
  FriChain& operator=(const FriChain&) = delete;

  std::vector<HashOut> commit_layer(unsigned arity_bits, unsigned cap_height) {
    std::vector<HashOut> cap(std::size_t(1) << cap_height);
    ctx_.check(vpbs_fri_commit_layer(h_, arity_bits, cap_height, cap[0].elements));
    pending_ = arity_bits;
    return cap;
  }
  void fold(const F (&beta)[2]) {
    ctx_.check(vpbs_fri_fold_layer(h_, beta));
    len_ >>= pending_;
  }
  std::vector<F> final_poly() const {  // (len >> rate_bits) extension coefficients
    std::vector<F> out(2 * (len_ >> rate_bits_));
    ctx_.check(vpbs_fri_final_poly(h_, rate_bits_, out.data()));
    return out;
  }

 private:
  const Context& ctx_;
  vpbs_fri* h_ = nullptr;
  unsigned rate_bits_ = 0, pending_ = 0;
  std::size_t len_ = 0;
};

}  // namespace vpbs
