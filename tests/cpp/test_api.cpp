// C++ host-mirror parity test: vpbs::PolynomialBatch / MerkleTree / fft (include/vpbs_commit.hpp,
// over the C ABI) against the CPU oracle (oracle/oracle.h — test infrastructure).  Written the way
// a plonky2 call site reads; exits non-zero on the first mismatch.  Built and run by
// tests/test_gpu_parity.py::test_cpp_host_mirror on the GPU box.
#include <array>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/vpbs_commit.hpp"
#include "../../oracle/oracle.h"

using vpbs::F;

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main() {
  vpbs::Context ctx(0);
  std::mt19937_64 rng(42);

  const unsigned shapes[5][4] = {{10, 135, 3, 4}, {8, 20, 3, 4}, {5, 3, 1, 0}, {0, 9, 2, 2}, {12, 16, 3, 4}};
  for (auto& shape : shapes) {
    const unsigned log_n = shape[0], ncols = shape[1], rate_bits = shape[2], cap_height = shape[3];
    const std::size_t n = std::size_t(1) << log_n, m = n << rate_bits;
    std::vector<std::vector<F>> values(ncols, std::vector<F>(n));
    for (auto& col : values)
      for (auto& x : col) x = rng();  // non-canonical inputs allowed
    for (int coeffs = 0; coeffs < 2; coeffs++) {
      auto batch = coeffs ? vpbs::PolynomialBatch::from_coeffs(ctx, values, rate_bits, false, cap_height)
                          : vpbs::PolynomialBatch::from_values(ctx, values, rate_bits, false, cap_height);
      std::vector<const uint64_t*> in(ncols);
      for (unsigned c = 0; c < ncols; c++) in[c] = values[c].data();
      std::vector<uint64_t> ocoef(ncols * n), oleaves(m * ncols), odig(8 * (m - (1u << cap_height))),
          ocap(4u << cap_height);
      CHECK(orc_commit(in.data(), ncols, log_n, rate_bits, cap_height, coeffs, nullptr, ocoef.data(),
                       nullptr, oleaves.data(), odig.data(), ocap.data()) == 0);
      CHECK(std::memcmp(batch.merkle_tree.cap.data(), ocap.data(), ocap.size() * 8) == 0);
      CHECK(std::memcmp(batch.merkle_tree.leaves.data(), oleaves.data(), oleaves.size() * 8) == 0);
      CHECK(odig.empty() ||
            std::memcmp(batch.merkle_tree.digests.data(), odig.data(), odig.size() * 8) == 0);
      for (unsigned c = 0; c < ncols; c++)
        CHECK(std::memcmp(batch.polynomials[c].data(), ocoef.data() + c * n, n * 8) == 0);
      // openings: prove + verify on the device, and against the oracle's verifier
      for (std::size_t i : {std::size_t(0), m / 3, m - 1}) {
        auto proof = batch.merkle_tree.prove(i);
        CHECK(vpbs::verify_merkle_proof_to_cap(ctx, batch.merkle_tree.get(i), ncols, i,
                                               batch.merkle_tree.cap, proof));
        CHECK(orc_merkle_verify(batch.merkle_tree.get(i), ncols, i,
                                proof.siblings.empty() ? nullptr : proof.siblings[0].elements,
                                (uint32_t)proof.siblings.size(), ocap.data(), cap_height) == 0);
        // get_lde_values(j) is the natural-order row j
        auto row = batch.get_lde_values(vpbs::reverse_bits(i, log_n + rate_bits));
        CHECK(std::memcmp(row.data(), batch.merkle_tree.get(i), ncols * 8) == 0);
      }
    }
  }
  // fft / ifft round trip and against the oracle
  for (unsigned lg : {0u, 3u, 9u, 14u}) {
    std::vector<F> v(std::size_t(1) << lg);
    for (auto& x : v) x = rng();
    auto ev = vpbs::fft(ctx, v);
    std::vector<uint64_t> o = v;
    orc_fft(o.data(), lg);
    CHECK(ev == o);
    auto back = vpbs::ifft(ctx, ev);
    for (std::size_t i = 0; i < v.size(); i++) CHECK(back[i] == v[i] % ORC_P);
  }
  // resident batch: lazy get / prove / openings; FRI layer + fold; PoW
  {
    const unsigned log_n = 9, ncols = 20, rate_bits = 3, cap_height = 4;
    const std::size_t n = std::size_t(1) << log_n, m = n << rate_bits;
    std::vector<std::vector<F>> values(ncols, std::vector<F>(n));
    for (auto& col : values)
      for (auto& x : col) x = rng();
    vpbs::ResidentBatch rb(ctx, values, rate_bits, cap_height, false);
    auto eager = vpbs::PolynomialBatch::from_values(ctx, values, rate_bits, false, cap_height);
    CHECK(std::memcmp(rb.cap.data(), eager.merkle_tree.cap.data(), rb.cap.size() * 32) == 0);
    for (std::size_t i : {std::size_t(1), m / 2 + 3, m - 1}) {
      auto row = rb.get(i);
      CHECK(std::memcmp(row.data(), eager.merkle_tree.get(i), ncols * 8) == 0);
      auto p1 = rb.prove(i), p2 = eager.merkle_tree.prove(i);
      CHECK(p1.siblings.size() == p2.siblings.size());
      CHECK(std::memcmp(p1.siblings.data(), p2.siblings.data(), p1.siblings.size() * 32) == 0);
    }
    const F zeta[2] = {rng(), rng()};
    auto op = rb.eval_ext2(zeta);
    std::vector<const uint64_t*> cp(ncols);
    for (unsigned c = 0; c < ncols; c++) cp[c] = eager.polynomials[c].data();
    std::vector<uint64_t> want(2 * ncols);
    orc_eval_ext2(cp.data(), ncols, n, zeta, want.data());
    CHECK(op == want);

    std::vector<F> ext(2 * 4096);
    for (auto& x : ext) x = rng() % ORC_P;
    auto tree = vpbs::fri_layer_commit(ctx, ext, 4, 4);
    std::vector<uint64_t> ol(ext.size()), od(8 * (256 - 16)), oc(64);
    CHECK(orc_fri_layer_commit(ext.data(), 4096, 4, 4, ol.data(), od.data(), oc.data()) == 0);
    CHECK(std::memcmp(tree.cap.data(), oc.data(), 64 * 8) == 0 && tree.leaves == ol);
    const F beta[2] = {rng() % ORC_P, rng() % ORC_P};
    auto fold = vpbs::fri_fold(ctx, ext, 4, beta, 33232930569601ULL /* 7^16 */);
    std::vector<uint64_t> fc(512), fv(512);
    orc_fri_fold(ext.data(), 4096, 4, beta, 33232930569601ULL, fc.data(), fv.data());
    CHECK(fold.coeffs == fc && fold.values == fv);

    F st[12];
    for (auto& x : st) x = rng() % ORC_P;
    long long w = vpbs::fri_proof_of_work(ctx, st, 4, 10);
    CHECK(w >= 0);
    uint64_t chk[12];
    std::memcpy(chk, st, sizeof chk);
    chk[4] = (uint64_t)w;
    orc_poseidon(chk);
    CHECK((chk[7] >> 54) == 0);
  }
  // failure behaviour: what plonky2 asserts on
  bool threw = false;
  try {
    vpbs::MerkleTree t(ctx, std::vector<F>(8 * 3), 3, 4);
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    vpbs::MerkleTree t(ctx, std::vector<F>(6 * 3), 3, 1);
  } catch (const std::invalid_argument&) {
    threw = true;
  }
  CHECK(threw);
  std::printf("cpp host mirror ok\n");
  return 0;
}
