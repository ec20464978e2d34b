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
