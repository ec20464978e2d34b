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
    {  // the same through the all-oracles-at-once forms (OpeningSet::new / initial_trees_proof)
      vpbs::ResidentBatch rb2(ctx, values, rate_bits, cap_height, true);
      auto all = vpbs::open_all_at_point({&rb, &rb2}, zeta);
      CHECK(all.size() == 2 && all[0] == want && all[1] == rb2.eval_ext2(zeta));
      auto opened = vpbs::open_all_at_leaf({&rb, &rb2}, m / 2 + 3);
      CHECK(opened[0].first == rb.get(m / 2 + 3) && opened[1].first == rb2.get(m / 2 + 3));
      auto q1 = rb.prove(m / 2 + 3), q2 = rb2.prove(m / 2 + 3);
      CHECK(std::memcmp(opened[0].second.siblings.data(), q1.siblings.data(), q1.siblings.size() * 32) == 0);
      CHECK(std::memcmp(opened[1].second.siblings.data(), q2.siblings.data(), q2.siblings.size() * 32) == 0);
    }

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
  // steps 4-5 and 9 of prove() through the mirror: Z / partial products (host and resident forms), then
  // prove_openings + one FRI layer from the resident batches
  {
    const unsigned log_n = 8, num_routed = 12, ncols = 15, max_degree = 5, rate_bits = 3, cap_height = 2;
    const std::size_t n = std::size_t(1) << log_n, K = (num_routed + max_degree - 1) / max_degree;
    std::vector<std::vector<F>> wires(ncols, std::vector<F>(n)), sig(num_routed, std::vector<F>(n));
    for (auto& col : wires)
      for (auto& x : col) x = rng();
    for (auto& col : sig)
      for (auto& x : col) x = rng();
    std::vector<F> k_is(num_routed), betas = {rng() % ORC_P, rng() % ORC_P}, gammas = {rng() % ORC_P, rng() % ORC_P};
    for (unsigned j = 0; j < num_routed; j++) k_is[j] = j ? orc_gl_mul(k_is[j - 1], 7) : 1;
    vpbs::Sigmas sigmas(ctx, sig, k_is);
    std::vector<std::vector<F>> routed(wires.begin(), wires.begin() + num_routed);
    auto zs = sigmas.partial_products(routed, betas, gammas, max_degree);
    CHECK(zs.size() == 2 * K);
    std::vector<const uint64_t*> wp(num_routed), sp(num_routed);
    for (unsigned j = 0; j < num_routed; j++) {
      wp[j] = wires[j].data();
      sp[j] = sig[j].data();
    }
    for (unsigned c = 0; c < 2; c++) {
      std::vector<uint64_t> want(K * n);
      CHECK(orc_zs_partial_products(wp.data(), sp.data(), k_is.data(), num_routed, log_n, max_degree, betas[c],
                                    gammas[c], want.data()) == 0);
      CHECK(std::memcmp(zs[c].data(), want.data(), n * 8) == 0);
      for (std::size_t k = 0; k + 1 < K; k++)
        CHECK(std::memcmp(zs[2 + c * (K - 1) + k].data(), want.data() + (1 + k) * n, n * 8) == 0);
    }
    vpbs::ResidentBatch wb(ctx, wires, rate_bits, cap_height, false);
    auto zb = sigmas.commit_partial_products(wb, betas, gammas, max_degree, rate_bits, cap_height);
    auto zref = vpbs::PolynomialBatch::from_values(ctx, zs, rate_bits, false, cap_height);
    CHECK(std::memcmp(zb.cap.data(), zref.merkle_tree.cap.data(), zb.cap.size() * 32) == 0);
    auto rows = zb.get_lde_rows(3, 2, 4);
    for (std::size_t k = 0; k < 4; k++)
      CHECK(std::memcmp(rows.data() + k * zb.ncols, zref.get_lde_values(3 + 2 * k).data(), zb.ncols * 8) == 0);

    // steps 6-7: the quotient from the resident batches, gate constraints from a program
    // (constraint 0: wire0 * wire1 - wire2 + public_inputs_hash[1], filter: the first sigma column)
    {
      const unsigned qdb = 2;
      vpbs::ResidentBatch cb(ctx, sig, rate_bits, cap_height, false);
      const std::vector<F> alphas = {rng() % ORC_P, rng() % ORC_P};
      const F pih[4] = {rng() % ORC_P, rng() % ORC_P, rng() % ORC_P, rng() % ORC_P};
      auto ins = [](unsigned op, unsigned dst, unsigned ka, unsigned ia, unsigned kb, unsigned ib) {
        return uint64_t(op) | uint64_t(dst) << 8 | uint64_t(ka) << 16 | uint64_t(kb) << 20 | uint64_t(ia) << 24 |
               uint64_t(ib) << 40;
      };
      const std::vector<uint64_t> code = {ins(2, 0, 1, 0, 1, 1), ins(1, 0, 0, 0, 1, 2), ins(0, 0, 0, 0, 4, 1),
                                          ins(3, 0, 0, 0, 0, 0), ins(4, 0, 2, 0, 0, 0)};
      vpbs::GateProgram prog(ctx, code, {}, 1, 1);
      auto qb = vpbs::GateProgram::quotient_polys(cb, 0, wb, zb, k_is, max_degree, qdb, betas, gammas, alphas,
                                                  rate_bits, cap_height, &prog, pih);
      auto wcoef = wb.coefficients(), scoef = cb.coefficients(), zcoef = zb.coefficients();
      std::vector<const uint64_t*> wq, sq, zq;
      for (auto& c : wcoef) wq.push_back(c.data());
      for (auto& c : scoef) sq.push_back(c.data());
      for (auto& c : zcoef) zq.push_back(c.data());
      const std::size_t q = n << qdb;
      std::vector<uint64_t> g0(q), g1(q), want((2u << qdb) * n);
      uint64_t* gout[2] = {g0.data(), g1.data()};
      CHECK(orc_gate_program_eval(code.data(), (uint32_t)code.size(), nullptr, 0, 1, 1, wq.data(), ncols, sq.data(),
                                  num_routed, log_n, qdb, pih, alphas.data(), 2, gout) == 0);
      const uint64_t* gin[2] = {g0.data(), g1.data()};
      CHECK(orc_quotient_polys(wq.data(), sq.data(), zq.data(), k_is.data(), num_routed, log_n, max_degree, qdb,
                               betas.data(), gammas.data(), alphas.data(), 2, gin, want.data()) == 0);
      auto got = qb.coefficients();
      CHECK(got.size() == (2u << qdb));
      for (std::size_t c = 0; c < got.size(); c++) CHECK(std::memcmp(got[c].data(), want.data() + c * n, n * 8) == 0);
    }

    // prove_openings: every polynomial of both batches at zeta, the two Zs at g zeta
    std::vector<std::vector<vpbs::FriPolynomialInfo>> fb(2);
    for (uint32_t j = 0; j < ncols; j++) fb[0].push_back({0, j});
    for (uint32_t j = 0; j < zb.ncols; j++) fb[0].push_back({1, j});
    fb[1] = {{1, 0}, {1, 1}};
    const std::vector<F> pts = {rng() % ORC_P, rng() % ORC_P, rng() % ORC_P, rng() % ORC_P};
    const F alpha[2] = {rng() % ORC_P, rng() % ORC_P};
    vpbs::FriChain fri({&wb, &zb}, fb, pts, alpha, rate_bits);
    auto fp = fri.final_poly();
    auto wco = vpbs::PolynomialBatch::from_values(ctx, wires, rate_bits, false, cap_height);  // coefficients
    std::vector<const uint64_t*> polys;
    for (unsigned j = 0; j < ncols; j++) polys.push_back(wco.polynomials[j].data());
    for (std::size_t j = 0; j < zb.ncols; j++) polys.push_back(zref.polynomials[j].data());
    polys.push_back(zref.polynomials[0].data());
    polys.push_back(zref.polynomials[1].data());
    const uint32_t sizes[2] = {(uint32_t)(ncols + zb.ncols), 2};
    std::vector<uint64_t> want(2 * n);
    CHECK(orc_fri_final_poly(polys.data(), sizes, 2, n, pts.data(), alpha, want.data()) == 0);
    CHECK(fp == want);
    vpbs::FriChain ref(ctx, want, rate_bits);
    auto c1 = fri.commit_layer(4, 2), c2 = ref.commit_layer(4, 2);
    CHECK(std::memcmp(c1.data(), c2.data(), c1.size() * 32) == 0);
    const F beta[2] = {rng() % ORC_P, rng() % ORC_P};
    fri.fold(beta);
    ref.fold(beta);
    CHECK(fri.final_poly() == ref.final_poly());
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
