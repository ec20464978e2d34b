// permutation.cuh — Z and partial-product polynomials of PLONK's permutation argument (sm_100a).
//
// Replaces, for F = GoldilocksField:
//   [P2] plonky2 0.2.0 src/plonk/prover.rs        wires_permutation_partial_products_and_zs
//   [P2] plonky2 0.2.0 src/util/partial_products.rs quotient_chunk_products, partial_products_and_z_gx
// i.e. step 4 of prove() ("compute partial products"), reached from
// /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364.  It is the first consumer of a
// committed batch that runs on the device: the routed wire values are
// recovered from the wires batch's coefficients in HBM and the 2 + 18 resulting columns go straight
// into the next commit without touching the host.
//
// Per row i (x_i = w_n^i), challenge (beta, gamma), routed wire j:
//     q_j = (wire_j + beta k_j x_i + gamma) / (wire_j + beta sigma_j(x_i) + gamma)
// chunk products c_k = prod of q_j over chunks of max_degree wires (K chunks), then
//     Z(x_0) = 1,   pp_k(x_i) = Z(x_i) c_0 .. c_k  (k < K - 1),   Z(x_{i+1}) = Z(x_i) c_0 .. c_{K-1}.
// Field inverses are unique, so dividing the chunk's numerator product by its denominator product
// (one Montgomery batch inversion per row) equals upstream's batch_multiplicative_inverse followed
// by per-wire multiplications bit for bit; the running product over the rows is a parallel prefix
// product (exact arithmetic: any association gives the same values).
#pragma once
#include "ntt.cuh"

namespace perm {

using gl::u32;
using gl::u64;

constexpr int MAX_CHUNKS = 32;  // K = ceil(num_routed / max_degree); plonky2's standard config: 10

// a^(p-2) for canonical a != 0.  p - 2 = 2^64 - 2^32 - 1 = (2^32 - 1) * 2^32 + (2^32 - 1) - ... is
// evaluated with the chain x^(2^32 - 1) -> shift by 32 -> times x^(2^32 - 2) ... ; 63 squarings + 8
// multiplies in the classic form for this prime:
//   e = p - 2 = 0xFFFFFFFE_FFFFFFFF:  bits 63..33 ones, bit 32 zero, bits 31..0 ones.
__device__ __forceinline__ u64 inv_nonzero(u64 a) {
  // t_k = a^(2^k - 1)
  u64 t1 = a;
  u64 t2 = gl::mul_lazy(gl::sqr_lazy(t1), t1);                              // 2^2 - 1
  u64 t4 = t2;
  for (int i = 0; i < 2; i++) t4 = gl::sqr_lazy(t4);
  t4 = gl::mul_lazy(t4, t2);                                                // 2^4 - 1
  u64 t8 = t4;
  for (int i = 0; i < 4; i++) t8 = gl::sqr_lazy(t8);
  t8 = gl::mul_lazy(t8, t4);                                                // 2^8 - 1
  u64 t16 = t8;
  for (int i = 0; i < 8; i++) t16 = gl::sqr_lazy(t16);
  t16 = gl::mul_lazy(t16, t8);                                              // 2^16 - 1
  u64 t31 = t16;
  for (int i = 0; i < 8; i++) t31 = gl::sqr_lazy(t31);
  t31 = gl::mul_lazy(t31, t8);                                              // 2^24 - 1
  for (int i = 0; i < 4; i++) t31 = gl::sqr_lazy(t31);
  t31 = gl::mul_lazy(t31, t4);                                              // 2^28 - 1
  for (int i = 0; i < 2; i++) t31 = gl::sqr_lazy(t31);
  t31 = gl::mul_lazy(t31, t2);                                              // 2^30 - 1
  t31 = gl::mul_lazy(gl::sqr_lazy(t31), t1);                                // 2^31 - 1
  u64 t32 = gl::mul_lazy(gl::sqr_lazy(t31), t1);                            // 2^32 - 1
  // exponent = (2^31 - 1) * 2^33 + (2^32 - 1)
  u64 r = t31;
  for (int i = 0; i < 33; i++) r = gl::sqr_lazy(r);
  return gl::canon(gl::mul_lazy(r, t32));
}

// One thread per row.  wires / sigmas: column-major (column j at + j * col_stride), natural order.
// quot[k * n + i] = chunk product c_k of row i; rowprod[i] = c_0 .. c_{K-1}.
// *zero_flag is set if some chunk's denominator product is zero (upstream panics there).
__global__ void __launch_bounds__(128)
chunk_quotients(const u64* __restrict__ wires, u64 wires_stride, const u64* __restrict__ sigmas,
                u64 sigmas_stride, const u64* __restrict__ k_is, u32 num_routed, u32 max_degree,
                unsigned log_n, u64 beta, u64 gamma, ntt::Roots R, u64* __restrict__ quot,
                u64* __restrict__ rowprod, int* __restrict__ zero_flag) {
  const u64 n = 1ULL << log_n;
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 x = log_n ? ntt::root_of<false>(R, log_n, i) : 1;
  const u64 bx = gl::mul(beta, x);  // beta * x_i
  const u32 K = (num_routed + max_degree - 1) / max_degree;
  u64 num[MAX_CHUNKS], den[MAX_CHUNKS];
#pragma unroll 1
  for (u32 k = 0; k < K; k++) {
    u64 pn = 1, pd = 1;
    const u32 j1 = (k + 1) * max_degree < num_routed ? (k + 1) * max_degree : num_routed;
#pragma unroll 1
    for (u32 j = k * max_degree; j < j1; j++) {
      const u64 w = gl::canon(__ldg(wires + (u64)j * wires_stride + i));
      const u64 s = __ldg(sigmas + (u64)j * sigmas_stride + i);
      const u64 wg = gl::add(w, gamma);
      const u64 a = gl::add(wg, gl::mul(__ldg(k_is + j), bx));   // wire + beta k_j x + gamma
      const u64 b = gl::add(wg, gl::mul(beta, s));               // wire + beta sigma_j + gamma
      pn = gl::mul_lazy(pn, a);
      pd = gl::mul_lazy(pd, b);
    }
    num[k] = gl::canon(pn);
    den[k] = gl::canon(pd);
  }
  // Montgomery batch inversion of den[0..K): prefix products, one inversion, walk back.
  u64 pre[MAX_CHUNKS];
  u64 acc = 1;
#pragma unroll 1
  for (u32 k = 0; k < K; k++) {
    pre[k] = acc;
    acc = gl::mul(acc, den[k]);
  }
  if (acc == 0) {
    atomicOr(zero_flag, 1);
    return;
  }
  u64 inv = inv_nonzero(acc);
  u64 row = 1;
#pragma unroll 1
  for (u32 k = K; k-- > 0;) {
    const u64 dinv = gl::mul(inv, pre[k]);  // 1 / den[k]
    inv = gl::mul(inv, den[k]);
    num[k] = gl::mul(num[k], dinv);
  }
#pragma unroll 1
  for (u32 k = 0; k < K; k++) {
    quot[(u64)k * n + i] = num[k];
    row = gl::mul(row, num[k]);
  }
  rowprod[i] = row;
}

// ---- inclusive prefix product (Goldilocks), 256 elements per CTA ----------------------------------
constexpr int SCAN_THREADS = 256;
__device__ __forceinline__ u64 shfl_up64(u64 v, unsigned d) {
  const u32 lo = __shfl_up_sync(0xffffffffu, (u32)v, d), hi = __shfl_up_sync(0xffffffffu, (u32)(v >> 32), d);
  return ((u64)hi << 32) | lo;
}
// data[i] <- data[first of its 256-block] * .. * data[i]; totals[block] <- product of the block
// (elements beyond count behave as 1).
__global__ void __launch_bounds__(SCAN_THREADS)
scan_blocks(u64* __restrict__ data, u64 count, u64* __restrict__ totals) {
  __shared__ u64 warp_tot[SCAN_THREADS / 32];
  const u64 i = (u64)blockIdx.x * SCAN_THREADS + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u64 v = i < count ? data[i] : 1;
#pragma unroll
  for (unsigned d = 1; d < 32; d <<= 1) {
    const u64 o = shfl_up64(v, d);
    if (lane >= d) v = gl::mul(v, o);
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    u64 t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 1;
#pragma unroll
    for (unsigned d = 1; d < SCAN_THREADS / 32; d <<= 1) {
      const u64 o = shfl_up64(t, d);
      if (lane >= d) t = gl::mul(t, o);
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = t;
  }
  __syncthreads();
  if (warp > 0) v = gl::mul(v, warp_tot[warp - 1]);
  if (i < count) data[i] = v;
  if (totals && threadIdx.x == SCAN_THREADS - 1) totals[blockIdx.x] = v;
}
// data[i] *= scanned_totals[block - 1] for every block but the first.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply(u64* __restrict__ data, u64 count, const u64* __restrict__ scanned_totals) {
  const u64 i = (u64)blockIdx.x * SCAN_THREADS + threadIdx.x;
  if (blockIdx.x == 0 || i >= count) return;
  data[i] = gl::mul(data[i], scanned_totals[blockIdx.x - 1]);
}

// out column 0 (Z) and columns 1 .. K-1 (partial products) of one challenge; rowscan = inclusive
// prefix product of the row products.  z_col / pp_col0: destination columns (stride n).
__global__ void __launch_bounds__(128)
finish_rows(const u64* __restrict__ quot, const u64* __restrict__ rowscan, unsigned log_n, u32 K,
            u64* __restrict__ z_col, u64* __restrict__ pp_col0) {
  const u64 n = 1ULL << log_n;
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 acc = i ? rowscan[i - 1] : 1;  // Z(x_i)
  z_col[i] = acc;
  for (u32 k = 0; k + 1 < K; k++) {
    acc = gl::mul(acc, quot[(u64)k * n + i]);
    pp_col0[(u64)k * n + i] = acc;
  }
}

// rows_out[r][c] = leaves[bitrev(first + r * step)][c] for c < ncols: PolynomialBatch::
// get_lde_values(first + r * step, 1) for a block of LDE indices, salt columns dropped.
__global__ void __launch_bounds__(128)
gather_lde_rows(const u64* __restrict__ leaves, u32 width, u32 ncols, unsigned log_m, u64 first,
                u64 step, u64 count, u64 leaf_off, u64* __restrict__ rows_out) {
  const u64 r = blockIdx.x;
  if (r >= count) return;
  const u64 idx = first + r * step;
  const u64 leaf = (log_m ? (__brevll(idx) >> (64 - log_m)) : 0) - leaf_off;  // leaf_off: first leaf of a shard
  const u64* src = leaves + leaf * width;
  for (u32 c = threadIdx.x; c < ncols; c += blockDim.x) rows_out[r * ncols + c] = src[c];
}

// ---- quotient polynomial: the gate-independent vanishing terms -------------------------------------------
// [P2] plonk/prover.rs compute_quotient_polys + plonk/vanishing_poly.rs eval_vanishing_poly_base_batch
// (step 6 of prove(), reached from /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364): on
// the quotient domain x_i = 7 w_q^i, q = n << qdb, per challenge c
//   Z(1) = 1 term      L_0(x) (Z_c(x) - 1),  L_0(x) = (x^n - 1) / (n (x - 1))
//   partial products   accs = [Z_c(x), pp_c,0 .. pp_c,K-2, Z_c(g x)];
//                      check_t = accs[t] prod_chunk (w_j + beta_c k_j x + gamma_c)
//                              - accs[t+1] prod_chunk (w_j + beta_c sigma_j(x) + gamma_c)
//   res_c = sum_j term_j alpha_c^j over [Z(1) terms of all challenges, checks of challenge 0, 1, ..]
//           (+ alpha_c^(nc + nc K) times the alpha-reduced gate constraints, if the caller supplies them)
//   value_c(i) = res_c / (x^n - 1).
// The LDE rows come straight from the resident batches: the points of the quotient domain are the
// first q leaves of the (bit-reversed) leaf matrices, leaf k <-> natural index bitrev_q(k), so thread
// k reads row k of each matrix (a warp reads 32 consecutive rows) and the "next" row i + 2^qdb.
struct QuotientParams {
  u64 zh[32], zh_inv[32];           // x^n - 1 on the coset takes 2^qdb values; and their inverses
  u64 beta[4], gamma[4];            // per challenge (at most 4)
  u64 apow[4][4 + 4 * MAX_CHUNKS];  // alpha_c^j for j < nc + nc K
  u64 agate[4];                     // alpha_c^(nc + nc K)
  u64 n_canon;                      // n mod p
};
__global__ void __launch_bounds__(128)
quotient_permutation_terms(const u64* __restrict__ wires, u32 wires_width, const u64* __restrict__ cs,
                           u32 cs_width, u32 sigmas_first, const u64* __restrict__ zs, u32 zs_width,
                           const u64* __restrict__ k_is, u32 num_routed, u32 max_degree, u32 K, u32 nc,
                           unsigned log_q, unsigned qdb, const __grid_constant__ QuotientParams qp,
                           ntt::Roots R, const u64* __restrict__ gate_terms, u64 k0, u64 kcount,
                           u64* __restrict__ vals) {
  // leaves [k0, k0 + kcount) of the quotient domain: all of it, or the shard whose rows the batches hold
  // (row r of the matrices is leaf k0 + r; the "next" point lies in the same LDE block, hence the same shard)
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 q = 1ULL << log_q;
  if (t >= kcount) return;
  const u64 k = k0 + t;
  const u64 i = log_q ? (__brevll(k) >> (64 - log_q)) : 0;
  const u64 i_next = (i + (1ULL << qdb)) & (q - 1);
  const u64 k_next = log_q ? (__brevll(i_next) >> (64 - log_q)) : 0;
  const u64 x = gl::mul(gl::COSET_SHIFT, ntt::root_of<false>(R, log_q, i));
  const u64* wrow = wires + t * wires_width;
  const u64* srow = cs + t * cs_width + sigmas_first;
  const u64* zrow = zs + t * zs_width;
  const u64* znext = zs + (k_next - k0) * zs_width;
  const unsigned cosetk = (unsigned)(i & ((1u << qdb) - 1));
  u64 res[4] = {0, 0, 0, 0};
  // Z(1) = 1 terms
  const u64 l0 = gl::mul(qp.zh[cosetk], inv_nonzero(gl::mul(qp.n_canon, gl::sub(x, 1))));
  // (loops over the challenges are unrolled to 4 with a guard so that res / num / den stay in registers)
#pragma unroll
  for (u32 c = 0; c < 4; c++) {
    if (c < nc) {
      const u64 term = gl::mul(l0, gl::sub(__ldg(zrow + c), 1));
#pragma unroll
      for (u32 d = 0; d < 4; d++)
        if (d < nc) res[d] = gl::add(res[d], gl::mul(term, qp.apow[d][c]));
    }
  }
  // partial-product checks, chunk by chunk (the wire / sigma values of a chunk serve every challenge)
  for (u32 t = 0; t < K; t++) {
    u64 num[4] = {1, 1, 1, 1}, den[4] = {1, 1, 1, 1};
    const u32 j1 = (t + 1) * max_degree < num_routed ? (t + 1) * max_degree : num_routed;
    for (u32 j0 = t * max_degree; j0 < j1; j0 += 8) {
      // eight wires' values and sigmas are requested before any of them is used: a thread's loads hit one
      // or two cache lines of its own rows, and issued one by one between dependent multiplies they made
      // the kernel latency-bound (ncu: 10 long-scoreboard stalls per issue, 1.75x the algorithmic DRAM bytes)
      u64 wv[8], sv[8];
#pragma unroll
      for (u32 u = 0; u < 8; u++) {
        const u32 j = j0 + u < j1 ? j0 + u : j1 - 1;
        wv[u] = __ldg(wrow + j);
        sv[u] = __ldg(srow + j);
      }
#pragma unroll
      for (u32 u = 0; u < 8; u++) {
        if (j0 + u < j1) {
          const u64 kx = gl::mul(__ldg(k_is + j0 + u), x);
#pragma unroll
          for (u32 c = 0; c < 4; c++) {
            if (c < nc) {
              // wire + beta s + gamma in one 128-bit multiply-add (exact, reduced once); the running
              // products stay arbitrary u64 representatives until the chunk is complete
              num[c] = gl::mul_lazy(num[c], ntt::mul_add2_lazy(qp.beta[c], kx, wv[u], qp.gamma[c]));
              den[c] = gl::mul_lazy(den[c], ntt::mul_add2_lazy(qp.beta[c], sv[u], wv[u], qp.gamma[c]));
            }
          }
        }
      }
    }
#pragma unroll
    for (u32 c = 0; c < 4; c++) {
      if (c < nc) {
        const u64 prev = t == 0 ? __ldg(zrow + c) : __ldg(zrow + nc + c * (K - 1) + (t - 1));
        const u64 next = t == K - 1 ? __ldg(znext + c) : __ldg(zrow + nc + c * (K - 1) + t);
        const u64 term = gl::sub(gl::mul(prev, num[c]), gl::mul(next, den[c]));
        const u32 idx = nc + c * K + t;
#pragma unroll
        for (u32 d = 0; d < 4; d++)
          if (d < nc) res[d] = gl::add(res[d], gl::mul(term, qp.apow[d][idx]));
      }
    }
  }
#pragma unroll
  for (u32 c = 0; c < 4; c++) {
    if (c < nc) {
      u64 r = res[c];
      if (gate_terms) r = gl::add(r, gl::mul(qp.agate[c], gl::canon(__ldg(gate_terms + (u64)c * q + i))));
      vals[(u64)c * q + i] = gl::mul(r, qp.zh_inv[cosetk]);
    }
  }
}
// ---- gate constraints as a straight-line program -----------------------------------------------------------
// [P2] plonk/vanishing_poly.rs evaluate_gate_constraints_base_batch: constraint_j(x) = sum over the
// gates of filter_g(selectors(x)) * gate_g.eval_unfiltered_base(local_constants(x), local_wires(x),
// public_inputs_hash)_j.  The gates' constraint polynomials are plonky2 source (gates/*.rs) that this
// repository does not restate; they reach the device as DATA: a program the host compiles once per
// circuit from the gates' own evaluation code, executed here at every point of the
// quotient domain.  One 64-bit word per instruction:
//     bits 0-7 op | 8-15 dst register | 16-19 kind of a | 20-23 kind of b | 24-39 index of a | 40-55 index of b
//     op:   0 ADD  1 SUB  2 MUL   dst <- a op b
//           3 EMIT     the current gate's constraint number (index of b) has value a
//           4 ENDGATE  the gate's constraints, times the filter value a, are added to the totals
//           5 MAD      dst <- dst + a b   (linear layers: one instruction and one reduction per term)
//     kind: 0 register  1 wire column  2 column of the constants/sigmas batch (selectors and
//           constants)  3 entry of the immediate table  4 public_inputs_hash element
// Constraint j enters the total of challenge c times alpha_c^j (reduce_with_powers), so per gate the
// kernel keeps sum_j alpha_c^j c_j and multiplies it by the filter at ENDGATE: exactly
// sum_j alpha_c^j sum_g filter_g c_{g,j}.  Registers live in shared memory ([register][thread]).
#ifndef VPBS_PROG_THREADS
#define VPBS_PROG_THREADS 128
#endif
constexpr int PROG_THREADS = VPBS_PROG_THREADS;
enum : unsigned { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_EMIT = 3, OP_ENDGATE = 4, OP_MAD = 5 };
enum : unsigned { K_REG = 0, K_WIRE = 1, K_CONST = 2, K_IMM = 3, K_PIH = 4 };
// PROG_POINTS points per thread share one decode of every instruction (the decode, dispatch and loop
// bookkeeping are about 40 % of the per-instruction cost with one point per thread); a CTA covers
// PROG_THREADS * PROG_POINTS consecutive leaves, point p of thread t being leaf base + p * PROG_THREADS + t
// so that a warp still reads consecutive rows.
#ifndef VPBS_PROG_POINTS
#define VPBS_PROG_POINTS 1
#endif
constexpr int PROG_POINTS = VPBS_PROG_POINTS;
constexpr unsigned MAX_PROG_REGS = 224u * 128u / (PROG_THREADS * PROG_POINTS);  // 224 KB of shared memory
__global__ void __launch_bounds__(PROG_THREADS)
gate_program_eval(const u64* __restrict__ code, u32 ncode, const u64* __restrict__ imm,
                  const u64* __restrict__ apow /* nc x num_constraints: alpha_c^j, then public_inputs_hash[4] */,
                  u32 num_constraints,
                  const u64* __restrict__ wires, u32 wires_width, const u64* __restrict__ cs, u32 cs_width,
                  u32 nc, unsigned log_q, u64 k0, u64 kcount, u64* __restrict__ out) {
  extern __shared__ u64 regs[];  // [register][point][thread]
  constexpr int PP = PROG_POINTS;
  const u64 q = 1ULL << log_q;
  const unsigned tid = threadIdx.x;
  const u64 base = (u64)blockIdx.x * (PROG_THREADS * PP);
  bool live[PP];
  const u64 *wrow[PP], *crow[PP];
#pragma unroll
  for (int p = 0; p < PP; p++) {
    const u64 t0 = base + (u64)p * PROG_THREADS + tid;
    live[p] = t0 < kcount;
    const u64 tt = live[p] ? t0 : 0;  // row of the (possibly sharded) leaf matrices; leaf k0 + tt
    wrow[p] = wires + tt * wires_width;
    crow[p] = cs + tt * cs_width;
  }
  u64 total[PP][4], gacc[PP][4];  // fully unrolled below: registers
#pragma unroll
  for (int p = 0; p < PP; p++)
#pragma unroll
    for (int c = 0; c < 4; c++) total[p][c] = gacc[p][c] = 0;
  // register file: plain 32-bit shared-memory addresses (LDS / STS; the generic-pointer form re-derives
  // the shared window base at every access) — register r, point p of this thread lives at
  // rbase + (r * PP + p) * PROG_THREADS * 8
  const unsigned rbase = (unsigned)__cvta_generic_to_shared(regs) + tid * 8u;
  constexpr unsigned RSTRIDE = PROG_THREADS * 8u;
  auto lds = [&](unsigned r, int p) -> u64 {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(rbase + (r * PP + p) * RSTRIDE));
    return v;
  };
  // operand of a kind other than "register": every such source is a global-memory table, so the kind
  // only SELECTS a base pointer (a switch / if-chain becomes an indirect branch through a
  // constant-memory jump table, the top stall of the first version of this kernel)
  const u64* pihp = apow + (u64)nc * num_constraints;  // the caller appends public_inputs_hash there
  auto fetch_slow = [&](unsigned kind, unsigned idx, int p) -> u64 {
    const u64* ptr = kind == K_IMM ? imm : kind == K_WIRE ? wrow[p] : kind == K_CONST ? crow[p] : pihp;
    return __ldg(ptr + idx);
  };
  // the instruction stream is the same for every thread: the next word is fetched while the current
  // one executes, and register operands (the common case) skip the operand-kind dispatch
  u64 ins = ncode ? __ldg(code) : 0;
  for (u32 pc = 0; pc < ncode; pc++) {
    const u64 next = pc + 1 < ncode ? __ldg(code + pc + 1) : 0;
    const unsigned lo = (unsigned)ins, hi = (unsigned)(ins >> 32);
    const unsigned op = lo & 0xff, dst = (lo >> 8) & 0xff;
    const unsigned ka = (lo >> 16) & 0xf, kb = (lo >> 20) & 0xf;
    const unsigned ia = (lo >> 24) | ((hi & 0xff) << 8), ib = (hi >> 8) & 0xffff;
    u64 a[PP];
    if (ka == K_REG) {
#pragma unroll
      for (int p = 0; p < PP; p++) a[p] = lds(ia, p);
    } else {
#pragma unroll
      for (int p = 0; p < PP; p++) a[p] = fetch_slow(ka, ia, p);
    }
    if (op <= OP_MUL || op == OP_MAD) {
      u64 b[PP], r[PP];
      if (kb == K_REG) {
#pragma unroll
        for (int p = 0; p < PP; p++) b[p] = lds(ib, p);
      } else {
#pragma unroll
        for (int p = 0; p < PP; p++) b[p] = fetch_slow(kb, ib, p);
      }
      if (op == OP_MAD) {
#pragma unroll
        for (int p = 0; p < PP; p++) r[p] = gl::canon(ntt::mul_add2_lazy(a[p], b[p], lds(dst, p), 0));
      } else if (op == OP_MUL) {
#pragma unroll
        for (int p = 0; p < PP; p++) r[p] = gl::mul(a[p], b[p]);
      } else if (op == OP_ADD) {
#pragma unroll
        for (int p = 0; p < PP; p++) r[p] = gl::add(a[p], b[p]);
      } else {
#pragma unroll
        for (int p = 0; p < PP; p++) r[p] = gl::sub(a[p], b[p]);
      }
#pragma unroll
      for (int p = 0; p < PP; p++)
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(rbase + (dst * PP + p) * RSTRIDE), "l"(r[p]) : "memory");
    } else if (op == OP_EMIT) {
#pragma unroll
      for (u32 c = 0; c < 4; c++)
        if (c < nc) {
          const u64 w = __ldg(apow + (u64)c * num_constraints + ib);
#pragma unroll
          for (int p = 0; p < PP; p++) gacc[p][c] = gl::add(gacc[p][c], gl::mul(a[p], w));
        }
    } else {  // OP_ENDGATE
#pragma unroll
      for (u32 c = 0; c < 4; c++)
        if (c < nc) {
#pragma unroll
          for (int p = 0; p < PP; p++) {
            total[p][c] = gl::add(total[p][c], gl::mul(gacc[p][c], a[p]));
            gacc[p][c] = 0;
          }
        }
    }
    ins = next;
  }
#pragma unroll
  for (int p = 0; p < PP; p++) {
    if (live[p]) {
      const u64 k = k0 + base + (u64)p * PROG_THREADS + tid;
      const u64 i = log_q ? (__brevll(k) >> (64 - log_q)) : 0;
#pragma unroll
      for (u32 c = 0; c < 4; c++)
        if (c < nc) out[(u64)c * q + i] = total[p][c];
    }
  }
}

// coset_ifft's tail: coefficient j of every column times shift^-j
__global__ void coset_unscale(u64* __restrict__ coeffs, u64 len, u32 ncols, u64 shift_inv) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= len) return;
  u64 s = 1, b = shift_inv;
  for (u64 e = t; e; e >>= 1) {
    if (e & 1) s = gl::mul(s, b);
    b = gl::mul(b, b);
  }
  for (u32 c = 0; c < ncols; c++) coeffs[(u64)c * len + t] = gl::mul(coeffs[(u64)c * len + t], s);
}

}  // namespace perm
