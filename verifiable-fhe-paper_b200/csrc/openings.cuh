// openings.cuh — the front half of PolynomialBatch::prove_openings on the device (sm_100a).
//
// Replaces, for F = GoldilocksField and its quadratic extension F[X]/(X^2 - 7):
//   [P2] plonky2 0.2.0 src/fri/oracle.rs          PolynomialBatch::prove_openings (up to final_poly)
//   [P2] plonky2 0.2.0 src/util/reducing.rs       ReducingFactor::{reduce_polys_base, shift_poly}
//   [P2] plonky2_field 0.2.0 src/polynomial/division.rs  PolynomialCoeffs::divide_by_linear
// reached from prove() at /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364 ("reduce batch
// of N polynomials" scopes).  Per FRI batch b (opening point z_b, polynomials f_b0, f_b1, ...):
//     F_b = sum_j alpha^j f_bj                                  (base coefficients x extension powers)
//     Q_b = (F_b(X) - F_b(z_b)) / (X - z_b),  padded back to n coefficients with a zero
//     final = final * alpha^(len_b) + Q_b
// The coefficient polynomials are read where the resident batches keep them (column-major in HBM);
// only alpha and the points come from the host, nothing goes back.
//
// divide_by_linear is the Horner recurrence b_i = b_{i+1} z + c_i run from the top (q_i = b_{i+1}):
// a suffix scan of affine maps.  Here: 16 coefficients per thread, 4096 per CTA; pass 1 reduces every
// CTA's block to its value at the block start, a serial pass combines the (few) blocks, pass 2
// replays each block with its incoming value and writes the quotient (fused with the final +=).
#pragma once
#include "ntt.cuh"

namespace openings {

using gl::u32;
using gl::u64;
using ntt::Ext2;
using ntt::ext_mul;

__device__ __forceinline__ Ext2 ext_add(Ext2 a, Ext2 b) { return Ext2{gl::add(a.re, b.re), gl::add(a.im, b.im)}; }
__device__ __forceinline__ Ext2 ext_pow(Ext2 b, u64 e) {
  Ext2 r{1, 0};
  while (e) {
    if (e & 1) r = ext_mul(r, b);
    b = ext_mul(b, b);
    e >>= 1;
  }
  return r;
}

// comp[i] = sum_j powers[j] * polys[j][i]  (i < n): one thread per coefficient, polynomials streamed
// (every read is a coalesced run of one column).  powers: alpha^j as (re, im) pairs, canonical.
__global__ void __launch_bounds__(256)
reduce_polys(const u64* const* __restrict__ polys, u32 npolys, const ulonglong2* __restrict__ powers,
             u64 n, ulonglong2* __restrict__ comp) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 re = 0, im = 0;
  for (u32 j = 0; j < npolys; j++) {
    const u64 c = gl::canon(__ldg(polys[j] + i));
    const ulonglong2 pw = __ldg(powers + j);
    re = gl::add(re, gl::mul(c, pw.x));
    im = gl::add(im, gl::mul(c, pw.y));
  }
  comp[i] = make_ulonglong2(re, im);
}

constexpr int DIV_THREADS = 256, DIV_PER_THREAD = 16, DIV_BLOCK = DIV_THREADS * DIV_PER_THREAD;

// Value of one CTA's block of coefficients at its own start: U = sum_{j in block} c_j z^(j - start),
// plus (in shared memory, for the caller) the inclusive suffix values U_t of its 256 chunks.
__device__ __forceinline__ void block_suffix(const ulonglong2* __restrict__ comp, u64 n, u64 start, Ext2 z,
                                             Ext2* sh_u, Ext2 (&c)[DIV_PER_THREAD]) {
  const unsigned t = threadIdx.x;
  // thread t owns coefficients start + 16 t .. + 15 (zero beyond n)
#pragma unroll
  for (int k = 0; k < DIV_PER_THREAD; k++) {
    const u64 idx = start + (u64)t * DIV_PER_THREAD + k;
    ulonglong2 v = make_ulonglong2(0, 0);
    if (idx < n) v = comp[idx];
    c[k] = Ext2{v.x, v.y};
  }
  Ext2 acc{0, 0};
#pragma unroll
  for (int k = DIV_PER_THREAD - 1; k >= 0; k--) acc = ext_add(ext_mul(acc, z), c[k]);
  sh_u[t] = acc;  // S_t
  __syncthreads();
  // inclusive suffix scan U_t = S_t + z^16 U_{t+1} by doubling: after the step with distance d,
  // U_t covers chunks t .. t + 2d - 1
  Ext2 zp = ext_pow(z, DIV_PER_THREAD);  // z^(16 d)
  for (unsigned d = 1; d < DIV_THREADS; d <<= 1) {
    Ext2 add{0, 0};
    if (t + d < DIV_THREADS) add = ext_mul(zp, sh_u[t + d]);
    __syncthreads();
    sh_u[t] = ext_add(sh_u[t], add);
    __syncthreads();
    zp = ext_mul(zp, zp);
  }
}

// pass 1: totals[block] = the block's value at its start
__global__ void __launch_bounds__(DIV_THREADS)
divide_pass1(const ulonglong2* __restrict__ comp, u64 n, u64 z_re, u64 z_im, ulonglong2* __restrict__ totals) {
  __shared__ Ext2 sh_u[DIV_THREADS];
  Ext2 c[DIV_PER_THREAD];
  block_suffix(comp, n, (u64)blockIdx.x * DIV_BLOCK, Ext2{z_re, z_im}, sh_u, c);
  if (threadIdx.x == 0) totals[blockIdx.x] = make_ulonglong2(sh_u[0].re, sh_u[0].im);
}
// between the passes: carries[b] = b_{end of block b} = sum_{b' > b} totals[b'] z^(4096 (b' - b - 1)),
// serial over the blocks (n / 4096 of them: 16 at the N=1024 step's size)
__global__ void divide_carries(const ulonglong2* __restrict__ totals, u64 nblocks, u64 z_re, u64 z_im,
                               ulonglong2* __restrict__ carries) {
  if (blockIdx.x || threadIdx.x) return;
  const Ext2 zb = ext_pow(Ext2{z_re, z_im}, DIV_BLOCK);
  Ext2 acc{0, 0};
  for (u64 b = nblocks; b-- > 0;) {
    carries[b] = make_ulonglong2(acc.re, acc.im);
    acc = ext_add(ext_mul(acc, zb), Ext2{totals[b].x, totals[b].y});
  }
}
// pass 2: quotient coefficients q_i = b_{i+1} (q_{n-1} = 0: "pad back to power of two"), fused with
// final = final * shift + q  (ReducingFactor::shift_poly, then +=); first batch: shift = 0 and the
// old contents of `final` are ignored.
__global__ void __launch_bounds__(DIV_THREADS)
divide_pass2(const ulonglong2* __restrict__ comp, u64 n, u64 z_re, u64 z_im,
             const ulonglong2* __restrict__ carries, u64 shift_re, u64 shift_im, int first,
             ulonglong2* __restrict__ final_poly) {
  __shared__ Ext2 sh_u[DIV_THREADS];
  Ext2 c[DIV_PER_THREAD];
  const Ext2 z{z_re, z_im}, shift{shift_re, shift_im};
  const u64 start = (u64)blockIdx.x * DIV_BLOCK;
  block_suffix(comp, n, start, z, sh_u, c);
  const unsigned t = threadIdx.x;
  // b at the top of this thread's chunk: the chunks above it inside the block, then the blocks above
  Ext2 acc = t + 1 < DIV_THREADS ? sh_u[t + 1] : Ext2{0, 0};
  const ulonglong2 cb = carries[blockIdx.x];
  acc = ext_add(acc, ext_mul(ext_pow(z, (u64)DIV_PER_THREAD * (DIV_THREADS - 1 - t)), Ext2{cb.x, cb.y}));
#pragma unroll
  for (int k = DIV_PER_THREAD - 1; k >= 0; k--) {
    const u64 idx = start + (u64)t * DIV_PER_THREAD + k;
    if (idx < n) {
      Ext2 q = acc;  // b_{idx + 1}
      if (!first) {
        const ulonglong2 f = final_poly[idx];
        q = ext_add(ext_mul(Ext2{f.x, f.y}, shift), q);
      }
      final_poly[idx] = make_ulonglong2(q.re, q.im);
    }
    acc = ext_add(ext_mul(acc, z), c[k]);
  }
}

}  // namespace openings
