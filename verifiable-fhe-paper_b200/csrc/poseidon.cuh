// poseidon.cuh — width-12 Poseidon permutation over Goldilocks, state held in registers.
//
// Replaces [P2] plonky2 0.2.0 src/hash/poseidon.rs (Poseidon::poseidon: 4 full + 22 partial + 4
// full rounds, S-box x^7) with the constants of src/hash/poseidon_goldilocks.rs
// (MDS_MATRIX_CIRC / MDS_MATRIX_DIAG / ALL_ROUND_CONSTANTS).  The reference selects it through
// `C = PoseidonGoldilocksConfig` (/root/reference/src/main.rs:34) and calls the same primitive
// natively at /root/reference/src/vtfhe/ivc_based_vpbs.rs:73.
//
// Design (one permutation per thread):
//  * the 12 state words live in 24 32-bit registers; the MDS layer works on the 32-bit halves
//    directly: out_r = sum_i c_i * lo(s_{i+r}) + 2^32 * sum_i c_i * hi(s_{i+r}), each sum an
//    IMAD.WIDE.U32 chain (c_i <= 41, so twelve terms stay below 2^42), recombined with one
//    96-bit reduction.  No 64x64 multiply is spent on the linear layer.
//  * the next round's constants are the initial value of those accumulators, so the constant
//    layer costs no extra instructions.
//  * S-box x^7 = (x^2 * x) * (x^2)^2: two squarings (3 wide products) and two multiplies (4).
//  * state words are kept as arbitrary u64 (lazy reduction); canonicalise once on output.
#pragma once
#include "gl64.cuh"

namespace poseidon {

using gl::u32;
using gl::u64;

constexpr int WIDTH = 12;
constexpr int RATE = 8;
constexpr int FULL_ROUNDS_HALF = 4;
constexpr int PARTIAL_ROUNDS = 22;
constexpr int ROUNDS = 2 * FULL_ROUNDS_HALF + PARTIAL_ROUNDS;

// RC[12*r + i] for r < 30, followed by one all-zero row (the "next round" of the last round).
__constant__ u64 RC[(ROUNDS + 1) * WIDTH] = {
#include "poseidon_rc.inc"
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

__device__ __forceinline__ u64 sbox7(u64 x) {
  u64 x2 = gl::sqr_lazy(x);
  u64 x4 = gl::sqr_lazy(x2);
  u64 x3 = gl::mul_lazy(x, x2);
  return gl::mul_lazy(x3, x4);
}

// acc + x * c as one IMAD.WIDE.U32 (opaque to the optimiser, which otherwise rewrites the small
// constant multiplies into shift/add chains that cost more issue slots).
template <u32 C>
__device__ __forceinline__ u64 mac32(u64 acc, u32 x) {
  u64 r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "n"(C), "l"(acc));
  return r;
}

// value = lo + 2^32 * hi with lo, hi < 2^43  ->  arbitrary-u64 representative mod p.
//   hi = hh * 2^32 + hl:  value = lo + hh * (2^32 - 1) + hl * 2^32   (2^64 = 2^32 - 1 mod p)
__device__ __forceinline__ u64 reduce96(u64 lo, u64 hi) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 hl, hh, t0, t1, bm;\n\t"
      ".reg .u64 t;\n\t"
      "mov.b64 {hl, hh}, %3;\n\t"
      "mad.wide.u32 t, hh, 0xffffffff, %2;\n\t"  // < 2^44: no carry
      "mov.b64 {t0, t1}, t;\n\t"
      "add.cc.u32 t1, t1, hl;\n\t"
      "addc.u32 bm, 0, 0;\n\t"                   // carry (0/1)
      "neg.s32 bm, bm;\n\t"                      // 0xffffffff on carry
      "add.cc.u32 %0, t0, bm;\n\t"               // carry: += 2^32 - 1
      "addc.u32 %1, t1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(lo), "l"(hi));
  return ((u64)r1 << 32) | r0;
}

template <int R, int I>
struct MdsRow {
  static constexpr u32 CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  __device__ __forceinline__ static void run(const u32 (&lo)[WIDTH], const u32 (&hi)[WIDTH],
                                             u64& acc_lo, u64& acc_hi) {
    acc_lo = mac32<CIRC[I]>(acc_lo, lo[(I + R) % WIDTH]);
    acc_hi = mac32<CIRC[I]>(acc_hi, hi[(I + R) % WIDTH]);
    if constexpr (I + 1 < WIDTH) MdsRow<R, I + 1>::run(lo, hi, acc_lo, acc_hi);
  }
};

template <int R>
__device__ __forceinline__ void mds_rows(u64 (&s)[WIDTH], const u32 (&lo)[WIDTH],
                                         const u32 (&hi)[WIDTH], const u64* __restrict__ rc_next) {
  const u64 c = rc_next[R];
  u64 acc_lo = (u32)c, acc_hi = c >> 32;
  MdsRow<R, 0>::run(lo, hi, acc_lo, acc_hi);
  if constexpr (R == 0) {  // MDS_MATRIX_DIAG = [8, 0, ..., 0]
    acc_lo = mac32<8>(acc_lo, lo[0]);
    acc_hi = mac32<8>(acc_hi, hi[0]);
  }
  s[R] = reduce96(acc_lo, acc_hi);
  if constexpr (R + 1 < WIDTH) mds_rows<R + 1>(s, lo, hi, rc_next);
}

// MDS layer fused with the following constant layer:
//   s'_r = RC_next[r] + sum_i CIRC[i] * s_{(i+r) mod 12} + DIAG[r] * s_r
__device__ __forceinline__ void mds_add_rc(u64 (&s)[WIDTH], const u64* __restrict__ rc_next) {
  u32 lo[WIDTH], hi[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; i++) {
    lo[i] = (u32)s[i];
    hi[i] = (u32)(s[i] >> 32);
  }
  mds_rows<0>(s, lo, hi, rc_next);
}

// In-place permutation; input words arbitrary u64, output words arbitrary u64 (lazy).
__device__ __forceinline__ void permute_lazy(u64 (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < WIDTH; i++) s[i] = gl::add_lazy(s[i], RC[i]);
  int r = 0;
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int k = 0; k < FULL_ROUNDS_HALF; k++, r++) {
#pragma unroll
      for (int i = 0; i < WIDTH; i++) s[i] = sbox7(s[i]);
      mds_add_rc(s, RC + (r + 1) * WIDTH);
    }
    if (half == 0) {
#pragma unroll 1
      for (int k = 0; k < PARTIAL_ROUNDS; k++, r++) {
        s[0] = sbox7(s[0]);
        mds_add_rc(s, RC + (r + 1) * WIDTH);
      }
    }
  }
}

}  // namespace poseidon
