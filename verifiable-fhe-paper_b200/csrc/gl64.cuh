// gl64.cuh — Goldilocks field (p = 2^64 - 2^32 + 1) on 32-bit integer lanes, sm_100a.
//
// Replaces [P2] plonky2_field 0.2.0 src/goldilocks_field.rs (GoldilocksField add/sub/mul/
// reduce128), which the reference reaches through `F = GoldilocksField`
// (/root/reference/src/main.rs:33-35).  Values are plain u64; like upstream, inputs may be
// non-canonical (any u64).  Functions say which operands must be canonical (< p).
#pragma once
#include <cstdint>

namespace gl {

typedef uint64_t u64;
typedef unsigned int u32;

constexpr u64 P = 0xFFFFFFFF00000001ULL;
constexpr u64 EPS = 0xFFFFFFFFULL;  // 2^64 mod p

// x mod p for any u64 x.  x >= p exactly when x + (2^32 - 1) carries out of 64 bits, and then the
// wrapped sum is x - p: on the device that is IADD3 + IADD3.X + 2 SEL on the carry predicate (the
// compare-based C form costs 6 instructions).
__host__ __device__ __forceinline__ u64 canon(u64 x) {
#if defined(__CUDA_ARCH__)
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 x0, x1, t0, t1, c;\n\t"
      ".reg .pred q;\n\t"
      "mov.b64 {x0, x1}, %2;\n\t"
      "add.cc.u32 t0, x0, 0xffffffff;\n\t"
      "addc.cc.u32 t1, x1, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "setp.ne.u32 q, c, 0;\n\t"
      "selp.u32 %0, t0, x0, q;\n\t"
      "selp.u32 %1, t1, x1, q;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(x));
  return ((u64)r1 << 32) | r0;
#else
  return x >= P ? x - P : x;
#endif
}

// a + b, a any u64, b canonical.  Result any u64 (not necessarily canonical).
// One wrap only: b < p bounds the wrapped sum below p - 1, so adding 2^64 mod p = 2^32 - 1 fits.
__device__ __forceinline__ u64 add_lazy(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 a0, a1, b0, b1, l, h, c, h2;\n\t"
      "mov.b64 {a0, a1}, %2;\n\t"
      "mov.b64 {b0, b1}, %3;\n\t"
      "add.cc.u32 l, a0, b0;\n\t"
      "addc.cc.u32 h, a1, b1;\n\t"
      "addc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, l, c;\n\t"    // + c (2^32 - 1): low word - c, high word + c - borrow
      "subc.u32 h2, h, 0;\n\t"
      "add.u32 %1, h2, c;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(a), "l"(b));
  return ((u64)r1 << 32) | r0;
}
// The same with the carry materialised on the FMA pipe (madc.lo -> IMAD.X): for kernels that are bound
// by the ALU pipe (the NTT passes: 24 ALU-pipe against 6 FMA-pipe instructions per butterfly).
__device__ __forceinline__ u64 add_lazy_fma(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 a0, a1, b0, b1, l, h, c, h2;\n\t"
      "mov.b64 {a0, a1}, %2;\n\t"
      "mov.b64 {b0, b1}, %3;\n\t"
      "add.cc.u32 l, a0, b0;\n\t"
      "addc.cc.u32 h, a1, b1;\n\t"
      "madc.lo.u32 c, 0, 0, 0;\n\t"
      "sub.cc.u32 %0, l, c;\n\t"
      "subc.u32 h2, h, 0;\n\t"
      "mad.lo.u32 %1, c, 1, h2;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(a), "l"(b));
  return ((u64)r1 << 32) | r0;
}
// a - b, a any u64, b canonical.  Result any u64.
// One wrap only: b < p keeps the wrapped difference >= 2^32 - 1.
__device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 a0, a1, b0, b1, l, h, bm;\n\t"
      "mov.b64 {a0, a1}, %2;\n\t"
      "mov.b64 {b0, b1}, %3;\n\t"
      "sub.cc.u32 l, a0, b0;\n\t"
      "subc.cc.u32 h, a1, b1;\n\t"
      "subc.u32 bm, 0, 0;\n\t"      // 0xffffffff on borrow
      "sub.cc.u32 %0, l, bm;\n\t"   // borrow: -= 2^32 - 1
      "subc.u32 %1, h, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(a), "l"(b));
  return ((u64)r1 << 32) | r0;
}
// a + b, both canonical, canonical result.
__host__ __device__ __forceinline__ u64 add(u64 a, u64 b) {
  u64 s = a + b;
  return (s < a || s >= P) ? s - P : s;
}
// a - b, both canonical, canonical result.
__host__ __device__ __forceinline__ u64 sub(u64 a, u64 b) { return a >= b ? a - b : a - b + P; }
__host__ __device__ __forceinline__ u64 neg(u64 a) { return a ? P - a : 0; }

// x = lo + 2^64*hi  ->  x mod p as an arbitrary u64 (upstream reduce128: 2^64 = EPS, 2^96 = -1).
__host__ __device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
  u32 hi_hi = (u32)(hi >> 32), hi_lo = (u32)hi;
  u64 t0 = lo - hi_hi;
  if (lo < (u64)hi_hi) t0 -= EPS;
  u64 t1 = (u64)hi_lo * (u32)EPS;
  u64 r = t0 + t1;
  return r < t1 ? r + EPS : r;
}

// 64x64 -> 128 from four 32x32->64 products (IMAD.WIDE.U32), no carry flags needed:
// every partial sum below fits in 64 bits.
__host__ __device__ __forceinline__ void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u64 p00 = (u64)a0 * b0;
  u64 mid = (u64)a0 * b1 + (p00 >> 32);
  u64 mid2 = (u64)a1 * b0 + (u32)mid;
  hi = (u64)a1 * b1 + (mid >> 32) + (mid2 >> 32);
  lo = (mid2 << 32) | (u32)p00;
}
__host__ __device__ __forceinline__ void sqr_wide(u64 a, u64& lo, u64& hi) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32);
  u64 p00 = (u64)a0 * a0;
  u64 p01 = (u64)a0 * a1;
  u64 mid = p01 + (p00 >> 32);
  u64 mid2 = p01 + (u32)mid;
  hi = (u64)a1 * a1 + (mid >> 32) + (mid2 >> 32);
  lo = (mid2 << 32) | (u32)p00;
}
// Products of arbitrary u64 operands; result arbitrary u64.
//
// Device sequence (16 SASS instructions; the first version, a chain of mad.wide with carry fix-ups
// after each reduce128 step, needed 26, the second — four mul.wide.u32 summed carry-save in PTX — 18):
//   * the 128-bit product is written as `unsigned __int128` arithmetic: ptxas then uses the 64-bit
//     addend, the carry-out predicate and the carry-in (.X) form of IMAD.WIDE.U32, which PTX cannot
//     express — 4 IMAD.WIDE.U32 + IMAD.X + IMAD.MOV + IADD3 = 7 instructions for (p0, s1, u, h1),
//     only one of them on the ALU pipe (the PTX carry-save form: 4 + 6, all six on the ALU pipe,
//     the pipe that bounds Poseidon's full rounds and the NTT butterflies);
//   * product = p0 + s1 2^32 + u 2^64 + h1 2^96 == p0 + s1 2^32 + u (2^32 - 1) - h1     (mod p)
//             = p0 - (u + h1) + (s1 + u) 2^32
//     (v, cv) = s1 + u; the carry is worth 2^64 == 2^32 - 1: v' = v + cv (cannot wrap: cv = 1 means
//               v <= 2^32 - 2) and w = u + h1 + cv (33 bits: w, cw)
//     r = (p0, v') - (w, cw); on borrow the wrapped value is >= 2^64 - 2^33, and r - (2^32 - 1)
//     is the representative.  One carry flag feeds two consumers (addc without .cc leaves CC.CF
//     alone); 9 SASS instructions.
__device__ __forceinline__ u64 reduce_words(u64 lo, u64 hi) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 p0, s1, u, h1, v, vv, w, cw, lo, hi, bm;\n\t"
      "mov.b64 {p0, s1}, %2;\n\t"
      "mov.b64 {u, h1}, %3;\n\t"
      "add.cc.u32 v, s1, u;\n\t"
      "addc.u32 vv, v, 0;\n\t"       // v' = v + cv
      "addc.cc.u32 w, u, h1;\n\t"    // w = u + h1 + cv (same flag)
      "addc.u32 cw, 0, 0;\n\t"
      "sub.cc.u32 lo, p0, w;\n\t"
      "subc.cc.u32 hi, vv, cw;\n\t"
      "subc.u32 bm, 0, 0;\n\t"       // 0xffffffff on borrow
      "sub.cc.u32 %0, lo, bm;\n\t"   // borrow: r -= 2^32 - 1
      "subc.u32 %1, hi, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(lo), "l"(hi));
  return ((u64)r1 << 32) | r0;
}
// reduce_words with its two carry-only adds on the FMA pipe (see add_lazy_fma).
__device__ __forceinline__ u64 reduce_words_fma(u64 lo, u64 hi) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 p0, s1, u, h1, v, vv, w, cw, lo, hi, bm;\n\t"
      "mov.b64 {p0, s1}, %2;\n\t"
      "mov.b64 {u, h1}, %3;\n\t"
      "add.cc.u32 v, s1, u;\n\t"
      "madc.lo.u32 vv, v, 1, 0;\n\t"
      "addc.cc.u32 w, u, h1;\n\t"
      "madc.lo.u32 cw, 0, 0, 0;\n\t"
      "sub.cc.u32 lo, p0, w;\n\t"
      "subc.cc.u32 hi, vv, cw;\n\t"
      "subc.u32 bm, 0, 0;\n\t"
      "sub.cc.u32 %0, lo, bm;\n\t"
      "subc.u32 %1, hi, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "l"(lo), "l"(hi));
  return ((u64)r1 << 32) | r0;
}
__device__ __forceinline__ u64 mul_lazy_fma(u64 a, u64 b) {
  const unsigned __int128 r = (unsigned __int128)a * b;
  return reduce_words_fma((u64)r, (u64)(r >> 64));
}
__host__ __device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  const unsigned __int128 r = (unsigned __int128)a * b;
  return reduce_words((u64)r, (u64)(r >> 64));
#else
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi);
#endif
}
// The 128-bit product only, as four 32-bit words (p0, s1, u, h1), without the reduction: for
// consumers that are linear in the result (Poseidon's MDS layer takes
// p0 - u - h1 + (s1 + u) 2^32 == a b (mod p) apart on the FP64 pipe, see poseidon.cuh).
struct Words128 {
  u32 p0, s1, u, h1;
};
__device__ __forceinline__ Words128 mul_words(u64 a, u64 b) {
  const unsigned __int128 r = (unsigned __int128)a * b;
  const u64 lo = (u64)r, hi = (u64)(r >> 64);
  return Words128{(u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)};
}
// a^2 from THREE 32x32 products.  The compiler's 128-bit a * a forms the cross product a0 a1 twice
// (4 IMAD.WIDE, an instruction that occupies the FMA-heavy pipe for four cycles and does not overlap
// with anything else on this part: tools/pipe_overlap.cu; a timing proxy that swapped one IMAD.WIDE per
// squaring for a shift measured -2.5 %).  Here it is formed once and doubled with three funnel
// shifts; the products are written as mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into
// IMAD.WIDE.U32 with a 64-bit addend and carry-out / carry-in (.X):
//     IMAD.WIDE c = a0 a1;  SHF x3 -> (d0, d1, d2) = 2 c;
//     IMAD.WIDE (w0, w1), P0 = a0 a0 + (0, d0);  IMAD.WIDE.X (w2, w3) = a1 a1 + (d1, d2) + P0
// 3 IMAD.WIDE + 3 SHF (+ one zero) instead of 4 IMAD.WIDE + 2 adds.  Leaf hashing 5.15 -> 5.06 ms.
__device__ __forceinline__ u64 sqr_lazy_dev(u64 a) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32);
  u32 w0, w1, w2, w3;
  asm("{\n\t"
      ".reg .u32 c0, c1, d0, d1, d2;\n\t"
      "mul.lo.u32 c0, %4, %5;\n\t"
      "mul.hi.u32 c1, %4, %5;\n\t"
      "shl.b32 d0, c0, 1;\n\t"
      "shf.l.wrap.b32 d1, c0, c1, 1;\n\t"
      "shr.u32 d2, c1, 31;\n\t"
      "mad.lo.cc.u32 %0, %4, %4, 0;\n\t"
      "madc.hi.cc.u32 %1, %4, %4, d0;\n\t"
      "madc.lo.cc.u32 %2, %5, %5, d1;\n\t"
      "madc.hi.u32 %3, %5, %5, d2;\n\t"
      "}"
      : "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(w3)
      : "r"(a0), "r"(a1));
  return reduce_words(((u64)w1 << 32) | w0, ((u64)w3 << 32) | w2);
}
__host__ __device__ __forceinline__ u64 sqr_lazy(u64 a) {
#if defined(__CUDA_ARCH__)
  return sqr_lazy_dev(a);
#else
  return mul_lazy(a, a);
#endif
}
__host__ __device__ __forceinline__ u64 mul(u64 a, u64 b) { return canon(mul_lazy(a, b)); }

__host__ __device__ inline u64 pow(u64 a, u64 e) {
  u64 r = 1, b = canon(a);
  while (e) {
    if (e & 1) r = mul(r, b);
    b = mul(b, b);
    e >>= 1;
  }
  return r;
}
__host__ __device__ inline u64 inv(u64 a) { return pow(a, P - 2); }

// [P2] GoldilocksField::POWER_OF_TWO_GENERATOR (= 7^((p-1)/2^32)), MULTIPLICATIVE_GROUP_GENERATOR
// = coset_shift() = 7, TWO_ADICITY = 32.
constexpr u64 POWER_OF_TWO_GENERATOR = 1753635133440165772ULL;
constexpr u64 COSET_SHIFT = 7ULL;
__host__ __device__ inline u64 primitive_root_of_unity(unsigned n_log) {
  u64 b = POWER_OF_TWO_GENERATOR;
  for (unsigned i = n_log; i < 32; i++) b = mul(b, b);
  return b;
}

}  // namespace gl
