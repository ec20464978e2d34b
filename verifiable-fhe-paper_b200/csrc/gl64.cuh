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

__host__ __device__ __forceinline__ u64 canon(u64 x) { return x >= P ? x - P : x; }

// a + b, a any u64, b canonical.  Result any u64 (not necessarily canonical).
__device__ __forceinline__ u64 add_lazy(u64 a, u64 b) {
  u64 s = a + b;
  return s < a ? s + EPS : s;  // one wrap only: b < p bounds the wrapped sum below p - 1
}
// a - b, a any u64, b canonical.  Result any u64.
__device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) {
  u64 d = a - b;
  return a < b ? d - EPS : d;  // one wrap only: b < p keeps the wrapped difference >= EPS
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
__host__ __device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
#if defined(__CUDA_ARCH__) && !defined(VPBS_MUL_C)
  // Hand-scheduled: 4 x IMAD.WIDE.U32 for the 128-bit product, carry flags (not compare/select)
  // for reduce128.  lo = {p0, n0}, hi = {h0, h1}:  r = lo - h1 + h0 * (2^32 - 1)  (mod p).
  // Measured in the Poseidon leaf kernel: 7.77 ms with this sequence vs 8.61 ms with the
  // compiler's version of the C++ fallback below (-DVPBS_MUL_C).
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 p0, p1, m0, m1, n0, n1, h0, h1, t0, t1, e0, e1, bm;\n\t"
      ".reg .u64 w, z;\n\t"
      "mul.wide.u32 w, %2, %4;\n\t"          // a0*b0
      "mov.b64 {p0, p1}, w;\n\t"
      "mov.b64 z, {p1, %6};\n\t"
      "mad.wide.u32 w, %2, %5, z;\n\t"       // a0*b1 + p1
      "mov.b64 {m0, m1}, w;\n\t"
      "mov.b64 z, {m0, %6};\n\t"
      "mad.wide.u32 w, %3, %4, z;\n\t"       // a1*b0 + m0
      "mov.b64 {n0, n1}, w;\n\t"
      "mov.b64 z, {m1, %6};\n\t"
      "mad.wide.u32 w, %3, %5, z;\n\t"       // a1*b1 + m1
      "mov.b64 {h0, h1}, w;\n\t"
      "add.cc.u32 h0, h0, n1;\n\t"
      "addc.u32 h1, h1, 0;\n\t"
      "sub.cc.u32 t0, p0, h1;\n\t"           // t = lo - h1
      "subc.cc.u32 t1, n0, 0;\n\t"
      "subc.u32 bm, 0, 0;\n\t"               // 0xffffffff on borrow
      "sub.cc.u32 t0, t0, bm;\n\t"           // borrow: t -= 2^32 - 1 (i.e. += p)
      "subc.u32 t1, t1, 0;\n\t"
      "sub.cc.u32 e0, 0, h0;\n\t"            // h0 * (2^32 - 1) = {-h0, h0 - (h0 != 0)}: two ALU ops
      "subc.u32 e1, h0, 0;\n\t"              // instead of a half-rate IMAD.WIDE on the busiest pipe
      "add.cc.u32 t0, t0, e0;\n\t"
      "addc.cc.u32 t1, t1, e1;\n\t"
      "addc.u32 bm, 0, 0;\n\t"               // carry (0/1); NB: subc after add.cc sees CF inverted
      "neg.s32 bm, bm;\n\t"                  // 0xffffffff on carry
      "add.cc.u32 %0, t0, bm;\n\t"
      "addc.u32 %1, t1, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(0u));
  return ((u64)r1 << 32) | r0;
#else
  u64 lo, hi;
  mul_wide(a, b, lo, hi);
  return reduce128(lo, hi);
#endif
}
__host__ __device__ __forceinline__ u64 sqr_lazy(u64 a) { return mul_lazy(a, a); }
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
