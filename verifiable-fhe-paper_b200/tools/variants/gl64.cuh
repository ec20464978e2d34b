// MEASURED ALTERNATIVES — not part of the product.  Round-1 form of csrc/gl64.cuh with every compile-time
// variant DESIGN.md quotes a measurement for (Goldilocks arithmetic): -DVPBS_MUL_C, -DVPBS_ADDSUB_C, -DVPBS_CANON_C,
// -DVPBS_MDS_INT32, -DVPBS_MDS_FP64_DENSE, -DVPBS_SBOX_REDUCED, -DVPBS_SBOX_OUTLINE, -DVPBS_SBOX_CALL4,
// -DVPBS_NO_PIPE_INTERLEAVE, -DVPBS_HALF_I2F.  Build a tool against them with
//   nvcc ... -I tools/variants -I csrc tools/selftest.cu   (this directory first)
// tests/test_abi_and_host.py keeps them compiling; tools/selftest.cu checks them bit for bit on a GPU.
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
#if defined(__CUDA_ARCH__) && !defined(VPBS_CANON_C)
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
#if !defined(VPBS_ADDSUB_C)
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
#else
  u64 s = a + b;
  return s < a ? s + EPS : s;
#endif
}
// a - b, a any u64, b canonical.  Result any u64.
// One wrap only: b < p keeps the wrapped difference >= 2^32 - 1.
__device__ __forceinline__ u64 sub_lazy(u64 a, u64 b) {
#if !defined(VPBS_ADDSUB_C)
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
#else
  u64 d = a - b;
  return a < b ? d - EPS : d;
#endif
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
// Device sequence (4 IMAD.WIDE.U32 + 14 integer instructions in SASS; the first version, a chain
// of mad.wide with carry fix-ups after each reduce128 step, needed 4 + 22):
//   p = a0 b0, x = a0 b1, y = a1 b0, z = a1 b1                     (32x32 -> 64 each)
//   (t0, t1, t2) = x + y                                           65-bit sum of the cross terms
//   (s1, u, h1)  = (t0, t1, t2) + (p1, z0, z1)                     product = (p0, s1, u, h1)
//   product = p0 + s1 2^32 + u 2^64 + h1 2^96 == p0 + s1 2^32 + u (2^32 - 1) - h1     (mod p)
//           = p0 - (u + h1) + (s1 + u) 2^32
//   (v, cv) = s1 + u; the carry is worth 2^64 == 2^32 - 1: v' = v + cv (cannot wrap: cv = 1 means
//             v <= 2^32 - 2) and w = u + h1 + cv (33 bits: w, cw)
//   r = (p0, v') - (w, cw); on borrow the wrapped value is >= 2^64 - 2^33, and r - (2^32 - 1)
//   is the representative.
// One carry flag feeds two consumers below (addc without .cc leaves CC.CF alone).
__host__ __device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
#if defined(__CUDA_ARCH__) && !defined(VPBS_MUL_C)
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 p0, p1, x0, x1, y0, y1, z0, z1, t0, t1, t2, s1, u, h1, v, vv, w, cw, lo, hi, bm;\n\t"
      ".reg .u64 q;\n\t"
      "mul.wide.u32 q, %2, %4;\n\t"
      "mov.b64 {p0, p1}, q;\n\t"
      "mul.wide.u32 q, %2, %5;\n\t"
      "mov.b64 {x0, x1}, q;\n\t"
      "mul.wide.u32 q, %3, %4;\n\t"
      "mov.b64 {y0, y1}, q;\n\t"
      "mul.wide.u32 q, %3, %5;\n\t"
      "mov.b64 {z0, z1}, q;\n\t"
      "add.cc.u32 t0, x0, y0;\n\t"
      "addc.cc.u32 t1, x1, y1;\n\t"
      "addc.u32 t2, z1, 0;\n\t"      // z1 + carry of the cross sum
      "add.cc.u32 s1, t0, p1;\n\t"
      "addc.cc.u32 u, t1, z0;\n\t"
      "addc.u32 h1, t2, 0;\n\t"
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
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return ((u64)r1 << 32) | r0;
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
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  Words128 w;
  asm("{\n\t"
      ".reg .u32 p1, x0, x1, y0, y1, z0, z1, t0, t1, t2;\n\t"
      ".reg .u64 q;\n\t"
      "mul.wide.u32 q, %4, %6;\n\t"
      "mov.b64 {%0, p1}, q;\n\t"
      "mul.wide.u32 q, %4, %7;\n\t"
      "mov.b64 {x0, x1}, q;\n\t"
      "mul.wide.u32 q, %5, %6;\n\t"
      "mov.b64 {y0, y1}, q;\n\t"
      "mul.wide.u32 q, %5, %7;\n\t"
      "mov.b64 {z0, z1}, q;\n\t"
      "add.cc.u32 t0, x0, y0;\n\t"
      "addc.cc.u32 t1, x1, y1;\n\t"
      "addc.u32 t2, z1, 0;\n\t"
      "add.cc.u32 %1, t0, p1;\n\t"
      "addc.cc.u32 %2, t1, z0;\n\t"
      "addc.u32 %3, t2, 0;\n\t"
      "}"
      : "=r"(w.p0), "=r"(w.s1), "=r"(w.u), "=r"(w.h1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
  return w;
}
// a * a: the cross product a0 a1 is formed once (3 IMAD.WIDE.U32), (t0, t1, t2) = 2 x.
__host__ __device__ __forceinline__ u64 sqr_lazy(u64 a) {
#if defined(__CUDA_ARCH__) && !defined(VPBS_MUL_C)
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32);
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 p0, p1, x0, x1, z0, z1, t0, t1, t2, s1, u, h1, v, vv, w, cw, lo, hi, bm;\n\t"
      ".reg .u64 q;\n\t"
      "mul.wide.u32 q, %2, %2;\n\t"
      "mov.b64 {p0, p1}, q;\n\t"
      "mul.wide.u32 q, %2, %3;\n\t"
      "mov.b64 {x0, x1}, q;\n\t"
      "mul.wide.u32 q, %3, %3;\n\t"
      "mov.b64 {z0, z1}, q;\n\t"
      "add.cc.u32 t0, x0, x0;\n\t"
      "addc.cc.u32 t1, x1, x1;\n\t"
      "addc.u32 t2, z1, 0;\n\t"
      "add.cc.u32 s1, t0, p1;\n\t"
      "addc.cc.u32 u, t1, z0;\n\t"
      "addc.u32 h1, t2, 0;\n\t"
      "add.cc.u32 v, s1, u;\n\t"
      "addc.u32 vv, v, 0;\n\t"
      "addc.cc.u32 w, u, h1;\n\t"
      "addc.u32 cw, 0, 0;\n\t"
      "sub.cc.u32 lo, p0, w;\n\t"
      "subc.cc.u32 hi, vv, cw;\n\t"
      "subc.u32 bm, 0, 0;\n\t"
      "sub.cc.u32 %0, lo, bm;\n\t"
      "subc.u32 %1, hi, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(a0), "r"(a1));
  return ((u64)r1 << 32) | r0;
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
