// poseidon.cuh — width-12 Poseidon permutation over Goldilocks, state held in registers.
//
// Replaces [P2] plonky2 0.2.0 src/hash/poseidon.rs (Poseidon::poseidon: 4 full + 22 partial + 4
// full rounds, S-box x^7) with the constants of src/hash/poseidon_goldilocks.rs
// (MDS_MATRIX_CIRC / MDS_MATRIX_DIAG / ALL_ROUND_CONSTANTS).  The reference selects it through
// `C = PoseidonGoldilocksConfig` (/root/reference/src/main.rs:34) and calls the same primitive
// natively at /root/reference/src/vtfhe/ivc_based_vpbs.rs:73.
//
// Design (one permutation per thread):
//  * the 12 state words live in 24 32-bit registers; the MDS layer works on the 32-bit halves:
//    out_r = sum_i c_i * lo(s_{i+r}) + 2^32 * sum_i c_i * hi(s_{i+r}); every sum stays below 2^42
//    (c_i <= 41), i.e. exact in a double, so the layer is DFMA work on the FP64 pipe while the
//    S-boxes keep the integer pipes busy.  No 64x64 multiply is spent on the linear layer.
//  * the next round's constants are the initial value of the accumulators, so the constant layer
//    costs no extra instructions.
//  * S-box x^7 = (x^2 * x) * (x^2)^2: four 64x64 -> 128 products (gl64.cuh), the last one left
//    unreduced in the full rounds.
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

// x^7 with the last product left unreduced (four 32-bit words): its only consumer in a full round is
// the linear layer, which takes the words apart on the FP64 pipe (mds_absorb_words) — the integer
// reduction of that product (7 ALU-pipe instructions, the pipe that bounds the full rounds) is not
// executed at all.
__device__ __forceinline__ gl::Words128 sbox7_words(u64 x) {
  u64 x2 = gl::sqr_lazy(x);
  u64 x4 = gl::sqr_lazy(x2);
  u64 x3 = gl::mul_lazy(x, x2);
  return gl::mul_words(x3, x4);
}

// value = lo + 2^32 * hi with lo < 2^55, hi < 2^43  ->  arbitrary-u64 representative mod p.
//   hi = hh * 2^32 + hl:  value = lo + hh * (2^32 - 1) + hl * 2^32   (2^64 = 2^32 - 1 mod p)
__device__ __forceinline__ u64 reduce96(u64 lo, u64 hi) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 hl, hh, t0, t1, bm;\n\t"
      ".reg .u64 t;\n\t"
      "mov.b64 {hl, hh}, %3;\n\t"
      "mov.b64 {t0, t1}, %2;\n\t"
      "sub.cc.u32 t0, t0, hh;\n\t"               // lo + hh * (2^32 - 1) = lo - hh + (hh << 32)
      "subc.u32 t1, t1, 0;\n\t"                  // (exact mod 2^64; the true value is < 2^56)
      "add.u32 t1, t1, hh;\n\t"
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

// Recombine two biased accumulators into one state word, with NO conditional fix-up.
//   dlo = 2^52 + L, dhi = 2^52 + H  (L, H < 2^42 exact integers): the low words of the doubles are
//   l0 = L mod 2^32 and h0 = H mod 2^32, the high words are 0x43300000 + l1 and 0x43300000 + h1
//   with l1, h1 < 2^10.   value = L + 2^32 H = l0 + (l1 + h0) 2^32 + h1 2^64,  2^64 = 2^32 - 1:
//       value = (l0 - h1 - c) + (m + c) 2^32,   m + c 2^32 = l1 + h0 + h1.
//   The low part can only borrow when h1 + c >= 1, and then m + c >= 1 absorbs the borrow; the
//   high part cannot exceed 2^32 - 1 (if c = 1 then m < 2^11): the result is exact as is.
__device__ __forceinline__ u64 combine_biased(double dlo, double dhi) {
  u32 r0, r1;
  asm("{\n\t"
      ".reg .u32 l0, lh, h0, hh, h1, s1, m, m2, kc;\n\t"
      "mov.b64 {l0, lh}, %2;\n\t"
      "mov.b64 {h0, hh}, %3;\n\t"
      "add.u32 h1, hh, 0xbcd00000;\n\t"      // hh - 0x43300000
      "add.u32 s1, lh, h1;\n\t"
      "add.u32 s1, s1, 0xbcd00000;\n\t"      // l1 + h1
      "add.cc.u32 m, s1, h0;\n\t"            // carry c
      "addc.u32 kc, h1, 0;\n\t"              // h1 + c   (addc without .cc leaves the flag alone:
      "addc.u32 m2, m, 0;\n\t"               // m + c     both read the same carry; 6 SASS instructions)
      "sub.cc.u32 %0, l0, kc;\n\t"
      "subc.u32 %1, m2, 0;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "d"(dlo), "d"(dhi));
  return ((u64)r1 << 32) | r0;
}

// ---- MDS layer --------------------------------------------------------------------------------
// The linear layer runs on the FP64 pipe as a split convolution with pair-merged partial rounds
// (below).  What was measured against it on B200 (tools/variants/): a dense FP64
// form (288 DFMA per layer) and a pure 32-bit integer form on 22/21/21-bit limbs, both bit-exact
// and both slower (7.80 and 8.21 ms against 5.32 ms for 8.39 M permutations in round 1).
// Round constants pre-split for the accumulators: RCD[2*(12*r+i)] = 2^52 + lo32(RC),
// RCD[2*(12*r+i)+1] = 2^52 + hi32(RC) (exact doubles; row 30 is the all-zero "no next round").
__constant__ double RCD[2 * (ROUNDS + 1) * WIDTH] = {
#include "poseidon_rcd.inc"
};
// The same table in global memory for permute_coop, whose threads index it divergently (the
// constant cache serialises distinct addresses within a warp; L1 does not).
__device__ const double RCD_G[2 * (ROUNDS + 1) * WIDTH] = {
#include "poseidon_rcd.inc"
};


// ---- MDS layer on the FP64 pipe ---------------------------------------------------------------
// On B200 a 64-bit integer multiply-accumulate costs three issue slots (IMAD.WIDE.U32 runs at half
// rate on the fmaheavy pipe and ptxas splits its 64-bit addend into IADD3 + IADD3.X), while DFMA is
// one full-rate slot on its own pipe with a 53-bit exact integer range (measured:
// tools/microbench.cu, profiles/).  The MDS constants are <= 41 and each state word is split in
// 32-bit halves, so every partial sum RC + sum_i c_i * half_i stays below 2^42: exact in doubles.
//   * half -> double: bit pattern {half, 0x43300000} is 2^52 + half; subtract 2^52.
//   * the accumulator starts at 2^52 + RC_half, so the final mantissa *is* the integer sum:
//     no conversion back, just mask off the exponent bits.
__device__ __forceinline__ double half_to_f64(u32 x) {
  return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0;
}

// The MDS matrix is circulant (plus one diagonal entry), i.e. y = c (*) s is a length-12 cyclic
// convolution.  x^12 - 1 = (x^6 - 1)(x^6 + 1) splits it into a cyclic and a negacyclic length-6
// convolution on p_t = s_t + s_{t+6} and m_t = s_t - s_{t+6}:
//     Z+_r = sum_j (c_j + c_{j+6})/2 p_{(j+r) mod 6},  Z-_r = sum_j (+-)(c_j - c_{j+6})/2 m_{(j+r) mod 6}
//     y_r = Z+_r + Z-_r,   y_{r+6} = Z+_r - Z-_r                          (r < 6)
// 144 DFMA per round instead of 288 (plus 24 + 24 adds).  Everything stays exact: the halved
// coefficients are integers (c+ and c- are even), |Z| < 2^42, and Z+ starts at 2^52 + its share
// of the round constants, so Z+ + Z- and Z+ - Z- are the biased results directly (the mantissa is
// the integer sum; no scaling or conversion back).
// RCS[24 * round + 4 * r + {0,1,2,3}] = {2^52 + (al+bl)/2, (al-bl)/2, 2^52 + (ah+bh)/2, (ah-bh)/2}
// where (al, ah), (bl, bh) are half-splits of RC[r], RC[r+6] (value = l + 2^32 h mod p) chosen by
// tools/gen_poseidon_consts.py so that both sums are even (row 30: RC = 0).
__constant__ double RCS[(ROUNDS + 1) * 24] = {
#include "poseidon_rcs.inc"
};

namespace mds_split {
constexpr int CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
// HALVED sums / differences: c_j + c_{j+6} and c_j - c_{j+6} are all even for this matrix
// (30 28 80 34 36 48 / 4 2 2 -2 -32 8), so Z+ and Z- below are the halves of the textbook ones and
// y_r = Z+_r + Z-_r, y_{r+6} = Z+_r - Z-_r come out without the 1/2.
static_assert((CIRC[0] + CIRC[6]) % 2 == 0 && (CIRC[1] + CIRC[7]) % 2 == 0 && (CIRC[2] + CIRC[8]) % 2 == 0 &&
              (CIRC[3] + CIRC[9]) % 2 == 0 && (CIRC[4] + CIRC[10]) % 2 == 0 && (CIRC[5] + CIRC[11]) % 2 == 0,
              "halved split convolution needs even c+ / c-");
__host__ __device__ constexpr double cplus(int j) { return (double)((CIRC[j] + CIRC[j + 6]) / 2); }
__host__ __device__ constexpr double cminus(int j) { return (double)((CIRC[j] - CIRC[j + 6]) / 2); }

// input t (p_t, m_t) into every accumulator r: the term j with (j + r) mod 6 == t
template <int T, int R>
__device__ __forceinline__ void col(double pl, double ph, double ml, double mh, double (&zpl)[6],
                                    double (&zph)[6], double (&zml)[6], double (&zmh)[6]) {
  constexpr int J = (T - R + 6) % 6;
  constexpr double CP = cplus(J);
  constexpr double CM = (J + R < 6) ? cminus(J) : -cminus(J);
  zpl[R] = fma(pl, CP, zpl[R]);
  zph[R] = fma(ph, CP, zph[R]);
  zml[R] = fma(ml, CM, zml[R]);
  zmh[R] = fma(mh, CM, zmh[R]);
  if constexpr (R + 1 < 6) col<T, R + 1>(pl, ph, ml, mh, zpl, zph, zml, zmh);
}
template <int T>
__device__ __forceinline__ void cols(const u64 (&s)[WIDTH], double (&zpl)[6], double (&zph)[6],
                                     double (&zml)[6], double (&zmh)[6], double& x0l, double& x0h) {
  const double al = half_to_f64((u32)s[T]), ah = half_to_f64((u32)(s[T] >> 32));
  const double bl = half_to_f64((u32)s[T + 6]), bh = half_to_f64((u32)(s[T + 6] >> 32));
  if constexpr (T == 0) {
    x0l = al;
    x0h = ah;
  }
  // (forming p, m on the integer side with carry into the exponent word was tried: the carry
  // costs ISETP + SEL per value and measured 1 % slower than these four DADDs)
  col<T, 0>(al + bl, ah + bh, al - bl, ah - bh, zpl, zph, zml, zmh);
  if constexpr (T + 1 < 6) cols<T + 1>(s, zpl, zph, zml, zmh, x0l, x0h);
}
}  // namespace mds_split

struct MdsAcc {
  double zpl[6], zph[6], zml[6], zmh[6], x0l, x0h;
};
__device__ __forceinline__ void mds_begin(MdsAcc& a, int next_round) {
  const double4* __restrict__ rcs = reinterpret_cast<const double4*>(RCS) + 6 * next_round;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const double4 c = rcs[r];
    a.zpl[r] = c.x;
    a.zml[r] = c.y;
    a.zph[r] = c.z;
    a.zmh[r] = c.w;
  }
}
// feed state words T and T + 6 (final for this round) into all accumulators
// biased view of a 32-bit half: the double 2^52 + x (no arithmetic, just a register pair)
__device__ __forceinline__ double half_biased(u32 x) { return __hiloint2double(0x43300000, (int)x); }
template <int T>
__device__ __forceinline__ void mds_absorb(MdsAcc& a, const u64 (&s)[WIDTH]) {
  // p = x_T + x_{T+6}, m = x_T - x_{T+6} straight from the biased views: (2^52 + a) - (2^52 + b)
  // is a - b, and (2^52 + a) + ((2^52 + b) - 2^53) is a + b; every step is exact (|.| < 2^53).
  const double TWO53 = 9007199254740992.0;
  const double cal = half_biased((u32)s[T]), cah = half_biased((u32)(s[T] >> 32));
  const double cbl = half_biased((u32)s[T + 6]), cbh = half_biased((u32)(s[T + 6] >> 32));
  if constexpr (T == 0) {
    a.x0l = cal - 4503599627370496.0;
    a.x0h = cah - 4503599627370496.0;
  }
  mds_split::col<T, 0>(cal + (cbl - TWO53), cah + (cbh - TWO53), cal - cbl, cah - cbh, a.zpl, a.zph,
                       a.zml, a.zmh);
}
// The same for two state words given as unreduced 128-bit products (p0, s1, u, h1):
//   p0 + s1 2^32 + u 2^64 + h1 2^96 == L + H 2^32 (mod p),  L = p0 - u - h1,  H = s1 + u,
// with L in (-2^34, 2^32) and H in [0, 2^33): the layer is linear, so (L, H) serve as the two
// "halves" as they are.  L and H are formed on the INTEGER side, directly as the bit patterns of
// biased doubles — pattern(2^52 + 2^40) + L is the double 2^52 + 2^40 + L and pattern(2^52) + H the
// double 2^52 + H (no exponent change in either range) — two or three IADD3 per value on the
// register pairs the product already lives in.  (Round 1 built four 2^52-biased views per product
// and took them apart with five DADDs: every view costs two register moves to assemble a pair
// with the constant high word; 52 instructions per full round more, 5.25 -> 5.15 ms.)  The sums and
// differences below are exact DADDs (every intermediate is an integer below 2^53); Z+ starts 2^42
// higher in its low half and 2^10 lower in its high half (value-neutral, poseidon_rcs.inc) so that
// the low sums stay positive.
template <int T>
__device__ __forceinline__ void mds_absorb_words(MdsAcc& a, const gl::Words128& wa, const gl::Words128& wb) {
  constexpr u64 PAT_L = 0x4330010000000000ULL, PAT_H = 0x4330000000000000ULL;
  const double BL = 4503599627370496.0 + 1099511627776.0, BH = 4503599627370496.0;  // 2^52 + 2^40, 2^52
  const double al = __longlong_as_double((long long)(PAT_L + (u64)wa.p0 - (u64)wa.u - (u64)wa.h1));
  const double ah = __longlong_as_double((long long)(PAT_H + (u64)wa.s1 + (u64)wa.u));
  const double bl = __longlong_as_double((long long)(PAT_L + (u64)wb.p0 - (u64)wb.u - (u64)wb.h1));
  const double bh = __longlong_as_double((long long)(PAT_H + (u64)wb.s1 + (u64)wb.u));
  if constexpr (T == 0) {
    a.x0l = al - BL;
    a.x0h = ah - BH;
  }
  mds_split::col<T, 0>(al + (bl - 2.0 * BL), ah + (bh - 2.0 * BH), al - bl, ah - bh, a.zpl, a.zph, a.zml,
                       a.zmh);
}
__device__ __forceinline__ void mds_finish(MdsAcc& a, u64 (&s)[WIDTH]) {
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double s1l = a.zpl[r] + a.zml[r], s1h = a.zph[r] + a.zmh[r];        // 2^52 + y_r
    const double s2l = a.zpl[r] - a.zml[r], s2h = a.zph[r] - a.zmh[r];  // 2^52 + y_{r+6}
    if (r == 0) {  // MDS_MATRIX_DIAG = [8, 0, ..., 0]
      s1l = fma(a.x0l, 8.0, s1l);
      s1h = fma(a.x0h, 8.0, s1h);
    }
    s[r] = combine_biased(s1l, s1h);  // Z+ starts at 2^52 + ..., so sums and differences are biased
    s[r + 6] = combine_biased(s2l, s2h);
  }
}
// ---- two partial rounds at once ---------------------------------------------------------------------
// In a partial round only lane 0 goes through the S-box, so two consecutive partial rounds are
//   u  = (sbox(s_0), s_1 .. s_11)                x'  = M u + RC_{r+1}
//   x'' = M (x' + delta e_0) + RC_{r+2},         delta = sbox(x'_0) - x'_0
//       = M^2 u + (M RC_{r+1} + RC_{r+2}) + delta M e_0.
// With M = C + 8 e_0 e_0^T (C circulant):  M^2 u = C^2 u + 8 u_0 C e_0 + 8 (M u)_0 e_0, and
// (M u)_0 + delta = sbox(x'_0) - RC_{r+1,0}, so
//   x'' = C^2 u + (8 u_0 + delta) C e_0 + 8 sbox(x'_0) e_0 + K',
//   K'  = M RC_{r+1} + RC_{r+2} - 8 RC_{r+1,0} e_0    (precomputed, poseidon_rcp.inc).
// C^2 is circulant too (constants c (*) c, row sum 2^16), so ONE split convolution serves both
// rounds: the entries are < 2^13, every partial sum stays below 2^50 and is exact in a double.  Per
// pair this is ~290 FP64 instructions and 13 recombinations instead of 2 x (236 and 12); x'_0
// itself only needs row 0 of C u, which falls out of the same p_t / m_t inputs
// ((C u)_0 = 1/2 sum_t (c+_t p_t + c-_t m_t)).
__constant__ double RCP[11 * 24] = {
#include "poseidon_rcp.inc"
};
namespace mds_pair {
constexpr int CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
__host__ __device__ constexpr int cc(int d) {  // (c (*) c)_d, cyclic
  int v = 0;
  for (int a = 0; a < WIDTH; a++) v += CIRC[a] * CIRC[(d - a + WIDTH) % WIDTH];
  return v;
}
// halved like mds_split: (c (*) c)+ = c+ (*) c+ and (c (*) c)- = c- (*) c- are multiples of 4
__host__ __device__ constexpr double ccplus(int j) { return (double)((cc(j) + cc(j + 6)) / 2); }
__host__ __device__ constexpr double ccminus(int j) { return (double)((cc(j) - cc(j + 6)) / 2); }
template <int T, int R>
__device__ __forceinline__ void col(double pl, double ph, double ml, double mh, double (&zpl)[6],
                                    double (&zph)[6], double (&zml)[6], double (&zmh)[6]) {
  constexpr int J = (T - R + 6) % 6;
  constexpr double CP = ccplus(J);
  constexpr double CM = (J + R < 6) ? ccminus(J) : -ccminus(J);
  zpl[R] = fma(pl, CP, zpl[R]);
  zph[R] = fma(ph, CP, zph[R]);
  zml[R] = fma(ml, CM, zml[R]);
  zmh[R] = fma(mh, CM, zmh[R]);
  if constexpr (R + 1 < 6) col<T, R + 1>(pl, ph, ml, mh, zpl, zph, zml, zmh);
}
// words T and T + 6 of u into the C^2 accumulators and into row 0 of C u (xl, xh)
template <int T>
__device__ __forceinline__ void absorb(MdsAcc& a, const u64 (&s)[WIDTH], double& xl, double& xh) {
  const double TWO53 = 9007199254740992.0;
  const double cal = half_biased((u32)s[T]), cah = half_biased((u32)(s[T] >> 32));
  const double cbl = half_biased((u32)s[T + 6]), cbh = half_biased((u32)(s[T + 6] >> 32));
  const double pl = cal + (cbl - TWO53), ph = cah + (cbh - TWO53), ml = cal - cbl, mh = cah - cbh;
  if constexpr (T == 0) {
    a.x0l = cal - 4503599627370496.0;
    a.x0h = cah - 4503599627370496.0;
  }
  constexpr double HP = mds_split::cplus(T), HM = mds_split::cminus(T);  // already halved
  xl = fma(pl, HP, fma(ml, HM, xl));
  xh = fma(ph, HP, fma(mh, HM, xh));
  col<T, 0>(pl, ph, ml, mh, a.zpl, a.zph, a.zml, a.zmh);
}
}  // namespace mds_pair

// ptxas schedules a full round as "all twelve S-boxes (integer pipes), then all of the linear layer
// (FP64 pipe)".  A real, always-zero data dependence from the accumulators of lane pair T into the
// S-box inputs of lane pair T + 2 (one LOP3 each: x | (sign bit of a 2^52-biased, hence positive,
// accumulator)) makes it interleave S-box T + 1 with the accumulation of pair T, which also shortens
// the live ranges (115 -> 96 registers).  Measured: 5.44 -> 5.38 ms; the same staging inside the
// partial-round pairs (S-box chain beside the other lanes' accumulation) measured no difference —
// the pipes already overlap across the warps of an SMSP.
__device__ __forceinline__ void tie_to(u64& x, double positive) {
  u32 lo = (u32)x;
  const u32 hi = (u32)(x >> 32);
  asm("lop3.b32 %0, %0, %1, 0x80000000, 0xf8;" : "+r"(lo) : "r"((u32)__double2hiint(positive)));
  x = ((u64)hi << 32) | lo;
}
#define VPBS_TIE(x, y, acc) \
  tie_to(x, (acc).zpl[5]);  \
  tie_to(y, (acc).zph[5]);

// Partial rounds r = 4 + 2 * pair and r + 1.  In: state with RC_r added; out: state with RC_{r+2}.
__device__ __forceinline__ void partial_pair(u64 (&s)[WIDTH], int pair) {
  MdsAcc acc;
  {
    const double4* __restrict__ k = reinterpret_cast<const double4*>(RCP) + 6 * pair;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const double4 c = k[r];
      acc.zpl[r] = c.x;
      acc.zml[r] = c.y;
      acc.zph[r] = c.z;
      acc.zmh[r] = c.w;
    }
  }
  // x'_0 = RC_{r+1,0} + 8 u_0 + (C u)_0, accumulated on top of the biased constant
  const int r1 = FULL_ROUNDS_HALF + 2 * pair + 1;
  double xl = RCD[2 * WIDTH * r1], xh = RCD[2 * WIDTH * r1 + 1];
  s[0] = sbox7(s[0]);                      // u_0
  mds_pair::absorb<0>(acc, s, xl, xh);
  xl = fma(acc.x0l, 8.0, xl);
  xh = fma(acc.x0h, 8.0, xh);
  mds_pair::absorb<1>(acc, s, xl, xh);
  mds_pair::absorb<2>(acc, s, xl, xh);
  mds_pair::absorb<3>(acc, s, xl, xh);
  mds_pair::absorb<4>(acc, s, xl, xh);
  mds_pair::absorb<5>(acc, s, xl, xh);
  const u64 x1 = gl::canon(combine_biased(xl, xh));  // x'_0
  const u64 u1 = sbox7(x1);                          // sbox(x'_0), any representative
  const u64 delta = gl::sub_lazy(u1, x1);            // x1 canonical; the halves only feed linear terms
  // A = 8 u_0 + delta, and sbox(x'_0), as plain doubles per half
  const double al = fma(acc.x0l, 8.0, half_to_f64((u32)delta));
  const double ah = fma(acc.x0h, 8.0, half_to_f64((u32)(delta >> 32)));
  const double u1l = half_to_f64((u32)u1), u1h = half_to_f64((u32)(u1 >> 32));
  constexpr double G[WIDTH] = {17, 20, 34, 18, 39, 13, 13, 28, 2, 16, 41, 15};  // (C e_0)_r = c_{-r}
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const double s1l = acc.zpl[r] + acc.zml[r], s1h = acc.zph[r] + acc.zmh[r];  // 2^52 + (C^2 u + K')_r
    const double s2l = acc.zpl[r] - acc.zml[r], s2h = acc.zph[r] - acc.zmh[r];  // ... _{r+6}
    double d1l = fma(al, G[r], s1l), d1h = fma(ah, G[r], s1h);
    const double d2l = fma(al, G[r + 6], s2l), d2h = fma(ah, G[r + 6], s2h);
    if (r == 0) {
      d1l = fma(u1l, 8.0, d1l);
      d1h = fma(u1h, 8.0, d1h);
    }
    s[r] = combine_biased(d1l, d1h);
    s[r + 6] = combine_biased(d2l, d2h);
  }
}
// The whole layer in one call (tools/selftest.cu).
__device__ __forceinline__ void mds_add_rc(u64 (&s)[WIDTH], int next_round) {
  MdsAcc acc;
  mds_begin(acc, next_round);
  mds_absorb<0>(acc, s);
  mds_absorb<1>(acc, s);
  mds_absorb<2>(acc, s);
  mds_absorb<3>(acc, s);
  mds_absorb<4>(acc, s);
  mds_absorb<5>(acc, s);
  mds_finish(acc, s);
}


// ---- the permutation -----------------------------------------------------------------------------
// In-place; input words arbitrary u64, output words arbitrary u64 (lazy).  The four full rounds of
// each half share one loop body and the eleven partial-round pairs another (~27 KB of code in all),
// which fits the 32 KB instruction cache; unrolled round bodies did not (ncu: 13 % icache misses,
// stall_no_instruction 1.3 per issue).  Measured alternatives (out-of-line S-boxes, integer and
// dense FP64 linear layers, reduced S-box outputs) live in tools/variants/, not here.
__device__ __forceinline__ void full_round(u64 (&s)[WIDTH], int r) {
  // S-boxes (integer pipes) and MDS accumulation (FP64 pipe) interleaved pair by pair; the last
  // product of every S-box stays unreduced (sbox7_words / mds_absorb_words).
  MdsAcc acc;
  mds_begin(acc, r + 1);
  {
    const gl::Words128 a0 = sbox7_words(s[0]), b0 = sbox7_words(s[6]);
    mds_absorb_words<0>(acc, a0, b0);
    const gl::Words128 a1 = sbox7_words(s[1]), b1 = sbox7_words(s[7]);
    VPBS_TIE(s[2], s[8], acc)
    mds_absorb_words<1>(acc, a1, b1);
    const gl::Words128 a2 = sbox7_words(s[2]), b2 = sbox7_words(s[8]);
    VPBS_TIE(s[3], s[9], acc)
    mds_absorb_words<2>(acc, a2, b2);
    const gl::Words128 a3 = sbox7_words(s[3]), b3 = sbox7_words(s[9]);
    VPBS_TIE(s[4], s[10], acc)
    mds_absorb_words<3>(acc, a3, b3);
    const gl::Words128 a4 = sbox7_words(s[4]), b4 = sbox7_words(s[10]);
    VPBS_TIE(s[5], s[11], acc)
    mds_absorb_words<4>(acc, a4, b4);
    const gl::Words128 a5 = sbox7_words(s[5]), b5 = sbox7_words(s[11]);
    mds_absorb_words<5>(acc, a5, b5);
  }
  mds_finish(acc, s);
}

__device__ __forceinline__ void permute_lazy(u64 (&s)[WIDTH]) {
#pragma unroll
  for (int i = 0; i < WIDTH; i++) s[i] = gl::add_lazy(s[i], RC[i]);
  // 4 full rounds, 11 pairs of partial rounds, 4 full rounds; one copy of each loop body.
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int k = 0; k < FULL_ROUNDS_HALF; k++)
      full_round(s, half * (FULL_ROUNDS_HALF + PARTIAL_ROUNDS) + k);
    if (half == 0) {
#pragma unroll 1
      for (int p = 0; p < PARTIAL_ROUNDS / 2; p++) partial_pair(s, p);
    }
  }
}

// ---- latency-optimised permutation: 12 threads of a 16-thread group share one state -----------
// The top levels of a Merkle tree have too few nodes to fill the machine, and a lone thread needs
// ~40 us per permutation (27 k dependent-ish instructions).  Here thread l (< 12) of a group owns
// state word l: the S-boxes of a full round run in parallel, and each thread computes its own MDS
// output from the other words' halves, exchanged as doubles through shared memory.
// `sh` = 48 doubles of shared memory private to the group (low halves at [0,24), high halves at
// [24,48), each stored twice so that the rotation (l + i) needs no wrap-around).
// All 32 threads of the warp must call this together (it uses __syncwarp()).
__device__ __forceinline__ u64 mad_wide(u32 x, u32 c, u64 acc) {
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(c));
  return acc;
}
// The MDS layer here is the INTEGER form (IMAD.WIDE chains on the 32-bit halves): this routine is
// latency bound, a dependent DFMA costs ~64 cycles on this part against ~6-10 for IMAD.WIDE + carry,
// and the FP64 form spent three quarters of every round waiting on its accumulation chains (one
// tree level 27 us -> see DESIGN.md 4.3).  Throughput does not matter at these sizes.
__device__ __forceinline__ void permute_coop(u64& w, double* __restrict__ sh_d, unsigned l) {
  u64* __restrict__ sh = reinterpret_cast<u64*>(sh_d);  // [0,24): the 12 words, stored twice
  const bool active = l < WIDTH;
  const unsigned ll = active ? l : 0;
  auto rc_of = [&](int r) -> u64 {  // RC[r][l] from the split table (2^52-biased halves)
    const double2 rc = __ldg(reinterpret_cast<const double2*>(RCD_G) + (WIDTH * r + ll));
    return ((u64)(u32)__double2loint(rc.y) << 32) | (u32)__double2loint(rc.x);
  };
  if (active) w = gl::add_lazy(w, rc_of(0));
  constexpr u32 CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#pragma unroll 1
  for (int r = 0; r < ROUNDS; r++) {
    const bool full = r < FULL_ROUNDS_HALF || r >= FULL_ROUNDS_HALF + PARTIAL_ROUNDS;
    // next round's constants: issued before the S-box so that the load latency hides behind it
    const u64 rc = rc_of(r + 1);
    if (active && (full || l == 0)) w = sbox7(w);
    if (active) {
      sh[l] = w;
      sh[l + 12] = w;
    }
    __syncwarp();
    if (active) {
      u64 al0 = (u32)rc, ah0 = rc >> 32, al1 = 0, ah1 = 0;  // two chains per half: shorter latency
#pragma unroll
      for (int i = 0; i < WIDTH; i += 2) {
        const u64 x0 = sh[l + i], x1 = sh[l + i + 1];
        al0 = mad_wide((u32)x0, CIRC[i], al0);
        ah0 = mad_wide((u32)(x0 >> 32), CIRC[i], ah0);
        al1 = mad_wide((u32)x1, CIRC[i + 1], al1);
        ah1 = mad_wide((u32)(x1 >> 32), CIRC[i + 1], ah1);
      }
      if (l == 0) {  // MDS_MATRIX_DIAG = [8, 0, ..., 0]
        const u64 x0 = sh[0];
        al1 = mad_wide((u32)x0, 8u, al1);
        ah1 = mad_wide((u32)(x0 >> 32), 8u, ah1);
      }
      w = reduce96(al0 + al1, ah0 + ah1);  // each sum < 2^43
    }
    __syncwarp();
  }
}

}  // namespace poseidon
