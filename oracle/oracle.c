/*
 * oracle.c — CPU restatement of plonky2 0.2.0's commitment path (see oracle.h header).
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product library.
 *
 * Plain C11 + OpenMP.  Parallel regions mirror where plonky2 uses rayon: over columns for the
 * FFTs ([P2] fri/oracle.rs from_values / lde_values), over rows for transpose, over leaves and
 * sibling pairs for the Merkle tree ([P2] hash/merkle_tree.rs fill_digests_buf / fill_subtree).
 */
#include "oracle.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define P ORC_P
#define EPS 0xFFFFFFFFULL

/* ------------------------------------------------------------------------------------------
 * Goldilocks field.  [P2] plonky2_field/src/goldilocks_field.rs
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t canon(uint64_t x) { return x >= P ? x - P : x; }

/* [P2] goldilocks_field.rs reduce128: x = lo + 2^64*hi, 2^64 = EPS, 2^96 = -1 (mod p). */
static inline uint64_t reduce128(u128 x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
  uint64_t t0, r;
  /* branch-free: the carry/borrow is taken ~half the time, so branches would mispredict */
  uint64_t borrow = __builtin_sub_overflow(lo, hi_hi, &t0);
  t0 -= EPS & (0 - borrow); /* borrow: add p */
  uint64_t t1 = hi_lo * EPS;
  uint64_t carry = __builtin_add_overflow(t0, t1, &r);
  r += EPS & (0 - carry); /* carry: subtract p */
  return r;             /* < 2^64, not necessarily canonical */
}
static inline uint64_t add_(uint64_t a, uint64_t b) { /* a,b canonical */
  uint64_t s = a + b;
  if (s < a || s >= P) s -= P;
  return s;
}
static inline uint64_t sub_(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (P - b); }
static inline uint64_t mul_(uint64_t a, uint64_t b) { return canon(reduce128((u128)a * b)); }

uint64_t orc_gl_add(uint64_t a, uint64_t b) { return add_(canon(a), canon(b)); }
uint64_t orc_gl_sub(uint64_t a, uint64_t b) { return sub_(canon(a), canon(b)); }
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return mul_(a, b); }
uint64_t orc_gl_pow(uint64_t a, uint64_t e) {
  uint64_t r = 1, b = canon(a);
  while (e) {
    if (e & 1) r = mul_(r, b);
    b = mul_(b, b);
    e >>= 1;
  }
  return r;
}
uint64_t orc_gl_inv(uint64_t a) { return orc_gl_pow(a, P - 2); }

/* [P2] GoldilocksField::POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32); two-adicity 32; generator 7. */
#define POWER_OF_TWO_GENERATOR 1753635133440165772ULL
#define COSET_SHIFT 7ULL
uint64_t orc_primitive_root_of_unity(unsigned n_log) {
  uint64_t b = POWER_OF_TWO_GENERATOR;
  for (unsigned i = n_log; i < 32; i++) b = mul_(b, b);
  return b;
}

static int g_threads = 0;
void orc_set_threads(int n) { g_threads = n; }
int orc_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * FFT.  [P2] plonky2_field/src/fft.rs fft_dispatch -> fft_classic: reverse_index_bits_in_place,
 * then lg n radix-2 decimation-in-time layers reading root_table[lg_m-1][j] = w_{2^lg_m}^j.
 * (zero_factor only skips butterflies whose inputs are known zeros: same values.)
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t bitrev(uint64_t x, unsigned bits) {
  uint64_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1ULL) << (bits - 1 - i);
  return r;
}
static void reverse_index_bits_in_place_u64(uint64_t* v, unsigned log_n) {
  uint64_t n = 1ULL << log_n;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t j = bitrev(i, log_n);
    if (i < j) {
      uint64_t t = v[i];
      v[i] = v[j];
      v[j] = t;
    }
  }
}
/* roots[j] = w_n^j for j < n/2 (the last row of plonky2's fft_root_table; row lg_m-1 is this
 * one read with stride n/2^lg_m). */
static uint64_t* make_roots(unsigned log_n) {
  uint64_t half = log_n ? (1ULL << (log_n - 1)) : 1;
  uint64_t* r = (uint64_t*)malloc(sizeof(uint64_t) * half);
  uint64_t w = orc_primitive_root_of_unity(log_n), acc = 1;
  for (uint64_t j = 0; j < half; j++) {
    r[j] = acc;
    acc = mul_(acc, w);
  }
  return r;
}
static void fft_classic(uint64_t* v, unsigned log_n, const uint64_t* roots) {
  uint64_t n = 1ULL << log_n;
  for (uint64_t i = 0; i < n; i++) v[i] = canon(v[i]);
  reverse_index_bits_in_place_u64(v, log_n);
  for (unsigned lg_half_m = 0; lg_half_m < log_n; lg_half_m++) {
    uint64_t half_m = 1ULL << lg_half_m, m = half_m << 1;
    uint64_t stride = n / m; /* w_m^j = w_n^(j*n/m) */
    for (uint64_t k = 0; k < n; k += m)
      for (uint64_t j = 0; j < half_m; j++) {
        uint64_t t = mul_(roots[j * stride], v[k + half_m + j]);
        uint64_t u = v[k + j];
        v[k + j] = add_(u, t);
        v[k + half_m + j] = sub_(u, t);
      }
  }
}
/* The same transform at the speed class of plonky2's packed-field fft_classic_simd: per-layer
 * contiguous twiddles ([P2] fft_root_table), SIMD butterflies (poseidon_simd.inc fft_layer), and
 * the first `zero_factor` layers as copies when the upper (2^zero_factor - 1)/2^zero_factor of the
 * input is zero padding ([P2] fft_classic's r parameter).  Used when orc_get_simd() > 1; the plain
 * loop above is its checker (tests/test_oracle_golden.py). */
static void fft_layer_x4(uint64_t* v, uint64_t n, uint64_t half_m, const uint64_t* tw);
static void fft_layer_x8(uint64_t* v, uint64_t n, uint64_t half_m, const uint64_t* tw);
int orc_get_simd(void);
static inline uint64_t brev64(uint64_t x) {
  x = __builtin_bswap64(x);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  return x;
}
/* layer_tw[lg][j] = w_{2^(lg+1)}^j, j < 2^lg, all layers in one allocation */
static uint64_t** make_layer_roots(unsigned log_n, const uint64_t* roots) {
  uint64_t n = 1ULL << log_n;
  uint64_t** t = (uint64_t**)malloc(sizeof(uint64_t*) * (log_n + 1) + sizeof(uint64_t) * (n + 8));
  uint64_t* data = (uint64_t*)(t + log_n + 1);
  for (unsigned lg = 0; lg < log_n; lg++) {
    t[lg] = data;
    uint64_t half_m = 1ULL << lg, stride = n / (2 * half_m);
    for (uint64_t j = 0; j < half_m; j++) data[j] = roots[j * stride];
    data += half_m;
  }
  return t;
}
static void fft_fast(uint64_t* v, unsigned log_n, uint64_t* const* layer_tw, unsigned zero_factor) {
  const uint64_t n = 1ULL << log_n;
  const int w = orc_get_simd();
  for (uint64_t i = 0; i < n; i++) { /* reverse_index_bits_in_place, canonicalising on the way */
    uint64_t j = log_n ? (brev64(i) >> (64 - log_n)) : 0;
    if (i < j) {
      uint64_t a = canon(v[i]), b = canon(v[j]);
      v[i] = b;
      v[j] = a;
    } else if (i == j) {
      v[i] = canon(v[i]);
    }
  }
  for (unsigned lg = 0; lg < log_n; lg++) {
    const uint64_t half_m = 1ULL << lg, m = half_m << 1;
    if (lg < zero_factor) { /* the second operand of every butterfly is zero padding */
      for (uint64_t k = 0; k < n; k += m)
        for (uint64_t j = 0; j < half_m; j++) v[k + half_m + j] = v[k + j];
    } else if (w == 8 && half_m >= 8) {
      fft_layer_x8(v, n, half_m, layer_tw[lg]);
    } else if (w >= 4 && half_m >= 4) {
      fft_layer_x4(v, n, half_m, layer_tw[lg]);
    } else {
      const uint64_t* tw = layer_tw[lg];
      for (uint64_t k = 0; k < n; k += m)
        for (uint64_t j = 0; j < half_m; j++) {
          uint64_t t = mul_(tw[j], v[k + half_m + j]);
          uint64_t u = v[k + j];
          v[k + j] = add_(u, t);
          v[k + half_m + j] = sub_(u, t);
        }
    }
  }
}
void orc_fft(uint64_t* v, unsigned log_n) {
  uint64_t* roots = make_roots(log_n);
  if (orc_get_simd() > 1) {
    uint64_t** lt = make_layer_roots(log_n, roots);
    fft_fast(v, log_n, lt, 0);
    free(lt);
  } else {
    fft_classic(v, log_n, roots);
  }
  free(roots);
}
/* [P2] fft.rs ifft_with_options: forward FFT, then reverse all values except the first and
 * scale by n^-1 (= F::inverse_2exp(lg n)). */
static void ifft_finish(uint64_t* v, unsigned log_n);
static void ifft_with_roots(uint64_t* v, unsigned log_n, const uint64_t* roots) {
  fft_classic(v, log_n, roots);
  ifft_finish(v, log_n);
}
static void ifft_finish(uint64_t* v, unsigned log_n) {
  uint64_t n = 1ULL << log_n;
  uint64_t n_inv = orc_gl_inv(n % P);
  v[0] = mul_(v[0], n_inv);
  if (n > 1) v[n / 2] = mul_(v[n / 2], n_inv);
  for (uint64_t i = 1; i < n / 2; i++) {
    uint64_t j = n - i;
    uint64_t ci = mul_(v[j], n_inv), cj = mul_(v[i], n_inv);
    v[i] = ci;
    v[j] = cj;
  }
}
void orc_ifft(uint64_t* v, unsigned log_n) {
  uint64_t* roots = make_roots(log_n);
  ifft_with_roots(v, log_n, roots);
  free(roots);
}
/* [P2] polynomial/mod.rs coset_fft_with_options: c_j *= shift^j, then fft. */
static void coset_fft_with_roots(uint64_t* v, unsigned log_n, uint64_t shift,
                                 const uint64_t* roots) {
  uint64_t n = 1ULL << log_n, s = 1;
  shift = canon(shift);
  for (uint64_t j = 0; j < n; j++) {
    v[j] = mul_(v[j], s);
    s = mul_(s, shift);
  }
  fft_classic(v, log_n, roots);
}
void orc_coset_fft(uint64_t* v, unsigned log_n, uint64_t shift) {
  uint64_t* roots = make_roots(log_n);
  coset_fft_with_roots(v, log_n, shift, roots);
  free(roots);
}
/* [P2] PolynomialCoeffs::lde = zero-pad to n<<rate_bits; lde_values() then takes
 * coset_fft_with_options(F::coset_shift(), Some(rate_bits), fft_root_table). */
void orc_lde(const uint64_t* coeffs, unsigned log_n, unsigned rate_bits, uint64_t* out) {
  uint64_t n = 1ULL << log_n, m = n << rate_bits;
  memcpy(out, coeffs, n * sizeof(uint64_t));
  memset(out + n, 0, (m - n) * sizeof(uint64_t));
  orc_coset_fft(out, log_n + rate_bits, COSET_SHIFT);
}

/* ------------------------------------------------------------------------------------------
 * Poseidon.  [P2] plonky2/src/hash/poseidon.rs (Poseidon trait: SPONGE_WIDTH 12, x^7,
 * HALF_N_FULL_ROUNDS 4, N_PARTIAL_ROUNDS 22) and hash/poseidon_goldilocks.rs (MDS_MATRIX_CIRC,
 * MDS_MATRIX_DIAG, ALL_ROUND_CONSTANTS).  Naive round form (constant layer, S-box layer, MDS
 * layer), which upstream's tests assert equal to its fast-partial-round form.
 *
 * ALL_ROUND_CONSTANTS is regenerated instead of transcribed: upstream documents it as 360 draws
 * of F::rand() from ChaCha8Rng::seed_from_u64(0) (rand 0.8 / rand_chacha 0.3).  SURVEY.md App. D.
 * The three plonky2 known-answer vectors in tests/golden/poseidon_kat.json check the result.
 * ---------------------------------------------------------------------------------------- */
static const uint64_t MDS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
static uint64_t RC[360];
static int rc_ready = 0;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define QR(a, b, c, d)                  \
  a += b; d ^= a; d = rotl32(d, 16);    \
  c += d; b ^= c; b = rotl32(b, 12);    \
  a += b; d ^= a; d = rotl32(d, 8);     \
  c += d; b ^= c; b = rotl32(b, 7);
static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
  uint32_t x[16];
  memcpy(x, s, sizeof x);
  for (int dr = 0; dr < 4; dr++) { /* 4 double rounds = ChaCha8 */
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
    QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
    QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}
static void gen_round_constants(void) {
  /* rand_core SeedableRng::seed_from_u64(0): PCG32 expansion of the seed into the key. */
  uint32_t key[8];
  uint64_t state = 0;
  for (int i = 0; i < 8; i++) {
    state = state * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
  uint32_t blk[16];
  uint64_t counter = 0;
  int pos = 16, k = 0;
  while (k < 360) {
    uint32_t w[2];
    for (int h = 0; h < 2; h++) {
      if (pos == 16) {
        chacha8_block(key, counter++, blk);
        pos = 0;
      }
      w[h] = blk[pos++];
    }
    uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    /* rand 0.8 UniformInt<u64>::sample_single(0, p): zone = (p << lz(p)) - 1 = p - 1. */
    u128 wide = (u128)v * P;
    if ((uint64_t)wide <= P - 1) RC[k++] = (uint64_t)(wide >> 64);
  }
  rc_ready = 1;
}
static void ensure_rc(void) {
  if (!rc_ready) {
#pragma omp critical(orc_rc)
    if (!rc_ready) gen_round_constants();
  }
}
void orc_poseidon_round_constants(uint64_t out[360]) {
  ensure_rc();
  memcpy(out, RC, sizeof RC);
}

/* Internally the state words are arbitrary u64 representatives (as upstream's GoldilocksField);
 * they are canonicalised when the permutation returns. */
static inline uint64_t mul_lazy(uint64_t a, uint64_t b) { return reduce128((u128)a * b); }
static inline uint64_t sbox7(uint64_t x) {
  uint64_t x2 = mul_lazy(x, x), x4 = mul_lazy(x2, x2), x3 = mul_lazy(x, x2);
  return mul_lazy(x3, x4);
}
/* [P2] Poseidon::mds_layer / mds_row_shf: out[r] = sum_i state[(i+r)%12]*CIRC[i] + state[r]*DIAG[r].
 * Computed on the 32-bit halves of each word (sums stay below 2^42), like upstream's
 * mds_row_shf does with u128 accumulators; `rc` = the next round's constants (or zeros). */
static inline void mds_layer_add_rc(uint64_t s[12], const uint64_t* rc) {
  static const uint32_t C32[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  uint32_t lo[24], hi[24];
  uint64_t al[12], ah[12];
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = (uint32_t)s[i];
    hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32);
    al[i] = ah[i] = 0;
  }
  for (int i = 0; i < 12; i++) /* 32x32->64 products: vectorises to vpmuludq */
    for (int r = 0; r < 12; r++) {
      al[r] += (uint64_t)lo[i + r] * C32[i];
      ah[r] += (uint64_t)hi[i + r] * C32[i];
    }
  al[0] += (uint64_t)lo[0] * (uint32_t)MDS_DIAG[0];
  ah[0] += (uint64_t)hi[0] * (uint32_t)MDS_DIAG[0];
  for (int r = 0; r < 12; r++) {
    u128 v = (u128)al[r] + ((u128)ah[r] << 32) + rc[r]; /* value = al + 2^32 * ah (+ rc) */
    s[r] = reduce128(v);
  }
}
static const uint64_t ZERO12[12] = {0};
static void poseidon_(uint64_t s[12]) {
  for (int i = 0; i < 12; i++) s[i] = add_(canon(s[i]), RC[i]);
  for (int r = 0; r < 30; r++) {
    if (r < 4 || r >= 26) {
      for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
    } else {
      s[0] = sbox7(s[0]);
    }
    mds_layer_add_rc(s, r < 29 ? RC + 12 * (r + 1) : ZERO12);
  }
  for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}
void orc_poseidon(uint64_t state[12]) {
  ensure_rc();
  for (int i = 0; i < 12; i++) state[i] = canon(state[i]);
  poseidon_(state);
}
/* [P2] hash/hashing.rs hash_n_to_m_no_pad with PoseidonPermutation (RATE 8): state starts at 0;
 * each chunk of <=8 inputs OVERWRITES state[0..len) (set_from_slice), then permute; squeeze
 * state[0..4). */
static void hash_no_pad_(const uint64_t* in, size_t len, uint64_t out[4]) {
  uint64_t s[12] = {0};
  for (size_t off = 0; off < len; off += 8) {
    size_t c = len - off < 8 ? len - off : 8;
    for (size_t i = 0; i < c; i++) s[i] = canon(in[off + i]);
    poseidon_(s);
  }
  memcpy(out, s, 4 * sizeof(uint64_t));
}
void orc_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]) {
  ensure_rc();
  hash_no_pad_(in, len, out);
}
/* [P2] plonk/config.rs Hasher::hash_or_noop: inputs that fit in a hash (<= 4 elements) are
 * copied (canonical, zero padded) instead of hashed. */
static void hash_or_noop_(const uint64_t* in, size_t len, uint64_t out[4]) {
  if (len <= 4) {
    for (size_t i = 0; i < 4; i++) out[i] = i < len ? canon(in[i]) : 0;
  } else {
    hash_no_pad_(in, len, out);
  }
}
void orc_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]) {
  ensure_rc();
  hash_or_noop_(in, len, out);
}
/* [P2] hash/hashing.rs compress (= PoseidonHash::two_to_one): perm(l || r || 0^4)[0..4). */
static void two_to_one_(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  uint64_t s[12] = {0};
  for (int i = 0; i < 4; i++) {
    s[i] = canon(l[i]);
    s[4 + i] = canon(r[i]);
  }
  poseidon_(s);
  memcpy(out, s, 4 * sizeof(uint64_t));
}
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
  ensure_rc();
  two_to_one_(l, r, out);
}

/* ------------------------------------------------------------------------------------------
 * SIMD evaluation of the same permutation on 4 (AVX2) or 8 (AVX-512) states at once: the speed
 * of the CPU arm (bench.py cpu_baseline / --impl reference); the scalar code above stays the
 * checker.  See poseidon_simd.inc.
 * ---------------------------------------------------------------------------------------- */
#define VSUF4_(name) name##_x4
#define VSUF8_(name) name##_x8
typedef uint64_t v4u __attribute__((vector_size(32)));
typedef uint64_t v8u __attribute__((vector_size(64)));
#define VW 4
#define VT v4u
#define VMU(a, b) ((v4u)_mm256_mul_epu32((__m256i)(a), (__m256i)(b)))
#define VSUF(name) VSUF4_(name)
#define VTARGET __attribute__((target("avx2")))
#include "poseidon_simd.inc"
#undef VW
#undef VT
#undef VMU
#undef VSUF
#undef VTARGET
#define VW 8
#define VT v8u
#define VMU(a, b) ((v8u)_mm512_mul_epu32((__m512i)(a), (__m512i)(b)))
#define VSUF(name) VSUF8_(name)
#define VTARGET __attribute__((target("avx512f,avx512dq,avx512bw,avx512vl")))
#include "poseidon_simd.inc"
#undef VW
#undef VT
#undef VMU
#undef VSUF
#undef VTARGET

static void fft_layer_x4(uint64_t* v, uint64_t n, uint64_t half_m, const uint64_t* tw) {
  fft_layer_impl_x4(v, n, half_m, tw);
}
static void fft_layer_x8(uint64_t* v, uint64_t n, uint64_t half_m, const uint64_t* tw) {
  fft_layer_impl_x8(v, n, half_m, tw);
}

static int g_simd = 0; /* 0 = widest the CPU supports, 1 = scalar (naive checker), 4, 8 */
void orc_set_simd(int width) { g_simd = width; }
int orc_get_simd(void) {
  int w = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") ? 8
          : __builtin_cpu_supports("avx2")                                          ? 4
                                                                                    : 1;
  if (g_simd == 1 || (g_simd == 4 && w >= 4) || (g_simd == 8 && w >= 8)) w = g_simd;
  return w;
}
/* count independent permutations (count x 12 words) through the SIMD path */
void orc_poseidon_batch(uint64_t* states, uint64_t count) {
  ensure_rc();
  const int w = orc_get_simd();
  uint64_t k = 0;
  if (w == 8)
    for (; k + 8 <= count; k += 8) permute_states_x8(states + 12 * k);
  if (w >= 4)
    for (; k + 4 <= count; k += 4) permute_states_x4(states + 12 * k);
  for (; k < count; k++) orc_poseidon(states + 12 * k);
}

/* ------------------------------------------------------------------------------------------
 * Merkle tree.  [P2] plonky2/src/hash/merkle_tree.rs
 * ---------------------------------------------------------------------------------------- */
static int log2_strict(uint64_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1ULL << l) < n) l++;
  return l;
}
/* [P2] fill_subtree, literally: digests_buf layout is
 *   left recursive output || left child digest || right child digest || right recursive output */
static void fill_subtree(uint64_t* digests_buf, uint64_t buf_len, const uint64_t* leaves,
                         uint64_t nleaves, uint32_t leaf_len, uint64_t out[4]) {
  if (buf_len == 0) {
    hash_or_noop_(leaves, leaf_len, out);
    return;
  }
  uint64_t half = buf_len / 2;
  uint64_t* left_buf = digests_buf;               /* [0, half-1) */
  uint64_t* left_mem = digests_buf + 4 * (half - 1);
  uint64_t* right_mem = digests_buf + 4 * half;
  uint64_t* right_buf = digests_buf + 4 * (half + 1); /* (half, buf_len) */
  uint64_t l[4], r[4];
  if (nleaves >= 4096) {
#pragma omp task shared(l)
    fill_subtree(left_buf, half - 1, leaves, nleaves / 2, leaf_len, l);
#pragma omp task shared(r)
    fill_subtree(right_buf, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len, r);
#pragma omp taskwait
  } else {
    fill_subtree(left_buf, half - 1, leaves, nleaves / 2, leaf_len, l);
    fill_subtree(right_buf, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len, r);
  }
  memcpy(left_mem, l, sizeof l);
  memcpy(right_mem, r, sizeof r);
  two_to_one_(l, r, out);
}
/* [P2] MerkleTree::new + fill_digests_buf. */
int orc_merkle_new(const uint64_t* leaves, uint64_t nleaves, uint32_t leaf_len, uint32_t cap_height,
                   uint64_t* digests, uint64_t* cap) {
  ensure_rc();
  int lg = log2_strict(nleaves);
  if (lg < 0 || (int)cap_height > lg) return -1;
  uint64_t ncap = 1ULL << cap_height;
  uint64_t num_digests = 2 * (nleaves - ncap);
  { /* SIMD path (level by level, same digests layout); the recursion below is the checker */
    const int w = orc_get_simd();
    if (w == 8 && nleaves >= 8) {
      merkle_new_x8(leaves, nleaves, leaf_len, (unsigned)lg - cap_height, digests, cap, orc_get_threads());
      return 0;
    }
    if (w >= 4 && nleaves >= 4) {
      merkle_new_x4(leaves, nleaves, leaf_len, (unsigned)lg - cap_height, digests, cap, orc_get_threads());
      return 0;
    }
  }
  if (num_digests == 0) { /* tree is all cap */
#pragma omp parallel for num_threads(orc_get_threads()) schedule(static)
    for (uint64_t k = 0; k < nleaves; k++) hash_or_noop_(leaves + k * leaf_len, leaf_len, cap + 4 * k);
    return 0;
  }
  uint64_t sub_digests = num_digests >> cap_height, sub_leaves = nleaves >> cap_height;
#pragma omp parallel num_threads(orc_get_threads())
#pragma omp single
  for (uint64_t s = 0; s < ncap; s++) {
#pragma omp task firstprivate(s)
    fill_subtree(digests + 4 * s * sub_digests, sub_digests, leaves + s * sub_leaves * leaf_len,
                 sub_leaves, leaf_len, cap + 4 * s);
  }
  return 0;
}
/* [P2] MerkleTree::prove. */
int orc_merkle_prove(const uint64_t* digests, uint64_t nleaves, uint32_t cap_height,
                     uint64_t leaf_index, uint64_t* siblings_out) {
  int lg = log2_strict(nleaves);
  if (lg < 0 || (int)cap_height > lg || leaf_index >= nleaves) return -1;
  uint32_t num_layers = (uint32_t)lg - cap_height;
  uint64_t num_digests = 2 * (nleaves - (1ULL << cap_height));
  uint64_t tree_index = leaf_index >> num_layers;
  uint64_t tree_len = num_digests >> cap_height;
  const uint64_t* digest_tree = digests + 4 * tree_len * tree_index;
  uint64_t pair_index = leaf_index & ((1ULL << num_layers) - 1);
  for (uint32_t i = 0; i < num_layers; i++) {
    uint64_t parity = pair_index & 1;
    pair_index >>= 1;
    uint64_t siblings_index = (pair_index << (i + 1)) + (1ULL << i) - 1;
    uint64_t sibling_index = 2 * siblings_index + (1 - parity);
    memcpy(siblings_out + 4 * i, digest_tree + 4 * sibling_index, 32);
  }
  return 0;
}
/* [P2] hash/merkle_proofs.rs verify_merkle_proof_to_cap. */
int orc_merkle_verify(const uint64_t* leaf, uint32_t leaf_len, uint64_t leaf_index,
                      const uint64_t* siblings, uint32_t nsiblings, const uint64_t* cap,
                      uint32_t cap_height) {
  ensure_rc();
  (void)cap_height;
  uint64_t cur[4], nxt[4];
  uint64_t index = leaf_index;
  hash_or_noop_(leaf, leaf_len, cur);
  for (uint32_t i = 0; i < nsiblings; i++) {
    uint64_t bit = index & 1;
    index >>= 1;
    if (bit) two_to_one_(siblings + 4 * i, cur, nxt);
    else two_to_one_(cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, sizeof cur);
  }
  return memcmp(cur, cap + 4 * index, 32) == 0 ? 0 : 1;
}

/* ------------------------------------------------------------------------------------------
 * PolynomialBatch.  [P2] plonky2/src/fri/oracle.rs from_values / from_coeffs / lde_values;
 * plonky2_util transpose + reverse_index_bits_in_place.  Reached from the reference at
 * /root/reference/src/vtfhe/ivc_based_vpbs.rs:275 (build) and :302/:333/:364 (prove).
 * ---------------------------------------------------------------------------------------- */
int orc_commit(const uint64_t* const* cols, uint32_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, int inputs_are_coeffs, const uint64_t* const* salt_cols,
               uint64_t* coeffs_out, uint64_t* lde_cols_out, uint64_t* leaves_out,
               uint64_t* digests_out, uint64_t* cap_out) {
  ensure_rc();
  if (!cols || ncols == 0 || log_n + rate_bits > 32 || cap_height > log_n + rate_bits) return -1;
  const uint64_t n = 1ULL << log_n, m = n << rate_bits;
  const unsigned log_m = log_n + rate_bits;
  const uint32_t salt = salt_cols ? 4 : 0, width = ncols + salt;
  int nt = orc_get_threads();
  uint64_t* roots_n = make_roots(log_n);
  uint64_t* roots_m = make_roots(log_m);
  const int fast = orc_get_simd() > 1;
  uint64_t** lt_n = fast ? make_layer_roots(log_n, roots_n) : NULL;
  uint64_t** lt_m = fast ? make_layer_roots(log_m, roots_m) : NULL;
  uint64_t* shift_pow = NULL; /* 7^j, shared by all columns */
  if (fast) {
    shift_pow = (uint64_t*)malloc(sizeof(uint64_t) * n);
    uint64_t sp = 1;
    for (uint64_t j = 0; j < n; j++) {
      shift_pow[j] = sp;
      sp = mul_(sp, COSET_SHIFT);
    }
  }
  uint64_t* coeffs = coeffs_out ? coeffs_out : (uint64_t*)malloc(sizeof(uint64_t) * ncols * n);
  uint64_t* lde = lde_cols_out ? lde_cols_out : (uint64_t*)malloc(sizeof(uint64_t) * (uint64_t)ncols * m);
  if (!coeffs || !lde) return -2;

  /* "IFFT": values.into_par_iter().map(|v| v.ifft()) */
#pragma omp parallel for num_threads(nt) schedule(dynamic)
  for (uint32_t c = 0; c < ncols; c++) {
    uint64_t* dst = coeffs + (uint64_t)c * n;
    for (uint64_t i = 0; i < n; i++) dst[i] = canon(cols[c][i]);
    if (!inputs_are_coeffs) {
      if (fast) {
        fft_fast(dst, log_n, lt_n, 0);
        ifft_finish(dst, log_n);
      } else {
        ifft_with_roots(dst, log_n, roots_n);
      }
    }
  }
  /* "FFT + blinding": polynomials.par_iter().map(|p| p.lde(r).coset_fft_with_options(7, ..)) */
#pragma omp parallel for num_threads(nt) schedule(dynamic)
  for (uint32_t c = 0; c < ncols; c++) {
    uint64_t* dst = lde + (uint64_t)c * m;
    memset(dst + n, 0, (m - n) * sizeof(uint64_t));
    if (fast) { /* c_j * 7^j on the n non-zero coefficients, then the transform minus its copy layers */
      const uint64_t* src = coeffs + (uint64_t)c * n;
      for (uint64_t j = 0; j < n; j++) dst[j] = mul_(src[j], shift_pow[j]);
      fft_fast(dst, log_m, lt_m, rate_bits);
    } else {
      memcpy(dst, coeffs + (uint64_t)c * n, n * sizeof(uint64_t));
      coset_fft_with_roots(dst, log_m, COSET_SHIFT, roots_m);
    }
  }
  /* "transpose LDEs" + reverse_index_bits_in_place: leaf k = natural row bitrev(k); the salt
   * columns are further entries of lde_values and are transposed/reordered with the rest. */
  if (leaves_out) {
#pragma omp parallel for num_threads(nt) schedule(static)
    for (uint64_t k = 0; k < m; k++) {
      uint64_t i = bitrev(k, log_m);
      uint64_t* row = leaves_out + k * width;
      for (uint32_t c = 0; c < ncols; c++) row[c] = lde[(uint64_t)c * m + i];
      for (uint32_t s = 0; s < salt; s++) row[ncols + s] = canon(salt_cols[s][i]);
    }
  }
  int rc = 0;
  /* "build Merkle tree" */
  if (leaves_out && cap_out)
    rc = orc_merkle_new(leaves_out, m, width, cap_height, digests_out, cap_out);
  if (!coeffs_out) free(coeffs);
  if (!lde_cols_out) free(lde);
  free(roots_n);
  free(roots_m);
  free(lt_n);
  free(lt_m);
  free(shift_pow);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * Quadratic extension and polynomial evaluation.  [P2] plonky2_field/src/extension/quadratic.rs
 * (W = 7), polynomial/mod.rs PolynomialCoeffs::eval (Horner from the highest coefficient).
 * ---------------------------------------------------------------------------------------- */
#define EXT_W 7ULL
void orc_eval_ext2(const uint64_t* const* cols, uint32_t ncols, uint64_t n, const uint64_t x[2],
                   uint64_t* out) {
  const uint64_t x0 = canon(x[0]), x1 = canon(x[1]);
#pragma omp parallel for num_threads(orc_get_threads()) schedule(dynamic)
  for (uint32_t c = 0; c < ncols; c++) {
    uint64_t a0 = 0, a1 = 0; /* acc = a0 + a1 X */
    for (uint64_t j = n; j-- > 0;) {
      /* acc = acc * x + coeff:  (a0 + a1 X)(x0 + x1 X) = a0 x0 + W a1 x1 + (a0 x1 + a1 x0) X */
      uint64_t n0 = add_(mul_(a0, x0), mul_(EXT_W, mul_(a1, x1)));
      uint64_t n1 = add_(mul_(a0, x1), mul_(a1, x0));
      a0 = add_(n0, canon(cols[c][j]));
      a1 = n1;
    }
    out[2 * c] = a0;
    out[2 * c + 1] = a1;
  }
}

/* ------------------------------------------------------------------------------------------
 * FRI commit phase, one layer.  [P2] plonky2/src/fri/prover.rs fri_committed_trees.
 * ---------------------------------------------------------------------------------------- */
int orc_fri_layer_commit(const uint64_t* values_ext, uint64_t len, uint32_t arity_bits,
                         uint32_t cap_height, uint64_t* leaves_out, uint64_t* digests_out,
                         uint64_t* cap_out) {
  int lg = log2_strict(len);
  if (lg < 0 || (int)arity_bits > lg) return -1;
  /* reverse_index_bits_in_place, then chunk: leaf j, slot i = values[bitrev(j * arity + i)] */
  for (uint64_t k = 0; k < len; k++) {
    uint64_t src = bitrev(k, (unsigned)lg);
    leaves_out[2 * k] = canon(values_ext[2 * src]);
    leaves_out[2 * k + 1] = canon(values_ext[2 * src + 1]);
  }
  return orc_merkle_new(leaves_out, len >> arity_bits, 2u << arity_bits, cap_height, digests_out, cap_out);
}
void orc_fri_fold(const uint64_t* coeffs_ext, uint64_t len, uint32_t arity_bits,
                  const uint64_t beta[2], uint64_t shift_next, uint64_t* coeffs_out,
                  uint64_t* values_out) {
  const uint64_t arity = 1ULL << arity_bits, out_len = len >> arity_bits;
  const uint64_t b0 = canon(beta[0]), b1 = canon(beta[1]);
  for (uint64_t j = 0; j < out_len; j++) {
    uint64_t a0 = 0, a1 = 0; /* Horner: acc = acc * beta + chunk[i], i from high to low */
    for (uint64_t i = arity; i-- > 0;) {
      uint64_t n0 = add_(mul_(a0, b0), mul_(EXT_W, mul_(a1, b1)));
      uint64_t n1 = add_(mul_(a0, b1), mul_(a1, b0));
      a0 = add_(n0, canon(coeffs_ext[2 * (j * arity + i)]));
      a1 = add_(n1, canon(coeffs_ext[2 * (j * arity + i) + 1]));
    }
    coeffs_out[2 * j] = a0;
    coeffs_out[2 * j + 1] = a1;
  }
  /* coset_fft over the extension = the base-field transform applied to each component */
  int lg = log2_strict(out_len);
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * out_len);
  for (int comp = 0; comp < 2; comp++) {
    for (uint64_t j = 0; j < out_len; j++) tmp[j] = coeffs_out[2 * j + comp];
    orc_coset_fft(tmp, (unsigned)lg, shift_next);
    for (uint64_t j = 0; j < out_len; j++) values_out[2 * j + comp] = tmp[j];
  }
  free(tmp);
}

/* ---- [P2] fri/oracle.rs PolynomialBatch::prove_openings, up to final_poly ------------------------ */
/* ReducingFactor (util/reducing.rs): reduce_polys_base multiplies the j-th polynomial of the call by
 * base^j (powers restart at 1 in every call) and counts the polynomials; shift_poly multiplies by
 * base^count and resets the count.  divide_by_linear (polynomial/division.rs): bs = scan from the top
 * of acc = acc * z + c; drop bs of the constant term; quotient = the rest, lowest first. */
int orc_fri_final_poly(const uint64_t* const* polys, const uint32_t* batch_sizes, uint32_t nbatches,
                       uint64_t n, const uint64_t* points, const uint64_t alpha[2], uint64_t* out) {
  if (!polys || !batch_sizes || !points || !alpha || !out || n == 0 || nbatches == 0) return -1;
  const uint64_t a0 = canon(alpha[0]), a1 = canon(alpha[1]);
  uint64_t* comp = (uint64_t*)malloc(sizeof(uint64_t) * 2 * n);
  if (!comp) return -1;
  memset(out, 0, sizeof(uint64_t) * 2 * n); /* final_poly = PolynomialCoeffs::empty() */
  size_t off = 0;
  for (uint32_t b = 0; b < nbatches; b++) {
    const uint32_t len = batch_sizes[b];
    if (len == 0) {
      free(comp);
      return -1;
    }
    /* composition_poly = alpha.reduce_polys_base(polys_coeff) */
    memset(comp, 0, sizeof(uint64_t) * 2 * n);
    uint64_t p0 = 1, p1 = 0; /* alpha^j */
    for (uint32_t j = 0; j < len; j++) {
      const uint64_t* f = polys[off + j];
      for (uint64_t i = 0; i < n; i++) {
        const uint64_t c = canon(f[i]);
        comp[2 * i] = add_(comp[2 * i], mul_(c, p0));
        comp[2 * i + 1] = add_(comp[2 * i + 1], mul_(c, p1));
      }
      const uint64_t n0 = add_(mul_(p0, a0), mul_(EXT_W, mul_(p1, a1)));
      const uint64_t n1 = add_(mul_(p0, a1), mul_(p1, a0));
      p0 = n0;
      p1 = n1;
    }
    /* after the loop (p0, p1) = alpha^len = the factor shift_poly applies */
    /* quotient = composition_poly.divide_by_linear(point); quotient.coeffs.push(ZERO) */
    const uint64_t z0 = canon(points[2 * b]), z1 = canon(points[2 * b + 1]);
    uint64_t acc0 = 0, acc1 = 0; /* b_{i+1} while visiting i from the top */
    for (uint64_t i = n; i-- > 0;) {
      const uint64_t q0 = acc0, q1 = acc1; /* quotient coefficient i (q_{n-1} = 0: the pushed zero) */
      const uint64_t m0 = add_(mul_(acc0, z0), mul_(EXT_W, mul_(acc1, z1)));
      const uint64_t m1 = add_(mul_(acc0, z1), mul_(acc1, z0));
      acc0 = add_(m0, comp[2 * i]);
      acc1 = add_(m1, comp[2 * i + 1]);
      /* alpha.shift_poly(&mut final_poly); final_poly += quotient */
      const uint64_t f0 = out[2 * i], f1 = out[2 * i + 1];
      out[2 * i] = add_(add_(mul_(f0, p0), mul_(EXT_W, mul_(f1, p1))), q0);
      out[2 * i + 1] = add_(add_(mul_(f0, p1), mul_(f1, p0)), q1);
    }
    off += len;
  }
  free(comp);
  return 0;
}

/* ---- [P2] plonk/prover.rs wires_permutation_partial_products_and_zs -------------------------- */
int orc_zs_partial_products(const uint64_t* const* wires, const uint64_t* const* sigmas,
                            const uint64_t* k_is, uint32_t num_routed, uint32_t log_n,
                            uint32_t max_degree, uint64_t beta, uint64_t gamma, uint64_t* out) {
  if (!wires || !sigmas || !k_is || !out || num_routed == 0 || max_degree < 2 || log_n > 30) return -1;
  const uint64_t n = 1ULL << log_n;
  const uint32_t K = (num_routed + max_degree - 1) / max_degree; /* chunks = partial products + 1 */
  beta = canon(beta);
  gamma = canon(gamma);
  /* all_quotient_chunk_products[i][k] (par_iter over the subgroup upstream) */
  uint64_t* chunks = malloc((size_t)n * K * sizeof(uint64_t));
  uint64_t* subgroup = malloc((size_t)n * sizeof(uint64_t));
  if (!chunks || !subgroup) {
    free(chunks);
    free(subgroup);
    return -1;
  }
  const uint64_t w = orc_primitive_root_of_unity(log_n);
  subgroup[0] = 1;
  for (uint64_t i = 1; i < n; i++) subgroup[i] = mul_(subgroup[i - 1], w);
  int zero_den = 0;
#pragma omp parallel for schedule(static) num_threads(orc_get_threads()) reduction(| : zero_den)
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t x = subgroup[i];
    for (uint32_t k = 0; k < K; k++) {
      uint64_t num = 1, den = 1;
      for (uint32_t j = k * max_degree; j < (k + 1) * max_degree && j < num_routed; j++) {
        const uint64_t wv = canon(wires[j][i]);
        const uint64_t s_id = mul_(canon(k_is[j]), x);
        num = mul_(num, add_(add_(wv, mul_(beta, s_id)), gamma));
        den = mul_(den, add_(add_(wv, mul_(beta, canon(sigmas[j][i]))), gamma));
      }
      /* prod(num_j / den_j) = prod(num_j) / prod(den_j): field inverses are unique, so this equals
       * upstream's batch_multiplicative_inverse + per-wire multiply bit for bit */
      if (den == 0) zero_den |= 1;
      chunks[i * K + k] = mul_(num, orc_gl_inv(den));
    }
  }
  if (!zero_den) {
    uint64_t z_x = 1; /* [P2] "let mut z_x = F::ONE" */
    for (uint64_t i = 0; i < n; i++) {
      uint64_t acc = z_x;
      out[i] = z_x; /* the last partial product is Z(gx); Z(x) takes its place in the row */
      for (uint32_t k = 0; k < K; k++) {
        acc = mul_(acc, chunks[i * K + k]);
        if (k + 1 < K) out[(uint64_t)(1 + k) * n + i] = acc;
      }
      z_x = acc;
    }
  }
  free(chunks);
  free(subgroup);
  return zero_den ? -2 : 0;
}

/* ---- [P2] plonk/prover.rs compute_quotient_polys + plonk/vanishing_poly.rs
 * eval_vanishing_poly_base_batch, the terms that do not depend on the gate set ------------------------
 * (reached from prove(), /root/reference/src/vtfhe/ivc_based_vpbs.rs:302, :333, :364).
 * Quotient domain: lde_q = n << qdb points x_i = 7 w_q^i (natural order), qdb =
 * log2_ceil(quotient_degree_factor).  Per point and challenge c, with K = ceil(num_routed / max_degree)
 * chunks (max_degree = quotient_degree_factor):
 *   Z(1) = 1 term:          L_0(x) (Z_c(x) - 1),  L_0(x) = (x^n - 1) / (n (x - 1))     [ZeroPolyOnCoset::eval_l_0]
 *   partial-product checks: accs = [Z_c(x), pp_c,0(x) .. pp_c,K-2(x), Z_c(g x)] (g = w_n, i.e. the point
 *                           i + 2^qdb); check_t = accs[t] prod_{j in chunk t} (w_j + beta_c k_j x + gamma_c)
 *                                              - accs[t+1] prod_{j in chunk t} (w_j + beta_c sigma_j(x) + gamma_c)
 *   vanishing terms, in upstream's order: Z(1) terms of all challenges, then the checks of challenge
 *   0, 1, .., then (lookups: none in this reference) the gate constraints.  reduce_with_powers_multi:
 *   res_c = sum_j term_j alpha_c^j.  The gate constraints enter through gate_terms[c][i] =
 *   sum_j alpha_c^j gate_constraint_j(x_i) (NULL: none), shifted by alpha_c^(nc + nc K).
 *   quotient value = res_c / (x^n - 1); then per challenge coset_ifft(7) and chunks of n coefficients.
 * Inputs are COEFFICIENT columns of length n (the batches' `polynomials`): wires[j] (j < num_routed),
 * sigmas[j], zs_pp[..] in commit order (Z_0 .. Z_{nc-1}, then the K - 1 partial products of challenge 0,
 * of challenge 1, ..).  out: nc * 2^qdb columns of n coefficients (challenge-major, chunk-minor). */
int orc_quotient_polys(const uint64_t* const* wires, const uint64_t* const* sigmas,
                       const uint64_t* const* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                       uint32_t log_n, uint32_t max_degree, uint32_t qdb, const uint64_t* betas,
                       const uint64_t* gammas, const uint64_t* alphas, uint32_t nc,
                       const uint64_t* const* gate_terms, uint64_t* out) {
  if (!wires || !sigmas || !zs_pp || !k_is || !betas || !gammas || !alphas || !out || num_routed == 0 ||
      max_degree < 2 || nc == 0 || log_n + qdb > 30)
    return -1;
  const uint64_t n = 1ULL << log_n, q = n << qdb, next_step = 1ULL << qdb;
  const uint32_t K = (num_routed + max_degree - 1) / max_degree;
  const uint32_t nzp = nc * K; /* columns of the Z / partial-products batch */
  const uint32_t nterms = nc + nc * K;
  uint64_t** lw = malloc(sizeof(uint64_t*) * (2 * num_routed + nzp));
  if (!lw) return -1;
#pragma omp parallel for schedule(dynamic) num_threads(orc_get_threads())
  for (uint32_t j = 0; j < 2 * num_routed + nzp; j++) {
    lw[j] = malloc(q * sizeof(uint64_t));
    const uint64_t* src = j < num_routed ? wires[j] : j < 2 * num_routed ? sigmas[j - num_routed] : zs_pp[j - 2 * num_routed];
    orc_lde(src, log_n, qdb, lw[j]);
  }
  uint64_t** ls = lw + num_routed;
  uint64_t** lz = lw + 2 * num_routed;
  /* ZeroPolyOnCoset: x^n - 1 takes 2^qdb values on the coset */
  uint64_t zh[1 << 10], zh_inv[1 << 10];
  const uint64_t g_pow_n = orc_gl_pow(7, n), wr = orc_primitive_root_of_unity(qdb);
  for (uint64_t k = 0, x = 1; k < next_step; k++, x = mul_(x, wr)) {
    zh[k] = sub_(mul_(g_pow_n, x), 1);
    zh_inv[k] = orc_gl_inv(zh[k]);
  }
  uint64_t* vals = malloc((size_t)nc * q * sizeof(uint64_t));
  uint64_t* xs = malloc(q * sizeof(uint64_t));
  const uint64_t wq = orc_primitive_root_of_unity(log_n + qdb);
  xs[0] = 7;
  for (uint64_t i = 1; i < q; i++) xs[i] = mul_(xs[i - 1], wq);
#pragma omp parallel for schedule(static) num_threads(orc_get_threads())
  for (uint64_t i = 0; i < q; i++) {
    const uint64_t x = xs[i], i_next = (i + next_step) % q;
    uint64_t terms[2 + 2 * 64];
    const uint64_t l0 = mul_(zh[i % next_step], orc_gl_inv(mul_(canon(n), sub_(x, 1))));
    for (uint32_t c = 0; c < nc; c++) terms[c] = mul_(l0, sub_(lz[c][i], 1));
    for (uint32_t c = 0; c < nc; c++) {
      const uint64_t beta = canon(betas[c]), gamma = canon(gammas[c]);
      for (uint32_t t = 0; t < K; t++) {
        uint64_t num = 1, den = 1;
        for (uint32_t j = t * max_degree; j < (t + 1) * max_degree && j < num_routed; j++) {
          const uint64_t wv = lw[j][i];
          num = mul_(num, add_(add_(wv, mul_(beta, mul_(canon(k_is[j]), x))), gamma));
          den = mul_(den, add_(add_(wv, mul_(beta, ls[j][i])), gamma));
        }
        const uint64_t prev = t == 0 ? lz[c][i] : lz[nc + c * (K - 1) + (t - 1)][i];
        const uint64_t next = t == K - 1 ? lz[c][i_next] : lz[nc + c * (K - 1) + t][i];
        terms[nc + c * K + t] = sub_(mul_(prev, num), mul_(next, den));
      }
    }
    for (uint32_t c = 0; c < nc; c++) {
      const uint64_t alpha = canon(alphas[c]);
      uint64_t acc = gate_terms && gate_terms[c] ? canon(gate_terms[c][i]) : 0;
      for (uint32_t j = nterms; j-- > 0;) acc = add_(mul_(acc, alpha), terms[j]);
      vals[(uint64_t)c * q + i] = mul_(acc, zh_inv[i % next_step]);
    }
  }
  /* coset_ifft(7): ifft, then coefficient j times 7^-j; chunks of n are contiguous */
  const uint64_t inv7 = orc_gl_inv(7);
  for (uint32_t c = 0; c < nc; c++) {
    uint64_t* v = vals + (uint64_t)c * q;
    orc_ifft(v, log_n + qdb);
    uint64_t s = 1;
    for (uint64_t j = 0; j < q; j++, s = mul_(s, inv7)) out[(uint64_t)c * q + j] = mul_(v[j], s);
  }
  for (uint32_t j = 0; j < 2 * num_routed + nzp; j++) free(lw[j]);
  free(lw);
  free(vals);
  free(xs);
  return 0;
}

/* ---- [P2] plonk/vanishing_poly.rs evaluate_gate_constraints_base_batch as a program ------------------
 * The gate constraints of a circuit given as straight-line code (instruction format: include/
 * vpbs_commit.h, vpbs_gate_program_upload), interpreted at every point x_i = 7 w_q^i of the quotient
 * domain over the LDE values of ALL wire columns and ALL columns of the constants/sigmas batch
 * (coefficient columns in).  out[c][i] = sum_j alpha_c^j sum_gates filter_g c_{g,j}: the gate_terms
 * argument of orc_quotient_polys. */
int orc_gate_program_eval(const uint64_t* code, uint32_t ncode, const uint64_t* imms, uint32_t nimm,
                          uint32_t nregs, uint32_t num_constraints, const uint64_t* const* wires,
                          uint32_t nwires, const uint64_t* const* cs, uint32_t ncs, uint32_t log_n,
                          uint32_t qdb, const uint64_t pih[4], const uint64_t* alphas, uint32_t nc,
                          uint64_t* const* out) {
  if ((!code && ncode) || !wires || !cs || !alphas || !out || nregs == 0 || nc == 0 || nc > 4 || log_n + qdb > 30)
    return -1;
  const uint64_t n = 1ULL << log_n, q = n << qdb;
  uint64_t** lw = malloc(sizeof(uint64_t*) * (nwires + ncs));
#pragma omp parallel for schedule(dynamic) num_threads(orc_get_threads())
  for (uint32_t j = 0; j < nwires + ncs; j++) {
    lw[j] = malloc(q * sizeof(uint64_t));
    orc_lde(j < nwires ? wires[j] : cs[j - nwires], log_n, qdb, lw[j]);
  }
  uint64_t* apow = malloc((size_t)nc * num_constraints * sizeof(uint64_t));
  for (uint32_t c = 0; c < nc; c++) {
    uint64_t pw = 1;
    for (uint32_t j = 0; j < num_constraints; j++, pw = mul_(pw, canon(alphas[c]))) apow[(size_t)c * num_constraints + j] = pw;
  }
  int bad = 0;
#pragma omp parallel for schedule(static) num_threads(orc_get_threads()) reduction(| : bad)
  for (uint64_t i = 0; i < q; i++) {
    uint64_t regs[256];
    uint64_t total[4] = {0, 0, 0, 0}, gacc[4] = {0, 0, 0, 0};
    for (uint32_t pc = 0; pc < ncode; pc++) {
      const uint64_t ins = code[pc];
      const unsigned op = ins & 0xff, dst = (ins >> 8) & 0xff;
      const unsigned kind[2] = {(unsigned)((ins >> 16) & 0xf), (unsigned)((ins >> 20) & 0xf)};
      const unsigned idx[2] = {(unsigned)((ins >> 24) & 0xffff), (unsigned)((ins >> 40) & 0xffff)};
      uint64_t v[2] = {0, 0};
      for (int o = 0; o < ((op <= 2 || op == 5) ? 2 : 1); o++) {
        switch (kind[o]) {
          case 0: v[o] = idx[o] < 256 ? regs[idx[o]] : (bad |= 1, 0); break;
          case 1: v[o] = idx[o] < nwires ? lw[idx[o]][i] : (bad |= 1, 0); break;
          case 2: v[o] = idx[o] < ncs ? lw[nwires + idx[o]][i] : (bad |= 1, 0); break;
          case 3: v[o] = idx[o] < nimm ? canon(imms[idx[o]]) : (bad |= 1, 0); break;
          case 4: v[o] = pih ? canon(pih[idx[o] & 3]) : 0; break;
          default: bad |= 1;
        }
      }
      if (op == 0) regs[dst] = add_(v[0], v[1]);
      else if (op == 1) regs[dst] = sub_(v[0], v[1]);
      else if (op == 2) regs[dst] = mul_(v[0], v[1]);
      else if (op == 5) regs[dst] = add_(regs[dst], mul_(v[0], v[1]));
      else if (op == 3) {
        if (idx[1] >= num_constraints) { bad |= 1; continue; }
        for (uint32_t c = 0; c < nc; c++) gacc[c] = add_(gacc[c], mul_(v[0], apow[(size_t)c * num_constraints + idx[1]]));
      } else if (op == 4) {
        for (uint32_t c = 0; c < nc; c++) {
          total[c] = add_(total[c], mul_(gacc[c], v[0]));
          gacc[c] = 0;
        }
      } else bad |= 1;
    }
    for (uint32_t c = 0; c < nc; c++) out[c][i] = total[c];
  }
  for (uint32_t j = 0; j < nwires + ncs; j++) free(lw[j]);
  free(lw);
  free(apow);
  return bad ? -1 : 0;
}
