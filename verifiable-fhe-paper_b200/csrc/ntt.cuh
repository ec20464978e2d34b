// ntt.cuh — batched Goldilocks NTT passes for the commit path (sm_100a).
//
// Replaces, for F = GoldilocksField:
//   [P2] plonky2_field 0.2.0 src/fft.rs            fft_dispatch / fft_classic / ifft_with_options
//   [P2] plonky2_field 0.2.0 src/polynomial/mod.rs PolynomialCoeffs::{lde, coset_fft_with_options}
//   [P2] plonky2_util 0.2.0  src/lib.rs            transpose + reverse_index_bits_in_place
// as reached from PolynomialBatch::from_values / from_coeffs ([P2] plonky2/src/fri/oracle.rs),
// i.e. from prove()/build() at /root/reference/src/vtfhe/ivc_based_vpbs.rs:275,302,333,364.
//
// Algorithm: decimation-in-frequency, split into passes of up to 8 layers.  A pass of s layers
// on blocks of size B is a batch of 2^s-point DFTs over stride B/2^s done in shared memory
// (bit-reversed in-place output), followed by one twiddle w_B^(low*k) per element (four-step
// form), so global memory is touched once per pass and the only global twiddle traffic is one
// table read per element.  After all passes position `pos` holds X[bitrev(pos)]:
//   * the forward LDE wants exactly that order (plonky2 bit-reverses the leaves), so its last pass
//     writes rows of the row-major leaf matrix directly (transpose fused, 128-byte row segments);
//   * the inverse transform / plain fft undo it in the last pass's store (natural order out).
// The LDE never materialises the zero padding: leaf block b (n rows) is the size-n transform of
// c_j * (7 w_m^bitrev(b))^j (what fft_classic's zero_factor skipping amounts to).
#pragma once
#include "gl64.cuh"

namespace ntt {

using gl::u32;
using gl::u64;

constexpr int THREADS = 256;
constexpr unsigned LOG_TILE = 12;  // 4096 elements (32 KB) per CTA
constexpr unsigned MAX_PASS_BITS = 8;

struct Roots {
  const u64* w;    // w[t] = omega_N^t, t < N/2
  unsigned log_N;  // >= 1
};

__device__ __forceinline__ unsigned brev(unsigned x, unsigned bits) {
  return bits ? (__brev(x) >> (32 - bits)) : 0u;
}
// omega_{2^lg}^(+-e), 0 <= e < 2^lg, lg <= log_N.
template <bool INVERSE>
__device__ __forceinline__ u64 root_of(const Roots& R, unsigned lg, u64 e) {
  u64 idx = e << (R.log_N - lg);
  const u64 N = 1ULL << R.log_N, half = N >> 1;
  if (INVERSE && idx) idx = N - idx;
  return idx < half ? __ldg(R.w + idx) : gl::P - __ldg(R.w + (idx - half));  // roots are never 0
}

// s radix-2 DIF layers over sm[q * pitch + t], q < 2^s, t < 2^log_T; tw[e] = omega_{2^s}^(+-e).
__device__ __forceinline__ void dif_layers(u64* sm, const u64* tw, unsigned s, unsigned log_T,
                                           unsigned pitch) {
  const unsigned half_elems = (1u << s >> 1) << log_T;
  for (unsigned l = 0; l < s; l++) {
    const unsigned log_dd = s - 1 - l, dd = 1u << log_dd;
    for (unsigned b = threadIdx.x; b < half_elems; b += THREADS) {
      const unsigned t = b & ((1u << log_T) - 1), pi = b >> log_T;
      const unsigned j = pi & (dd - 1);
      const unsigned q_lo = ((pi >> log_dd) << (log_dd + 1)) + j;
      u64* pa = sm + q_lo * pitch + t;
      u64* pb = pa + dd * pitch;
      const u64 a = *pa, c = *pb;
      *pa = gl::add(a, c);
      *pb = gl::mul(gl::sub(a, c), tw[j << l]);
    }
    __syncthreads();
  }
}

// ---- pass over a strided sub-transform (every pass but the last) ------------------------------
// Column-major data; blocks of size 2^log_B; 2^s-point DFT over stride sigma = 2^(log_B - s);
// a CTA takes 2^log_T consecutive `low` offsets (contiguous in memory).  blockIdx.y = column.
template <bool INVERSE>
__global__ void __launch_bounds__(THREADS)
pass_strided(const u64* __restrict__ src, u64 src_col_stride, u64* __restrict__ dst,
             u64 dst_col_stride, unsigned log_B, unsigned s, unsigned log_T,
             const u64* __restrict__ in_scale, Roots R) {
  extern __shared__ u64 sm[];
  const unsigned S = 1u << s, T = 1u << log_T;
  const unsigned log_sigma = log_B - s;
  u64* tw = sm + (S << log_T);
  const unsigned tiles_per_block_log = log_sigma - log_T;
  const u64 blk = blockIdx.x >> tiles_per_block_log;
  const u64 low0 = (u64)(blockIdx.x & ((1u << tiles_per_block_log) - 1)) << log_T;
  const u64 base = blk << log_B;
  src += (u64)blockIdx.y * src_col_stride;
  dst += (u64)blockIdx.y * dst_col_stride;

  for (unsigned e = threadIdx.x; e < S / 2; e += THREADS) tw[e] = root_of<INVERSE>(R, s, e);
  for (unsigned idx = threadIdx.x; idx < (S << log_T); idx += THREADS) {
    const unsigned t = idx & (T - 1), q = idx >> log_T;
    const u64 pos = base + ((u64)q << log_sigma) + low0 + t;
    u64 x = gl::canon(__ldg(src + pos));
    if (in_scale) x = gl::mul(x, __ldg(in_scale + pos));
    sm[idx] = x;
  }
  __syncthreads();
  dif_layers(sm, tw, s, log_T, T);
  for (unsigned idx = threadIdx.x; idx < (S << log_T); idx += THREADS) {
    const unsigned t = idx & (T - 1), q = idx >> log_T;
    const u64 pos = base + ((u64)q << log_sigma) + low0 + t;
    const u64 e = (low0 + t) * (u64)brev(q, s);  // < 2^log_B
    u64 x = sm[idx];
    if (e) x = gl::mul(x, root_of<INVERSE>(R, log_B, e));
    dst[pos] = x;
  }
}

// ---- last pass: contiguous 2^s-point blocks ------------------------------------------------------
enum StoreMode { STORE_LEAF = 0, STORE_NATURAL = 1 };
// STORE_LEAF   : lanes = 2^log_T consecutive columns (blockIdx.y), block = blockIdx.x;
//                dst[(row0 + pos) * dst_stride + col]  (row-major leaf matrix; pos = leaf index
//                inside this LDE block: transpose + reverse_index_bits fused).
// STORE_NATURAL: lanes = 2^log_T blocks whose bit-reversed ids are consecutive, column =
//                blockIdx.y; dst[col * dst_stride + bitrev(pos)]  (natural order, column-major).
template <bool INVERSE, int MODE>
__global__ void __launch_bounds__(THREADS)
pass_final(const u64* __restrict__ src, u64 src_col_stride, unsigned ncols,
           u64* __restrict__ dst, u64 dst_stride, u64 row0, unsigned log_n, unsigned s,
           unsigned log_T, const u64* __restrict__ in_scale, u64 out_scale, Roots R) {
  extern __shared__ u64 sm[];
  const unsigned S = 1u << s, T = 1u << log_T, pitch = T + 1;
  u64* tw = sm + S * pitch;
  const unsigned log_nb = log_n - s;  // blocks per column
  for (unsigned e = threadIdx.x; e < S / 2; e += THREADS) tw[e] = root_of<INVERSE>(R, s, e);

  for (unsigned idx = threadIdx.x; idx < (S << log_T); idx += THREADS) {
    const unsigned q = idx & (S - 1), lane = idx >> s;
    u64 x = 0;
    if (MODE == STORE_LEAF) {
      const unsigned col = blockIdx.y * T + lane;
      if (col < ncols) {
        const u64 pos = ((u64)blockIdx.x << s) + q;
        x = gl::canon(__ldg(src + (u64)col * src_col_stride + pos));
        if (in_scale) x = gl::mul(x, __ldg(in_scale + pos));
      }
    } else {
      const u64 blk = brev(blockIdx.x * T + lane, log_nb);
      const u64 pos = (blk << s) + q;
      x = gl::canon(__ldg(src + (u64)blockIdx.y * src_col_stride + pos));
      if (in_scale) x = gl::mul(x, __ldg(in_scale + pos));
    }
    sm[q * pitch + lane] = x;
  }
  __syncthreads();
  dif_layers(sm, tw, s, log_T, pitch);
  for (unsigned idx = threadIdx.x; idx < (S << log_T); idx += THREADS) {
    const unsigned lane = idx & (T - 1), q = idx >> log_T;
    u64 x = sm[q * pitch + lane];
    if (out_scale != 1) x = gl::mul(x, out_scale);
    if (MODE == STORE_LEAF) {
      const unsigned col = blockIdx.y * T + lane;
      const u64 pos = ((u64)blockIdx.x << s) + q;
      if (col < ncols) dst[(row0 + pos) * dst_stride + col] = x;
    } else {
      const u64 nat = ((u64)brev(q, s) << log_nb) + (u64)blockIdx.x * T + lane;
      dst[(u64)blockIdx.y * dst_stride + nat] = x;
    }
  }
}

// ---- radix-16 register kernels for 256-point passes (s = 8) --------------------------------------
// The 2^16-row commits (the microbench and every N=1024 step commit) are two passes of 8 layers.
// Here each thread keeps 16 elements in registers and runs 4 layers on them, the tile is exchanged
// through shared memory ONCE, and 4 more layers run in registers: 1 barrier and 1 shared-memory
// round trip per pass instead of 8, no per-butterfly index arithmetic, and the twiddles that are 1
// (15 of the 64 butterflies per thread) cost nothing.  Tile = 256 points x 16 lanes, 256 threads.

// 4 DIF layers over x[j], j = 0..15 (distance 8, 4, 2, 1 in j).  Butterfly (j, j + dd) of layer l
// multiplies its difference by tw[((j mod dd) * jstride + off) << l]; tw[0] = 1 is skipped when the
// exponent is known at compile time to be 0 (off == 0 && j mod dd == 0).
// Values inside the register layers are arbitrary u64 representatives ("lazy"): add_lazy / sub_lazy
// need only their SECOND operand canonical, so each butterfly canonicalises one input (4 slots)
// instead of paying for canonical add, sub and mul results (saves ~11 slots per butterfly).
template <bool HAS_OFF>
__device__ __forceinline__ void dif16(u64 (&x)[16], const u64* __restrict__ tw, unsigned jstride,
                                      unsigned off, unsigned l0) {
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const int dd = 8 >> l;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if ((j & dd) == 0) {
        const u64 a = x[j], c = gl::canon(x[j + dd]);
        // the _fma forms materialise their carries on the FMA pipe (IMAD.X instead of SEL): these
        // passes are bound by the ALU pipe (65-69 % against 36 % FMA-heavy), LDE 1.05 -> 1.00 ms
        x[j] = gl::add_lazy_fma(a, c);
        const u64 d = gl::sub_lazy(a, c);
        const int jm = j & (dd - 1);
        if (!HAS_OFF && jm == 0) x[j + dd] = d;
        else x[j + dd] = gl::mul_lazy_fma(d, tw[((unsigned)jm * jstride + off) << (l0 + l)]);
      }
    }
  }
}

// A 256-point DIF on a 256 x 16 tile is two of these with one exchange through shared memory in
// between: layers 0..3 on x[j] = element (q = 16 j + q_lo, lane) with dif16<true>(x, tw, 16, q_lo, 0),
// then layers 4..7 on x[j] = element (q = 16 q_hi + j, lane) with dif16<false>(x, tw, 1, 0, 4);
// afterwards position q holds frequency bitrev8(q).  tw[e] = omega_256^(+-e), 128 words.

// Two refinements of the four-step scheme used when a 2^16-point transform is two 256-point passes
// (template / run-time switches of the kernels below):
//  * OUT_TW = false leaves out the four-step twiddle w_B^(low * brev(q)) at the first pass's store:
//    the final pass applies it at its load (in_tw_log_B), where one CTA needs only 256 distinct
//    twiddles shared by its 16 columns instead of 4096 scattered table reads per CTA;
//  * the input scaling is split the same way: in_scale[pos] = g^pos with pos = q * 2^log_sigma + low
//    factors into g^(q 2^log_sigma) * g^low, and g^low is constant along the 256-point DFT (over
//    q), so the first pass applies only the 256 factors g^(q 2^log_sigma) (gathered into shared
//    memory once per CTA) and the final pass folds g^low into its twiddle table (in_tw_scale).  No
//    per-element scale loads from global memory remain (they were the largest stall).
// (The first, non-persistent versions of these kernels — one tile per CTA, loads straight into
// registers — are in the history; profiles/r1_ntt_r16_kernels_v1/v2.txt are their ncu captures.)


// ---- persistent radix-16 passes with asynchronous tile prefetch ---------------------------------
// ncu on the one-tile-per-CTA versions (profiles/r1_ntt_r16_kernels_v2.txt): issue slots 41-47 % busy, the
// largest stall by far is the global-load scoreboard — every CTA loads, then computes, then stores,
// and with 4 CTAs per SM the load phases are not covered.  Here a CTA walks over tiles and the
// NEXT tile's elements travel global -> shared with cp.async while the current tile is being
// transformed.  Every thread copies exactly the 16 elements it will later pick up, into exactly
// the shared-memory slots it later uses for the register exchange, so the staging buffer doubles
// as the exchange buffer and no barrier is needed for the copies themselves (two buffers of one
// tile each; 3 CTAs per SM).
__device__ __forceinline__ void cp_async8(u64* smem_dst, const u64* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int R16P_MIN_BLOCKS = 3;
constexpr size_t R16P_STRIDED_SMEM = (2 * 256 * 16 + 128 + 256) * sizeof(u64);
constexpr size_t R16P_FINAL_SMEM = (2 * 256 * 17 + 128 + 2 * 256) * sizeof(u64);

// pass_strided_r16 over tiles (tile_x < tiles_x, column < ncols), tile = column * tiles_x + tile_x.
template <bool INVERSE, bool OUT_TW>
__global__ void __launch_bounds__(THREADS, R16P_MIN_BLOCKS)
pass_strided_r16p(const u64* __restrict__ src, u64 src_col_stride, u64* __restrict__ dst,
                  u64 dst_col_stride, unsigned log_B, const u64* __restrict__ in_scale, Roots R,
                  unsigned tiles_x, unsigned ntiles) {
  extern __shared__ u64 dyn[];
  u64* buf = dyn;                  // [2][256 * 16]
  u64* tw = dyn + 2 * 256 * 16;    // [128]
  u64* sq = tw + 128;              // [256]
  const unsigned log_sigma = log_B - 8;
  const unsigned tiles_per_block_log = log_sigma - 4;
  const unsigned t = threadIdx.x & 15, qa = threadIdx.x >> 4;
  if (threadIdx.x < 128) tw[threadIdx.x] = root_of<INVERSE>(R, 8, threadIdx.x);
  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  auto issue = [&](unsigned tl, u64* b) {
    const unsigned bx = tl % tiles_x, by = tl / tiles_x;
    const u64 blk = bx >> tiles_per_block_log;
    const u64 low0 = (u64)(bx & ((1u << tiles_per_block_log) - 1)) << 4;
    const u64* p = src + (u64)by * src_col_stride + (blk << log_B) + low0 + t;
#pragma unroll
    for (int j = 0; j < 16; j++) cp_async8(b + (16 * j + qa) * 16 + t, p + ((u64)(16 * j + qa) << log_sigma));
  };
  issue(tile, buf);
  cp_async_commit();
  // the split input scaling needs one block per transform (blk == 0), which is what OUT_TW = false
  // is used for; the general form reads the scale per element below
  if (!OUT_TW && in_scale) sq[threadIdx.x] = __ldg(in_scale + ((u64)threadIdx.x << log_sigma));
  __syncthreads();  // tw, sq ready
  for (unsigned it = 0;; it++) {
    u64* cur = buf + (it & 1) * (256 * 16);
    const unsigned next = tile + gridDim.x;
    if (next < ntiles) issue(next, buf + ((it & 1) ^ 1) * (256 * 16));
    cp_async_commit();
    cp_async_wait<1>();  // this tile's copies (issued one iteration ago) have landed
    const unsigned bx = tile % tiles_x, by = tile / tiles_x;
    const u64 blk = bx >> tiles_per_block_log;
    const u64 low0 = (u64)(bx & ((1u << tiles_per_block_log) - 1)) << 4;
    const u64 base = blk << log_B;
    u64 x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[(16 * j + qa) * 16 + t];
    if (in_scale) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (OUT_TW) x[j] = gl::mul_lazy_fma(x[j], __ldg(in_scale + base + ((u64)(16 * j + qa) << log_sigma) + low0 + t));
        else x[j] = gl::mul_lazy_fma(x[j], sq[16 * j + qa]);
      }
    }
    dif16<true>(x, tw, 16, qa, 0);
#pragma unroll
    for (int j = 0; j < 16; j++) cur[(16 * j + qa) * 16 + t] = x[j];  // own slots only
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[(16 * qa + j) * 16 + t];
    __syncthreads();  // everyone is done with `cur` before the next iteration refills it
    dif16<false>(x, tw, 1, 0, 4);
    u64* out = dst + (u64)by * dst_col_stride + base + low0 + t;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const unsigned q = 16 * qa + j;
      if (OUT_TW) {
        const u64 e = (low0 + t) * (u64)brev(q, 8);
        out[(u64)q << log_sigma] = e ? gl::mul_lazy_fma(x[j], root_of<INVERSE>(R, log_B, e)) : x[j];
      } else {
        out[(u64)q << log_sigma] = x[j];
      }
    }
    if (next >= ntiles) break;
    tile = next;
  }
  cp_async_wait<0>();
}

// pass_final_r16 over tiles (tile_x < tiles_x, tile_y < tiles_y), tile = tile_y * tiles_x + tile_x
// (STORE_LEAF: tile_y = group of 16 columns; STORE_NATURAL: tile_y = column).
template <bool INVERSE, int MODE>
__global__ void __launch_bounds__(THREADS, R16P_MIN_BLOCKS)
pass_final_r16p(const u64* __restrict__ src, u64 src_col_stride, unsigned ncols,
                u64* __restrict__ dst, u64 dst_stride, u64 row0, unsigned log_n,
                const u64* __restrict__ in_scale, u64 out_scale, Roots R, unsigned in_tw_log_B,
                const u64* __restrict__ in_tw_scale, unsigned tiles_x, unsigned ntiles) {
  extern __shared__ u64 dyn[];
  u64* buf = dyn;                  // [2][256 * 17]
  u64* tw = dyn + 2 * 256 * 17;    // [128]
  u64* tws = tw + 128;             // [2][256]
  const unsigned log_nb = log_n - 8;
  const bool in_tw = MODE == STORE_LEAF && in_tw_log_B != 0;
  const unsigned q_lo = threadIdx.x & 15, lane_a = threadIdx.x >> 4;
  const unsigned lane_b = threadIdx.x & 15, q_hi = threadIdx.x >> 4;
  if (threadIdx.x < 128) tw[threadIdx.x] = root_of<INVERSE>(R, 8, threadIdx.x);
  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  // source of this thread's 16 elements of tile tl (nullptr: column beyond ncols)
  auto source = [&](unsigned tl) -> const u64* {
    const unsigned bx = tl % tiles_x, by = tl / tiles_x;
    if (MODE == STORE_LEAF) {
      const unsigned col = by * 16 + lane_a;
      return col < ncols ? src + (u64)col * src_col_stride + ((u64)bx << 8) + q_lo : nullptr;
    }
    return src + (u64)by * src_col_stride + ((u64)brev(bx * 16 + lane_a, log_nb) << 8) + q_lo;
  };
  auto issue = [&](unsigned tl, u64* b) {
    const u64* p = source(tl);
    if (p) {
#pragma unroll
      for (int j = 0; j < 16; j++) cp_async8(b + (16 * j + q_lo) * 17 + lane_a, p + 16 * j);
    }
  };
  auto twiddle = [&](unsigned tl) -> u64 {  // pending four-step twiddle (times the g^low scaling)
    return root_of<INVERSE>(R, in_tw_log_B, (u64)threadIdx.x * brev((tl % tiles_x) & 255u, 8));
  };
  const u64 tw_scale = (in_tw && in_tw_scale) ? __ldg(in_tw_scale + threadIdx.x) : 1;
  issue(tile, buf);
  cp_async_commit();
  if (in_tw) tws[threadIdx.x] = gl::mul_lazy_fma(twiddle(tile), tw_scale);
  for (unsigned it = 0;; it++) {
    u64* cur = buf + (it & 1) * (256 * 17);
    const u64* twc = tws + (it & 1) * 256;
    const unsigned next = tile + gridDim.x;
    u64 w_next = 0;
    if (next < ntiles) {
      issue(next, buf + ((it & 1) ^ 1) * (256 * 17));
      if (in_tw) w_next = twiddle(next);  // the load completes behind this tile's arithmetic
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();  // tw / this tile's tws visible
    const bool have = source(tile) != nullptr;
    u64 x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = have ? cur[(16 * j + q_lo) * 17 + lane_a] : 0;
    if (in_scale && have) {
      const unsigned bx = tile % tiles_x;
      const u64 pos0 = MODE == STORE_LEAF ? ((u64)bx << 8) : ((u64)brev(bx * 16 + lane_a, log_nb) << 8);
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] = gl::mul_lazy_fma(x[j], __ldg(in_scale + pos0 + 16 * j + q_lo));
    }
    if (in_tw) {
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] = gl::mul_lazy_fma(x[j], twc[16 * j + q_lo]);
    }
    dif16<true>(x, tw, 16, q_lo, 0);
#pragma unroll
    for (int j = 0; j < 16; j++) cur[(16 * j + q_lo) * 17 + lane_a] = x[j];  // own slots only
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[(16 * q_hi + j) * 17 + lane_b];
    if (in_tw && next < ntiles) tws[((it & 1) ^ 1) * 256 + threadIdx.x] = gl::mul_lazy_fma(w_next, tw_scale);
    __syncthreads();  // everyone is done with `cur` before the next iteration refills it
    dif16<false>(x, tw, 1, 0, 4);
    const unsigned bx = tile % tiles_x, by = tile / tiles_x;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const unsigned q = 16 * q_hi + j;
      u64 v = x[j];
      v = (out_scale != 1) ? gl::mul(v, out_scale) : gl::canon(v);
      if (MODE == STORE_LEAF) {
        const unsigned col = by * 16 + lane_b;
        const u64 pos = ((u64)bx << 8) + q;
        if (col < ncols) dst[(row0 + pos) * dst_stride + col] = v;
      } else {
        const u64 nat = ((u64)brev(q, 8) << log_nb) + (u64)bx * 16 + lane_b;
        dst[(u64)by * dst_stride + nat] = v;
      }
    }
    if (next >= ntiles) break;
    tile = next;
  }
  cp_async_wait<0>();
}

// ---- the same passes with TMA tile loads ------------------------------------------------------------
// The tile of a strided pass is a box of a strided tensor — 16 consecutive elements (128 B) x 256
// rows of pitch 2^log_sigma — and the tile of the leaf-order final pass a box of 256 elements x 16
// columns: ONE cp.async.bulk.tensor per tile, issued by one thread and completed on an mbarrier,
// replaces 16 cp.async + their address arithmetic per thread (4096 LDGSTS per tile), and columns
// beyond ncols arrive as zeros (out-of-bounds fill) instead of through a branch.  The tensor maps
// are encoded on the host per launch (vpbs_commit.cu, make_tile_map) and passed as __grid_constant__.
// Hazards: the staging buffer is also the exchange buffer (generic-proxy stores), so every thread
// executes fence.proxy.async before the barrier that precedes the next bulk copy into it.
namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void load_4d(void* dst, const void* map, u64* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void load_2d(void* dst, const void* map, u64* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
struct alignas(64) TileMap {  // same size and alignment as the driver's CUtensorMap
  unsigned long long opaque[16];
};
constexpr unsigned TILE_BYTES = 256 * 16 * sizeof(u64);
}  // namespace tma

constexpr size_t R16T_STRIDED_SMEM = 2 * tma::TILE_BYTES + (128 + 256 + 2) * sizeof(u64) + 128;
constexpr size_t R16T_FINAL_SMEM = 2 * tma::TILE_BYTES + (128 + 2 * 256 + 2) * sizeof(u64) + 128;

// pass_strided_r16p with the tile fetched by TMA.  map: rank-4 tensor over the source,
// (low < 2^log_sigma, q < 256, block < 2^(log_n - log_B), column < ncols), box 16 x 256 x 1 x 1.
template <bool INVERSE, bool OUT_TW>
__global__ void __launch_bounds__(THREADS, R16P_MIN_BLOCKS)
pass_strided_r16t(const __grid_constant__ tma::TileMap map, u64* __restrict__ dst, u64 dst_col_stride,
                  unsigned log_B, const u64* __restrict__ in_scale, Roots R, unsigned tiles_x,
                  unsigned ntiles) {
  extern __shared__ u64 dyn_raw[];
  // 128-byte aligned start, by pointer arithmetic ON the shared array (a round trip through an integer
  // would turn every access below into a generic LD / ST instead of LDS / STS)
  u64* dyn = dyn_raw + (((128u - (tma::smem_u32(dyn_raw) & 127u)) & 127u) >> 3);
  u64* buf = dyn;                  // [2][256 * 16], dense [q][low]: what the box lands as
  u64* tw = dyn + 2 * 256 * 16;    // [128]
  u64* sq = tw + 128;              // [256]
  u64* bar = sq + 256;             // [2]
  const unsigned log_sigma = log_B - 8;
  const unsigned tiles_per_block_log = log_sigma - 4;
  const unsigned t = threadIdx.x & 15, qa = threadIdx.x >> 4;
  if (threadIdx.x < 128) tw[threadIdx.x] = root_of<INVERSE>(R, 8, threadIdx.x);
  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    tma::mbar_init(bar + 1, 1);
    tma::fence_mbar_init();
  }
  auto issue = [&](unsigned tl, unsigned b) {  // one thread
    const unsigned bx = tl % tiles_x, by = tl / tiles_x;
    tma::mbar_expect_tx(bar + b, tma::TILE_BYTES);
    tma::load_4d(buf + b * (256 * 16), &map, bar + b, (int)((bx & ((1u << tiles_per_block_log) - 1)) << 4), 0,
                 (int)(bx >> tiles_per_block_log), (int)by);
  };
  if (!OUT_TW && in_scale) sq[threadIdx.x] = __ldg(in_scale + ((u64)threadIdx.x << log_sigma));
  __syncthreads();  // barriers initialised; tw, sq ready
  if (threadIdx.x == 0) issue(tile, 0);
  for (unsigned it = 0;; it++) {
    u64* cur = buf + (it & 1) * (256 * 16);
    const unsigned next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) issue(next, (it & 1) ^ 1);
    tma::mbar_wait(bar + (it & 1), (it >> 1) & 1);
    const unsigned bx = tile % tiles_x, by = tile / tiles_x;
    const u64 blk = bx >> tiles_per_block_log;
    const u64 low0 = (u64)(bx & ((1u << tiles_per_block_log) - 1)) << 4;
    const u64 base = blk << log_B;
    u64 x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[(16 * j + qa) * 16 + t];
    if (in_scale) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (OUT_TW) x[j] = gl::mul_lazy_fma(x[j], __ldg(in_scale + base + ((u64)(16 * j + qa) << log_sigma) + low0 + t));
        else x[j] = gl::mul_lazy_fma(x[j], sq[16 * j + qa]);
      }
    }
    dif16<true>(x, tw, 16, qa, 0);
#pragma unroll
    for (int j = 0; j < 16; j++) cur[(16 * j + qa) * 16 + t] = x[j];  // own slots only
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[(16 * qa + j) * 16 + t];
    tma::fence_proxy_async();  // our stores to `cur` are ordered before the bulk copy that refills it
    __syncthreads();
    dif16<false>(x, tw, 1, 0, 4);
    u64* out = dst + (u64)by * dst_col_stride + base + low0 + t;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const unsigned q = 16 * qa + j;
      if (OUT_TW) {
        const u64 e = (low0 + t) * (u64)brev(q, 8);
        out[(u64)q << log_sigma] = e ? gl::mul_lazy_fma(x[j], root_of<INVERSE>(R, log_B, e)) : x[j];
      } else {
        out[(u64)q << log_sigma] = x[j];
      }
    }
    if (next >= ntiles) break;
    tile = next;
  }
}

// pass_final_r16p with the tile fetched by TMA.
//   STORE_LEAF:    map = rank-2 tensor (position < n, column < ncols), box 256 x 16.
//   STORE_NATURAL: map = rank-4 tensor (position < 256, block_lo < 2^(log_nb - 4), block_hi < 16,
//                  column), box 256 x 1 x 16 x 1 at block_lo = bitrev(tile_x): the 16 blocks whose
//                  bit-reversed ids are consecutive lie 2^(log_nb - 4) blocks apart; row i of the tile
//                  is block_hi = i, i.e. the block the cp.async form calls lane bitrev4(i).
// The tile lands dense as [lane][q] (no padding possible), so the register exchange swizzles the low
// four bits of q with the lane (slot = lane * 256 + (q ^ lane)): the stage-1 stores then touch slots
// read by threads of the same half-warp (a __syncwarp orders them) and both exchange accesses are
// bank-conflict free (the padded cp.async form had 7.7e5 conflicts per launch).
template <bool INVERSE, int MODE>
__global__ void __launch_bounds__(THREADS, R16P_MIN_BLOCKS)
pass_final_r16t(const __grid_constant__ tma::TileMap map, unsigned ncols, u64* __restrict__ dst,
                u64 dst_stride, u64 row0, unsigned log_n, u64 out_scale, Roots R, unsigned in_tw_log_B,
                const u64* __restrict__ in_tw_scale, unsigned tiles_x, unsigned ntiles) {
  extern __shared__ u64 dyn_raw[];
  // 128-byte aligned start, by pointer arithmetic ON the shared array (a round trip through an integer
  // would turn every access below into a generic LD / ST instead of LDS / STS)
  u64* dyn = dyn_raw + (((128u - (tma::smem_u32(dyn_raw) & 127u)) & 127u) >> 3);
  u64* buf = dyn;                  // [2][16 * 256], dense [lane][q]
  u64* tw = dyn + 2 * 256 * 16;    // [128]
  u64* tws = tw + 128;             // [2][256]
  u64* bar = tws + 2 * 256;        // [2]
  const unsigned log_nb = log_n - 8;
  const bool in_tw = MODE == STORE_LEAF && in_tw_log_B != 0;
  const unsigned q_lo = threadIdx.x & 15, lane_a = threadIdx.x >> 4;
  const unsigned lane_b = threadIdx.x & 15, q_hi = threadIdx.x >> 4;
  if (threadIdx.x < 128) tw[threadIdx.x] = root_of<INVERSE>(R, 8, threadIdx.x);
  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    tma::mbar_init(bar + 1, 1);
    tma::fence_mbar_init();
  }
  auto issue = [&](unsigned tl, unsigned b) {  // one thread
    tma::mbar_expect_tx(bar + b, tma::TILE_BYTES);
    if (MODE == STORE_LEAF)
      tma::load_2d(buf + b * (256 * 16), &map, bar + b, (int)((tl % tiles_x) << 8), (int)((tl / tiles_x) << 4));
    else
      tma::load_4d(buf + b * (256 * 16), &map, bar + b, 0, (int)brev(tl % tiles_x, log_nb - 4), 0,
                   (int)(tl / tiles_x));
  };
  auto twiddle = [&](unsigned tl) -> u64 {  // pending four-step twiddle (times the g^low scaling)
    return root_of<INVERSE>(R, in_tw_log_B, (u64)threadIdx.x * brev((tl % tiles_x) & 255u, 8));
  };
  const u64 tw_scale = (in_tw && in_tw_scale) ? __ldg(in_tw_scale + threadIdx.x) : 1;
  if (in_tw) tws[threadIdx.x] = gl::mul_lazy_fma(twiddle(tile), tw_scale);
  __syncthreads();  // barriers initialised; tw and the first tws visible
  if (threadIdx.x == 0) issue(tile, 0);
  for (unsigned it = 0;; it++) {
    u64* cur = buf + (it & 1) * (256 * 16);
    const u64* twc = tws + (it & 1) * 256;
    const unsigned next = tile + gridDim.x;
    u64 w_next = 0;
    if (next < ntiles) {
      if (threadIdx.x == 0) issue(next, (it & 1) ^ 1);
      if (in_tw) w_next = twiddle(next);  // the load completes behind this tile's arithmetic
    }
    tma::mbar_wait(bar + (it & 1), (it >> 1) & 1);
    u64 x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[lane_a * 256 + 16 * j + q_lo];
    if (in_tw) {
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] = gl::mul_lazy_fma(x[j], twc[16 * j + q_lo]);
    }
    dif16<true>(x, tw, 16, q_lo, 0);
    __syncwarp();  // the slots written next were read by threads of this half-warp
#pragma unroll
    for (int j = 0; j < 16; j++) cur[lane_a * 256 + 16 * j + (q_lo ^ lane_a)] = x[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = cur[lane_b * 256 + 16 * q_hi + (j ^ lane_b)];
    if (in_tw && next < ntiles) tws[((it & 1) ^ 1) * 256 + threadIdx.x] = gl::mul_lazy_fma(w_next, tw_scale);
    tma::fence_proxy_async();
    __syncthreads();  // everyone is done with `cur` before the next bulk copy refills it
    dif16<false>(x, tw, 1, 0, 4);
    const unsigned bx = tile % tiles_x, by = tile / tiles_x;
    if (MODE == STORE_LEAF) {
      const unsigned col = by * 16 + lane_b;
      if (col < ncols) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const u64 pos = ((u64)bx << 8) + 16 * q_hi + j;
          const u64 v = (out_scale != 1) ? gl::mul(x[j], out_scale) : gl::canon(x[j]);
          // streaming: leaf rows are not re-read before the whole LDE is done
          __stcs(reinterpret_cast<unsigned long long*>(dst + (row0 + pos) * dst_stride + col), v);
        }
      }
    } else {
      u64* out = dst + (u64)by * dst_stride + (u64)bx * 16 + brev(lane_b, 4);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const u64 v = (out_scale != 1) ? gl::mul(x[j], out_scale) : gl::canon(x[j]);
        out[(u64)brev(16 * q_hi + j, 8) << log_nb] = v;
      }
    }
    if (next >= ntiles) break;
    tile = next;
  }
}

// ---- polynomial evaluation at quadratic-extension points --------------------------------------------
// F[X]/(X^2 - 7): (a0 + a1 X)(b0 + b1 X) = (a0 b0 + 7 a1 b1) + (a0 b1 + a1 b0) X.
struct Ext2 {
  u64 re, im;
};
__device__ __forceinline__ Ext2 ext_mul(Ext2 a, Ext2 b) {
  const u64 t = gl::mul(a.im, b.im);
  return Ext2{gl::add(gl::mul(a.re, b.re), gl::mul(7, t)), gl::add(gl::mul(a.re, b.im), gl::mul(a.im, b.re))};
}
// One CTA of EVAL_THREADS threads per (polynomial, point).  Thread t sums the coefficients
// j = t (mod T) by Horner in y = x^T (coalesced reads), scales by x^t and the CTA adds the partial
// values.  T = 1024 for long polynomials: the Horner chain of a thread is a serial run of extension
// multiplies (2^16 coefficients: 64 steps instead of 256 with T = 256, and four times the warps to
// hide them behind; 170 -> see DESIGN.md §9.2 for the measured effect), T = 256 below 2^14.
// a b + c + d formed in 128 bits (it fits: (2^64-1)^2 + 2 (2^64-1) = 2^128 - 1) and reduced once:
// the Horner step acc <- acc y + c_k costs two of these and two plain products (~66 instructions)
// instead of five canonical multiplies and three canonical adds (~124).
__device__ __forceinline__ u64 mul_add2_lazy(u64 a, u64 b, u64 c, u64 d) {
  const unsigned __int128 r = (unsigned __int128)a * b + c + d;
  return gl::reduce_words((u64)r, (u64)(r >> 64));
}
template <int T>
__global__ void __launch_bounds__(T)
eval_ext2(const u64* __restrict__ coeffs, u64 col_stride, unsigned log_n,
          const u64* __restrict__ points, u64* __restrict__ out, unsigned ncols) {
  constexpr int LOG_T = T == 1024 ? 10 : 8;
  static_assert(T == 1024 || T == 256, "eval_ext2: 256 or 1024 threads");
  __shared__ u64 sre[T], sim[T];
  __shared__ u64 plo[2][32], phi[2][32];  // x^i and x^(32 i), i < 32 (re, im)
  const unsigned col = blockIdx.x, pt = blockIdx.y, t = threadIdx.x;
  const u64 n = 1ULL << log_n;
  const Ext2 x{gl::canon(points[2 * pt]), gl::canon(points[2 * pt + 1])};
  Ext2 y = x;  // x^T
#pragma unroll
  for (int i = 0; i < LOG_T; i++) y = ext_mul(y, y);
  // x^t = x^(t mod 32) * x^(32 (t / 32)): two 32-entry tables built by the first two warps (a
  // per-thread binary exponentiation cost a quarter of the kernel)
  if (t < 64) {
    Ext2 b = x;
    if (t >= 32) {
#pragma unroll
      for (int i = 0; i < 5; i++) b = ext_mul(b, b);  // x^32
    }
    Ext2 p{1, 0};
    for (unsigned e = t & 31; e; e >>= 1) {
      if (e & 1) p = ext_mul(p, b);
      b = ext_mul(b, b);
    }
    if (t < 32) {
      plo[0][t] = p.re;
      plo[1][t] = p.im;
    } else {
      phi[0][t - 32] = p.re;
      phi[1][t - 32] = p.im;
    }
  }
  __syncthreads();
  const u64* c = coeffs + (u64)col * col_stride;
  Ext2 acc{0, 0};
  if (t < n) {
    const u64 y7 = gl::mul(7, y.im);
    const u64 k_hi = (n - 1 - t) >> LOG_T;  // largest k with T k + t < n
    u64 are = 0, aim = 0;                   // arbitrary u64 representatives inside the chain
    for (u64 k = k_hi + 1; k-- > 0;) {
      const u64 ck = __ldg(c + (k << LOG_T) + t);
      const u64 t1 = gl::mul_lazy_fma(are, y.re), t2 = gl::mul_lazy_fma(are, y.im);
      are = mul_add2_lazy(aim, y7, t1, ck);   // re: a.re y.re + 7 a.im y.im + c_k
      aim = mul_add2_lazy(aim, y.re, t2, 0);  // im: a.re y.im + a.im y.re
    }
    acc = Ext2{gl::canon(are), gl::canon(aim)};
    const Ext2 lo{plo[0][t & 31], plo[1][t & 31]}, hi{phi[0][(t >> 5) & 31], phi[1][(t >> 5) & 31]};
    acc = ext_mul(acc, ext_mul(lo, hi));
  }
  sre[t] = acc.re;
  sim[t] = acc.im;
  __syncthreads();
  for (unsigned s = T / 2; s > 0; s >>= 1) {
    if (t < s) {
      sre[t] = gl::add(sre[t], sre[t + s]);
      sim[t] = gl::add(sim[t], sim[t + s]);
    }
    __syncthreads();
  }
  if (t == 0) {
    out[2 * ((u64)pt * ncols + col)] = sre[0];
    out[2 * ((u64)pt * ncols + col) + 1] = sim[0];
  }
}

// ---- FRI commit-phase helpers (extension elements as (re, im) pairs) ---------------------------------
// leaves[k] = canon(values[bitrev(k)]): reverse_index_bits_in_place + chunking is a pure re-indexing
// because a leaf is `arity` consecutive elements of the reversed vector.
__global__ void fri_gather_leaves(const ulonglong2* __restrict__ values, unsigned log_len,
                                  ulonglong2* __restrict__ leaves) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (1ULL << log_len)) return;
  const u64 src = log_len ? (__brevll(k) >> (64 - log_len)) : 0;
  const ulonglong2 v = values[src];
  leaves[k] = make_ulonglong2(gl::canon(v.x), gl::canon(v.y));
}
// coeffs'[j] = sum_i coeffs[j * arity + i] * beta^i  (Horner from the top); written interleaved to
// `folded` and planar (re column, im column) to `planar` for the transform that follows.
__global__ void fri_fold(const ulonglong2* __restrict__ coeffs, u64 out_len, unsigned arity_bits,
                         u64 beta_re, u64 beta_im, ulonglong2* __restrict__ folded,
                         u64* __restrict__ planar) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= out_len) return;
  const Ext2 beta{beta_re, beta_im};
  const u64 arity = 1ULL << arity_bits;
  Ext2 acc{0, 0};
  for (u64 i = arity; i-- > 0;) {
    const ulonglong2 c = coeffs[j * arity + i];
    acc = ext_mul(acc, beta);
    acc.re = gl::add(acc.re, gl::canon(c.x));
    acc.im = gl::add(acc.im, gl::canon(c.y));
  }
  folded[j] = make_ulonglong2(acc.re, acc.im);
  planar[j] = acc.re;
  planar[out_len + j] = acc.im;
}
// planar[j] = canon(re), planar[out_len + j] = canon(im) for j < in_len, zero beyond (zero padding
// of PolynomialCoeffs::lde made explicit for the extension-field final polynomial).
__global__ void deinterleave2_pad(const ulonglong2* __restrict__ in, u64 in_len, u64 out_len,
                                  u64* __restrict__ planar) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= out_len) return;
  ulonglong2 v = make_ulonglong2(0, 0);
  if (j < in_len) v = in[j];
  planar[j] = gl::canon(v.x);
  planar[out_len + j] = gl::canon(v.y);
}
__global__ void interleave2(const u64* __restrict__ planar, u64 len, ulonglong2* __restrict__ out) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < len) out[j] = make_ulonglong2(planar[j], planar[len + j]);
}

// ---- tables ------------------------------------------------------------------------------------
// w[t] = omega_N^t for t < N/2.
__global__ void fill_roots(u64* w, unsigned log_N) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (1ULL << (log_N - 1))) return;
  w[t] = gl::pow(gl::primitive_root_of_unity(log_N), t);
}
// out[b * n + j] = (shift * omega_m^bitrev_r(b))^j  for b < 2^rate_bits, j < n = 2^log_n,
// m = n << rate_bits: the per-block coset factors of the LDE (rate_bits = 0: plain shift^j).
__global__ void fill_coset_powers(u64* out, unsigned log_n, unsigned rate_bits, u64 shift) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1ULL << (log_n + rate_bits))) return;
  const u64 j = i & ((1ULL << log_n) - 1);
  const unsigned b = (unsigned)(i >> log_n);
  const u64 g = gl::mul(gl::canon(shift),
                        gl::pow(gl::primitive_root_of_unity(log_n + rate_bits), brev(b, rate_bits)));
  out[i] = gl::pow(g, j);
}

}  // namespace ntt
