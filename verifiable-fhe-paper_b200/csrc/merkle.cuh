// merkle.cuh — Poseidon leaf hashing and level-by-level Merkle reduction (sm_100a).
//
// Replaces, for H = PoseidonHash over GoldilocksField:
//   [P2] plonky2 0.2.0 src/hash/hashing.rs      hash_n_to_m_no_pad (overwrite-mode sponge), compress
//   [P2] plonky2 0.2.0 src/plonk/config.rs      Hasher::hash_or_noop / two_to_one
//   [P2] plonky2 0.2.0 src/hash/merkle_tree.rs  MerkleTree::new / fill_digests_buf / fill_subtree
// reached from PolynomialBatch::from_coeffs ("build Merkle tree" scope) in every prove()/build() of
// the reference (/root/reference/src/vtfhe/ivc_based_vpbs.rs:275,302,333,364).
//
// The digests buffer uses plonky2's layout so MerkleTree::prove's index arithmetic is unchanged:
// inside one cap subtree, the sibling pair q of layer i (layer 0 = leaf digests) lives at hash
// indices 2*((q << (i+1)) + (1 << i) - 1) and +1.
#pragma once
#include "poseidon.cuh"

namespace merkle {

using gl::u32;
using gl::u64;

struct Hash4 {
  u64 e[4];
};

__device__ __forceinline__ void store_hash(u64* dst, const u64 (&s)[poseidon::WIDTH]) {
  // hashes are 32-byte aligned (cudaMalloc base + 32 * index)
  ulonglong2* d = reinterpret_cast<ulonglong2*>(dst);
  d[0] = make_ulonglong2(gl::canon(s[0]), gl::canon(s[1]));
  d[1] = make_ulonglong2(gl::canon(s[2]), gl::canon(s[3]));
}

// hash_or_noop of one row of `width` elements (any u64 values).
__device__ __forceinline__ void hash_row(const u64* __restrict__ row, u32 width, u64* out) {
  u64 s[poseidon::WIDTH];
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++) s[i] = 0;
  if (width <= 4) {  // [P2] hash_or_noop: inputs that fit in a hash are copied, not hashed
#pragma unroll
    for (int i = 0; i < 4; i++)
      if ((u32)i < width) s[i] = __ldg(row + i);
    store_hash(out, s);
    return;
  }
  // one permutation call site (code size); a short last chunk overwrites only the first lanes.
  // (128-bit loads for rows of even width — four LDG.128 instead of eight LDG.64 per absorb — were
  // measured in round 2: 5.17 against 5.15 ms for 2^19 x 128, the extra path costs more than the
  // four loads it saves out of ~14,600 instructions per permutation; L1 already merges the 64-bit
  // loads of a row into whole sectors, DRAM traffic equals the algorithmic bytes.)
  for (u32 off = 0; off < width; off += poseidon::RATE) {
#pragma unroll
    for (int i = 0; i < poseidon::RATE; i++)
      if (off + i < width) s[i] = __ldg(row + off + i);
    poseidon::permute_lazy(s);
  }
  store_hash(out, s);
}

// Position (in hashes) of leaf digest `l` of a cap subtree inside that subtree's digest buffer.
__device__ __forceinline__ u64 leaf_digest_pos(u64 l) { return 4 * (l >> 1) + (l & 1); }

#ifndef VPBS_HASH_THREADS
#define VPBS_HASH_THREADS 128
#endif
#ifndef VPBS_HASH_MIN_BLOCKS
// Occupancy does not matter any more (the kernel is bound by the ALU pipe in the full rounds and
// the FP64 pipe in the pair steps): 3 / 4 / 5 / 6 CTAs per SM measure 5.55 / 5.54 / 5.52 / 5.59 ms,
// 64- and 256-thread CTAs the same.
#define VPBS_HASH_MIN_BLOCKS 5  // round 2: with the 3-product squaring ptxas takes 102 registers unless held to 96
#endif
// One thread per leaf.  all_cap: the tree has no digests, leaf hashes are the cap.
__global__ void __launch_bounds__(VPBS_HASH_THREADS, VPBS_HASH_MIN_BLOCKS)
hash_leaves(const u64* __restrict__ leaves, u64 nleaves, u32 width, u64* __restrict__ out,
            unsigned log_sub, u64 sub_digests, int all_cap) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nleaves) return;
  u64 pos = k;
  if (!all_cap) {
    const u64 sub = k >> log_sub, l = k & ((1ULL << log_sub) - 1);
    pos = sub * sub_digests + leaf_digest_pos(l);
  }
  hash_row(leaves + k * width, width, out + 4 * pos);
}

// The same sponge in pieces: columns [c0, c1) of every row absorbed on top of the state the previous
// piece left (c0 a multiple of the rate; `state` holds the 12 words of every leaf, word-major so
// that a warp's accesses coalesce).  The sponge state after the first c columns depends on those
// columns only, so a wide batch whose columns arrive chunk by chunk from a slow source (pageable host
// memory through the staging ring) can be hashed while the later chunks are still on their way.
// first: start from the zero state; last: write the digest instead of the state.
__global__ void __launch_bounds__(VPBS_HASH_THREADS, VPBS_HASH_MIN_BLOCKS)
hash_leaves_part(const u64* __restrict__ leaves, u64 nleaves, u32 width, u32 c0, u32 c1,
                 u64* __restrict__ state, int first, int last, u64* __restrict__ out, unsigned log_sub,
                 u64 sub_digests, int all_cap) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nleaves) return;
  u64 s[poseidon::WIDTH];
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++) s[i] = first ? 0 : state[(u64)i * nleaves + k];
  const u64* row = leaves + k * width;
  for (u32 off = c0; off < c1; off += poseidon::RATE) {
#pragma unroll
    for (int i = 0; i < poseidon::RATE; i++)
      if (off + i < c1) s[i] = __ldg(row + off + i);
    poseidon::permute_lazy(s);
  }
  if (!last) {
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) state[(u64)i * nleaves + k] = s[i];
    return;
  }
  u64 pos = k;
  if (!all_cap) {
    const u64 sub = k >> log_sub, l = k & ((1ULL << log_sub) - 1);
    pos = sub * sub_digests + leaf_digest_pos(l);
  }
  store_hash(out + 4 * pos, s);
}

// Layer `level` (>= 1) of every cap subtree: node jj = two_to_one(children pair jj of level-1).
// level == log_sub writes the subtree roots into `cap`.
__global__ void __launch_bounds__(128)
reduce_level(u64* __restrict__ digests, u64* __restrict__ cap, unsigned level, unsigned log_sub,
             u64 sub_digests, u64 nnodes) {
  const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnodes) return;
  const unsigned log_nodes = log_sub - level;
  const u64 sub = j >> log_nodes, jj = j & ((1ULL << log_nodes) - 1);
  u64* base = digests + 4 * sub * sub_digests;
  const ulonglong2* ch = reinterpret_cast<const ulonglong2*>(
      base + 4 * (2 * ((jj << level) + (1ULL << (level - 1)) - 1)));
  u64 s[poseidon::WIDTH];
  const ulonglong2 a = ch[0], b = ch[1], c = ch[2], d = ch[3];
  s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
  s[4] = c.x; s[5] = c.y; s[6] = d.x; s[7] = d.y;
  s[8] = s[9] = s[10] = s[11] = 0;
  poseidon::permute_lazy(s);
  u64* out = (level == log_sub)
                 ? cap + 4 * sub
                 : base + 4 * (2 * (((jj >> 1) << (level + 1)) + (1ULL << level) - 1) + (jj & 1));
  store_hash(out, s);
}

// Levels level_from..log_sub of every cap subtree in ONE launch: one CTA per subtree, one
// 16-thread group per node (poseidon::permute_coop), __syncthreads() between levels.  Used for the
// top of the tree, where per-level launches of the thread-per-node kernel are latency bound
// (~40 us per level; this kernel: ~7 us per level).
constexpr int COOP_THREADS = 1024;
__global__ void __launch_bounds__(COOP_THREADS)
reduce_levels_coop(u64* __restrict__ digests, u64* __restrict__ cap, unsigned level_from,
                   unsigned log_sub, u64 sub_digests) {
  extern __shared__ double coop_sh[];
  const u64 sub = blockIdx.x;
  u64* base = digests + 4 * sub * sub_digests;
  const unsigned group = threadIdx.x >> 4, l = threadIdx.x & 15, ngroups = blockDim.x >> 4;
  double* sh = coop_sh + 48 * group;
  for (unsigned level = level_from; level <= log_sub; level++) {
    const u64 nodes = 1ULL << (log_sub - level);
    for (u64 j0 = 0; j0 < nodes; j0 += ngroups) {  // same trip count for every thread of the CTA
      const u64 jj = j0 + group;
      const bool valid = jj < nodes;
      u64 w = 0;
      if (valid && l < 8)
        w = __ldcg(base + 4 * (2 * ((jj << level) + (1ULL << (level - 1)) - 1)) + l);
      // warps whose two groups both lie beyond the last node skip the work (warp-uniform branch)
      if (j0 + (group & ~1u) < nodes) poseidon::permute_coop(w, sh, l);
      if (valid && l < 4) {
        u64* out = (level == log_sub)
                       ? cap + 4 * sub
                       : base + 4 * (2 * (((jj >> 1) << (level + 1)) + (1ULL << level) - 1) + (jj & 1));
        out[l] = gl::canon(w);
      }
    }
    __syncthreads();  // this level's digests are visible to the CTA before the next level reads them
  }
}

// One level with one 16-thread group per node, nodes of all subtrees spread over the grid (for
// levels with a few thousand nodes: enough CTAs to use every SM, ~4x lower latency than the
// thread-per-node kernel).
__global__ void __launch_bounds__(128)
reduce_level_coop(u64* __restrict__ digests, u64* __restrict__ cap, unsigned level,
                  unsigned log_sub, u64 sub_digests, u64 nnodes) {
  __shared__ double sh_all[8 * 48];
  const unsigned group = threadIdx.x >> 4, l = threadIdx.x & 15;
  const u64 j = (u64)blockIdx.x * 8 + group;
  const bool valid = j < nnodes;
  const unsigned log_nodes = log_sub - level;
  const u64 sub = j >> log_nodes, jj = j & ((1ULL << log_nodes) - 1);
  u64* base = digests + 4 * sub * sub_digests;
  u64 w = 0;
  if (valid && l < 8) w = __ldcg(base + 4 * (2 * ((jj << level) + (1ULL << (level - 1)) - 1)) + l);
  if ((u64)blockIdx.x * 8 + (group & ~1u) < nnodes) poseidon::permute_coop(w, sh_all + 48 * group, l);
  if (valid && l < 4) {
    u64* out = (level == log_sub)
                   ? cap + 4 * sub
                   : base + 4 * (2 * (((jj >> 1) << (level + 1)) + (1ULL << level) - 1) + (jj & 1));
    out[l] = gl::canon(w);
  }
}

// Blinding columns: leaves[k][first_col + s] = salt[s][bitrev(first_leaf + k)] (salt vectors are
// extra lde_values() columns in natural order; they get transposed/bit-reversed like the rest).
__global__ void scatter_salt(const u64* __restrict__ salt, u64 m, unsigned log_m, u64 first_leaf,
                             u64 nleaves, u64* __restrict__ leaves, u32 width, u32 first_col) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nleaves * 4) return;
  const u64 k = i >> 2;
  const u32 s = (u32)(i & 3);
  const u64 g = first_leaf + k;
  const u64 nat = log_m ? (__brevll(g) >> (64 - log_m)) : 0;
  leaves[k * width + first_col + s] = gl::canon(__ldg(salt + (u64)s * m + nat));
}

// ---- lazy openings of a device-resident tree ----------------------------------------------------
// rows_out[i][:] = leaves[idx[i]][:]
__global__ void gather_rows(const u64* __restrict__ leaves, u32 width, const u64* __restrict__ idx,
                            u64 count, u64* __restrict__ rows_out) {
  const u64 row = blockIdx.x;
  if (row >= count) return;
  const u64* src = leaves + idx[row] * width;
  for (u32 c = threadIdx.x; c < width; c += blockDim.x) rows_out[row * width + c] = src[c];
}
// [P2] MerkleTree::prove index arithmetic, one thread per (query, layer).
__global__ void gather_siblings(const u64* __restrict__ digests, const u64* __restrict__ idx,
                                u64 count, unsigned num_layers, u64 sub_digests,
                                u64* __restrict__ out) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * num_layers) return;
  const u64 q = t / num_layers;
  const unsigned i = (unsigned)(t % num_layers);
  const u64 leaf = idx[q];
  const u64 tree = leaf >> num_layers, l = leaf & ((1ULL << num_layers) - 1);
  const u64 parity = (l >> i) & 1, pair = l >> (i + 1);
  const u64 sib = 2 * ((pair << (i + 1)) + (1ULL << i) - 1) + (1 - parity);
  const ulonglong2* src = reinterpret_cast<const ulonglong2*>(digests + 4 * (tree * sub_digests + sib));
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + 4 * t);
  dst[0] = src[0];
  dst[1] = src[1];
}

// ---- FRI proof-of-work grind ---------------------------------------------------------------------
// best: smallest qualifying candidate so far (init ~0); one candidate per thread.
__global__ void __launch_bounds__(128)
pow_grind(const u64* __restrict__ state, u32 witness_pos, u32 response_lane, u32 min_leading_zeros,
          u64 first, u64 count, unsigned long long* __restrict__ best) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const u64 cand = first + t;
  if (cand >= *best) return;  // a smaller witness is already known
  u64 s[poseidon::WIDTH];
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++) s[i] = (i == (int)witness_pos) ? gl::canon(cand) : state[i];
  poseidon::permute_lazy(s);
  u64 resp = 0;
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++)
    if (i == (int)response_lane) resp = gl::canon(s[i]);
  if ((u32)__clzll((long long)resp) >= min_leading_zeros) atomicMin(best, (unsigned long long)cand);
}

// ---- small batch entry points (tests / callers that hash outside a tree) -----------------------
__global__ void permute_batch(u64* states, u64 count) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  u64 s[poseidon::WIDTH];
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++) s[i] = states[k * poseidon::WIDTH + i];
  poseidon::permute_lazy(s);
#pragma unroll
  for (int i = 0; i < poseidon::WIDTH; i++) states[k * poseidon::WIDTH + i] = gl::canon(s[i]);
}
__global__ void two_to_one_batch(const u64* __restrict__ l, const u64* __restrict__ r, u64 count,
                                 u64* __restrict__ out) {
  const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  u64 s[poseidon::WIDTH];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s[i] = l[4 * k + i];
    s[4 + i] = r[4 * k + i];
    s[8 + i] = 0;
  }
  poseidon::permute_lazy(s);
  store_hash(out + 4 * k, s);
}

}  // namespace merkle
