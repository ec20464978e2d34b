// vpbs_commit.cu — context, launch orchestration and the C ABI of libvpbs_commit.so.
// See include/vpbs_commit.h for the contract and the plonky2 0.2.0 items each entry replaces.
//
// HBM layout of one commit (n = 2^log_n rows, C columns, r = rate_bits, m = n << r):
//   values   C x n   column-major   (caller / arena "in")
//   coeffs   C x n   column-major   (PolynomialBatch.polynomials)
//   work     C x n   column-major   scratch of the multi-pass NTT (stays L2-resident per block)
//   leaves   m x W   row-major      W = C (+4 salt); leaf k = natural LDE row bitrev(k)
//   digests  2(m - 2^h) x 4, plonky2 layout;  cap 2^h x 4
// There is no CPU fallback anywhere in this file: every path ends in a kernel launch or an error.
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched through the runtime
#include <cuda_runtime.h>

#include <cstdlib>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/vpbs_commit.h"
#include "host_stage.h"
#include "merkle.cuh"
#include "ntt.cuh"
#include "openings.cuh"
#include "permutation.cuh"

using gl::u32;
using gl::u64;

namespace {

std::mutex g_err_mu;
std::string g_global_err;

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct vpbs_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // twiddles
  u64* roots = nullptr;
  unsigned roots_log = 0;
  std::map<std::pair<unsigned, std::pair<unsigned, u64>>, u64*> coset_tables;  // (log_n,(r,shift))
  // grow-only arena
  std::map<std::string, Buf> arena;
  cudaEvent_t ev[10] = {};
  // host API: device->host copies run on their own stream, overlapped with the remaining kernels
  cudaStream_t copy_stream = nullptr;
  cudaStream_t h2d_stream = nullptr;  // inputs of wide batches arrive chunk by chunk on this one
  // Odd LDE blocks run on this stream with their own work buffer: one 256-point pass is 2048 CTAs
  // = 3.46 waves of 148 x 4, and the other stream's CTAs fill the tail wave (LDE 1.67 -> 1.46 ms).
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;

  // vpbs_commit_multi: events of the chunk uploads this context performs for all contexts, and the
  // event that marks the tail of this context's compute stream when such an upload starts
  std::vector<cudaEvent_t> peer_ev;
  cudaEvent_t peer_tail = nullptr;
  std::set<int> peers_enabled;

  unsigned sms = 148;  // persistent NTT passes launch sms * ntt::R16P_MIN_BLOCKS CTAs
  std::vector<cudaEvent_t> ov;  // coeffs ready, one per LDE block, commit done
  // Buffers of destroyed resident batches, kept for the next batch of the same shape (a prover
  // commits the same shapes every step; cudaMalloc/cudaFree of ~0.6 GB cost milliseconds).
  std::multimap<size_t, void*> pool;
  size_t pool_bytes = 0;
  // Live resident batches of this context: vpbs_ctx_destroy frees their device buffers and orphans
  // the handles (a batch outliving its context is then inert instead of a use-after-free).
  std::set<struct vpbs_batch*> batches;
  std::set<struct vpbs_sigmas*> sigma_sets;
  std::set<struct vpbs_fri*> fri_chains;
  // Host columns that are not page-locked travel through this pinned ring, filled by host_threads
  // copy threads (host_stage.h); 0 threads: leave such copies to the driver's own staging.
  hoststage::Ring ring;
  unsigned host_threads = 4;
  // Row-range sharding of the resident batches this context creates (vpbs_ctx_set_shard): batch
  // leaves / digests cover leaves [index * m / count, (index + 1) * m / count) only.
  u32 shard_index = 0, shard_count = 1;
};

// A commit kept in HBM (vpbs_batch_*): owns its device buffers, reads go through the context.
struct vpbs_batch {
  vpbs_ctx* ctx = nullptr;
  u32 ncols = 0, log_n = 0, rate_bits = 0, cap_height = 0, width = 0;
  bool coeff_inputs = false;
  u64 *coeffs = nullptr, *leaves = nullptr, *digests = nullptr, *cap = nullptr;
  size_t coeffs_bytes = 0, leaves_bytes = 0, digests_bytes = 0, cap_bytes = 0;
  // rows held: leaves [first_leaf, first_leaf + nleaves) (all m of them unless the context shards);
  // `cap` always has all 2^cap_height entries, the ones of other shards zero
  u64 first_leaf = 0, nleaves = 0;
  u64* own_roots() const { return cap + 4 * (first_leaf >> (log_n + rate_bits - cap_height)); }
  bool sharded() const { return nleaves != (1ULL << (log_n + rate_bits)); }
};

// Sigma polynomials' values + coset shifts of one circuit, resident in HBM (vpbs_sigmas_upload).
struct vpbs_sigmas {
  vpbs_ctx* ctx = nullptr;
  u32 num_routed = 0, log_n = 0;
  u64 *sigmas = nullptr, *k_is = nullptr;  // num_routed x n column-major; num_routed
};

// The gate constraints of one circuit as a straight-line program in HBM (vpbs_gate_program_upload;
// perm::gate_program_eval).  Owns plain device allocations on `device`.
struct vpbs_gate_program {
  int device = 0;
  u64 *code = nullptr, *imm = nullptr;
  u32 ncode = 0, nimm = 0, nregs = 0, num_constraints = 0;
  u32 max_wire = 0, max_const = 0;   // highest column indices the program reads (+ 1)
};

// FRI commit phase kept in HBM (vpbs_fri_*): the current polynomial (coefficients and coset
// evaluations over the quadratic extension, (re, im) pairs) and the Merkle tree of every layer.
struct vpbs_fri {
  vpbs_ctx* ctx = nullptr;
  u64 len = 0;             // current number of extension elements (power of two)
  u64 shift = 0;           // current coset shift
  u32 pending_arity_bits = 0;
  bool layer_open = false;  // a layer has been committed and not folded yet
  u64 *coeffs = nullptr, *values = nullptr, *scratch = nullptr;  // 2 * cap_len u64 each (+ scratch 4x)
  u64 cap_len = 0;         // allocated extension elements
  struct Layer {
    u64 nleaves = 0;
    u32 leaf_len = 0, cap_height = 0;
    u64 *leaves = nullptr, *digests = nullptr, *cap = nullptr;
    size_t leaves_bytes = 0, digests_bytes = 0, cap_bytes = 0;
  };
  std::vector<Layer> layers;
};

namespace {

int fail(vpbs_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  else {
    std::lock_guard<std::mutex> g(g_err_mu);
    g_global_err = msg;
  }
  return code;
}
#define CU(ctx, call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? VPBS_ERR_OOM : VPBS_ERR_CUDA,     \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                     \
  } while (0)

constexpr size_t POOL_LIMIT_BYTES = 8ULL << 30;  // at most 8 GiB parked in the pool
int batch_alloc(vpbs_ctx* ctx, u32 ncols, u32 log_n, u32 rate_bits, u32 cap_height, bool salted,
                bool coeff_inputs, vpbs_batch** out);

cudaError_t pool_alloc(vpbs_ctx* ctx, size_t bytes, u64** out) {
  auto it = ctx->pool.find(bytes);
  if (it != ctx->pool.end()) {
    *out = (u64*)it->second;
    ctx->pool_bytes -= bytes;
    ctx->pool.erase(it);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation && !ctx->pool.empty()) {  // give the pool back and retry
    for (auto& kv : ctx->pool) cudaFree(kv.second);
    ctx->pool.clear();
    ctx->pool_bytes = 0;
    cudaGetLastError();
    e = cudaMalloc(out, bytes);
  }
  return e;
}
void pool_free(vpbs_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  if (ctx && ctx->pool_bytes + bytes <= POOL_LIMIT_BYTES) {
    ctx->pool.emplace(bytes, p);
    ctx->pool_bytes += bytes;
  } else {
    cudaFree(p);
  }
}

int arena_get(vpbs_ctx* ctx, const char* name, size_t bytes, void** out) {
  Buf& b = ctx->arena[name];
  if (b.bytes < bytes) {
    if (b.p) {
      CU(ctx, cudaStreamSynchronize(ctx->stream));
      CU(ctx, cudaFree(b.p));
      b.p = nullptr;
      b.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    CU(ctx, cudaMalloc(&b.p, want));
    b.bytes = want;
  }
  *out = b.p;
  return VPBS_OK;
}

int ensure_roots(vpbs_ctx* ctx, unsigned log_N) {
  if (log_N < 1) log_N = 1;
  if (ctx->roots && ctx->roots_log >= log_N) return VPBS_OK;
  if (ctx->roots) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaFree(ctx->roots));
    ctx->roots = nullptr;
  }
  const u64 half = 1ULL << (log_N - 1);
  CU(ctx, cudaMalloc(&ctx->roots, half * sizeof(u64)));
  ntt::fill_roots<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>(ctx->roots, log_N);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  ctx->roots_log = log_N;
  return VPBS_OK;
}

int get_coset_table(vpbs_ctx* ctx, unsigned log_n, unsigned rate_bits, u64 shift, const u64** out) {
  auto key = std::make_pair(log_n, std::make_pair(rate_bits, shift));
  auto it = ctx->coset_tables.find(key);
  if (it != ctx->coset_tables.end()) {
    *out = it->second;
    return VPBS_OK;
  }
  if (ctx->coset_tables.size() >= 16) {  // bounded cache
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->coset_tables) cudaFree(kv.second);
    ctx->coset_tables.clear();
  }
  const u64 total = 1ULL << (log_n + rate_bits);
  u64* t = nullptr;
  CU(ctx, cudaMalloc(&t, total * sizeof(u64)));
  ntt::fill_coset_powers<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(t, log_n,
                                                                                 rate_bits, shift);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  ctx->coset_tables[key] = t;
  *out = t;
  return VPBS_OK;
}

// Pass plan: the last pass takes min(L, 8) layers, the rest is split evenly (<= 8 each).
struct Plan {
  unsigned npass = 0;
  unsigned s[8];
};
Plan make_plan(unsigned L) {
  Plan p;
  const unsigned last = L < ntt::MAX_PASS_BITS ? L : ntt::MAX_PASS_BITS;
  unsigned rest = L - last;
  const unsigned k = (rest + ntt::MAX_PASS_BITS - 1) / ntt::MAX_PASS_BITS;
  for (unsigned i = 0; i < k; i++) {
    unsigned si = rest / (k - i);
    if (rest % (k - i)) si++;
    p.s[p.npass++] = si;
    rest -= si;
  }
  p.s[p.npass++] = last;
  return p;
}

enum class Out { Leaf, Natural };

// ---- TMA tile maps for the 256-point passes (ntt::pass_*_r16t) ------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static_assert(sizeof(CUtensorMap) == sizeof(ntt::tma::TileMap) && alignof(CUtensorMap) <= alignof(ntt::tma::TileMap),
              "ntt::tma::TileMap must mirror CUtensorMap");
EncodeTiledFn tile_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// A u64 tensor of `rank` dimensions (dims[0] contiguous; strides in elements for dims 1..) with the
// given box.  false: this source cannot be described (alignment) and the caller uses the cp.async pass.
bool make_tile_map(ntt::tma::TileMap* out, const u64* base, unsigned rank, const u64* dims,
                   const u64* strides_elems, const unsigned* box) {
  EncodeTiledFn enc = tile_map_encoder();
  if (!enc || ((uintptr_t)base & 15)) return false;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (unsigned i = 0; i < rank; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i) {
      gstr[i - 1] = strides_elems[i - 1] * sizeof(u64);
      if (gstr[i - 1] & 15) return false;
    }
  }
  return enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT64, rank, const_cast<u64*>(base),
             gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// One size-2^log_n transform of `ncols` columns.
//   src (column-major, stride src_stride) -> dst.  work: scratch ncols x n (needed if npass > 1).
//   Out::Leaf   : dst = row-major matrix, row stride dst_stride, rows row0.. in bit-reversed order
//   Out::Natural: dst = column-major, column stride dst_stride, natural order
template <bool INVERSE>
int run_transform(vpbs_ctx* ctx, const u64* src, u64 src_stride, unsigned ncols, unsigned log_n,
                  u64* work, Out mode, u64* dst, u64 dst_stride, u64 row0, const u64* in_scale,
                  u64 out_scale, cudaStream_t stream = nullptr) {
  if (!stream) stream = ctx->stream;
  const ntt::Roots R{ctx->roots, ctx->roots_log};
  const Plan plan = make_plan(log_n);
  const u64 n = 1ULL << log_n;
  unsigned log_B = log_n;
  const u64* cur = src;
  u64 cur_stride = src_stride;
  // 2^16-point forward transforms into leaf rows (two 256-point passes): the four-step twiddle
  // between the passes is applied by the second pass at its load, see ntt::pass_strided_r16.
  const bool tw_at_load = !INVERSE && mode == Out::Leaf && plan.npass == 2 && plan.s[0] == 8 &&
                          plan.s[1] == 8 && log_n == 16;
  for (unsigned p = 0; p + 1 < plan.npass; p++) {
    const unsigned s = plan.s[p];
    const unsigned log_sigma = log_B - s;
    unsigned log_T = ntt::LOG_TILE - s;
    if (log_T > log_sigma) log_T = log_sigma;
    const size_t smem = ((size_t)(1u << s << log_T) + (1u << s) / 2) * sizeof(u64);
    dim3 grid((unsigned)(n >> (s + log_T)), ncols);
    const unsigned ntiles = grid.x * grid.y;
    const unsigned pgrid = ntiles < ctx->sms * ntt::R16P_MIN_BLOCKS ? ntiles : ctx->sms * ntt::R16P_MIN_BLOCKS;
    ntt::tma::TileMap map;
    bool tiled = false;
    if (s == 8 && log_T == 4) {  // the tile is a 16 x 256 box of (low, q, block, column)
      const u64 dims[4] = {1ULL << log_sigma, 256, n >> log_B, ncols};
      const u64 strides[3] = {1ULL << log_sigma, 1ULL << log_B, cur_stride};
      const unsigned box[4] = {16, 256, 1, 1};
      tiled = make_tile_map(&map, cur, 4, dims, strides, box);
    }
    if (tiled && tw_at_load)  // 256-point pass: persistent radix-16 register kernel, TMA tile loads
      ntt::pass_strided_r16t<INVERSE, false><<<pgrid, ntt::THREADS, ntt::R16T_STRIDED_SMEM, stream>>>(
          map, work, n, log_B, p == 0 ? in_scale : nullptr, R, grid.x, ntiles);
    else if (tiled)
      ntt::pass_strided_r16t<INVERSE, true><<<pgrid, ntt::THREADS, ntt::R16T_STRIDED_SMEM, stream>>>(
          map, work, n, log_B, p == 0 ? in_scale : nullptr, R, grid.x, ntiles);
    else if (s == 8 && log_T == 4 && tw_at_load)  // sources TMA cannot describe: cp.async tile loads
      ntt::pass_strided_r16p<INVERSE, false><<<pgrid, ntt::THREADS, ntt::R16P_STRIDED_SMEM, stream>>>(
          cur, cur_stride, work, n, log_B, p == 0 ? in_scale : nullptr, R, grid.x, ntiles);
    else if (s == 8 && log_T == 4)
      ntt::pass_strided_r16p<INVERSE, true><<<pgrid, ntt::THREADS, ntt::R16P_STRIDED_SMEM, stream>>>(
          cur, cur_stride, work, n, log_B, p == 0 ? in_scale : nullptr, R, grid.x, ntiles);
    else
      ntt::pass_strided<INVERSE><<<grid, ntt::THREADS, smem, stream>>>(
          cur, cur_stride, work, n, log_B, s, log_T, p == 0 ? in_scale : nullptr, R);
    ctx->launches++;
    cur = work;
    cur_stride = n;
    log_B -= s;
  }
  {
    const unsigned s = plan.s[plan.npass - 1];
    const u64* scale = plan.npass == 1 ? in_scale : nullptr;
    if (mode == Out::Leaf) {
      const unsigned log_T = 4;
      const size_t smem = ((size_t)(1u << s) * ((1u << log_T) + 1) + (1u << s) / 2) * sizeof(u64);
      dim3 grid((unsigned)(n >> s), (ncols + (1u << log_T) - 1) >> log_T);
      const unsigned ntiles = grid.x * grid.y;
      const unsigned pgrid = ntiles < ctx->sms * ntt::R16P_MIN_BLOCKS ? ntiles : ctx->sms * ntt::R16P_MIN_BLOCKS;
      ntt::tma::TileMap map;
      bool tiled = false;
      if (s == 8 && !scale && ncols >= 16) {  // the tile is a 256 x 16 box of (position, column)
        const u64 dims[2] = {n, ncols};
        const u64 strides[1] = {cur_stride};
        const unsigned box[2] = {256, 16};
        tiled = make_tile_map(&map, cur, 2, dims, strides, box);
      }
      if (tiled)
        ntt::pass_final_r16t<INVERSE, ntt::STORE_LEAF><<<pgrid, ntt::THREADS, ntt::R16T_FINAL_SMEM, stream>>>(
            map, ncols, dst, dst_stride, row0, log_n, out_scale, R, tw_at_load ? log_n : 0u,
            tw_at_load ? in_scale : nullptr, grid.x, ntiles);
      else if (s == 8)
        ntt::pass_final_r16p<INVERSE, ntt::STORE_LEAF><<<pgrid, ntt::THREADS, ntt::R16P_FINAL_SMEM, stream>>>(
            cur, cur_stride, ncols, dst, dst_stride, row0, log_n, scale, out_scale, R,
            tw_at_load ? log_n : 0u, tw_at_load ? in_scale : nullptr, grid.x, ntiles);
      else
        ntt::pass_final<INVERSE, ntt::STORE_LEAF><<<grid, ntt::THREADS, smem, stream>>>(
            cur, cur_stride, ncols, dst, dst_stride, row0, log_n, s, log_T, scale, out_scale, R);
    } else {
      unsigned log_T = 4;
      if (log_T > log_n - s) log_T = log_n - s;
      const size_t smem = ((size_t)(1u << s) * ((1u << log_T) + 1) + (1u << s) / 2) * sizeof(u64);
      dim3 grid((unsigned)(n >> (s + log_T)), ncols);
      const unsigned ntiles = grid.x * grid.y;
      const unsigned pgrid = ntiles < ctx->sms * ntt::R16P_MIN_BLOCKS ? ntiles : ctx->sms * ntt::R16P_MIN_BLOCKS;
      ntt::tma::TileMap map;
      bool tiled = false;
      if (s == 8 && log_T == 4 && !scale) {  // 16 blocks 2^(log_nb - 4) apart, 256 positions each
        const unsigned log_nb = log_n - 8;
        const u64 dims[4] = {256, 1ULL << (log_nb - 4), 16, ncols};
        const u64 strides[3] = {256, 256ULL << (log_nb - 4), cur_stride};
        const unsigned box[4] = {256, 1, 16, 1};
        tiled = make_tile_map(&map, cur, 4, dims, strides, box);
      }
      if (tiled)
        ntt::pass_final_r16t<INVERSE, ntt::STORE_NATURAL><<<pgrid, ntt::THREADS, ntt::R16T_FINAL_SMEM, stream>>>(
            map, ncols, dst, dst_stride, row0, log_n, out_scale, R, 0u, nullptr, grid.x, ntiles);
      else if (s == 8 && log_T == 4)
        ntt::pass_final_r16p<INVERSE, ntt::STORE_NATURAL><<<pgrid, ntt::THREADS, ntt::R16P_FINAL_SMEM, stream>>>(
            cur, cur_stride, ncols, dst, dst_stride, row0, log_n, scale, out_scale, R, 0u, nullptr,
            grid.x, ntiles);
      else
        ntt::pass_final<INVERSE, ntt::STORE_NATURAL><<<grid, ntt::THREADS, smem, stream>>>(
            cur, cur_stride, ncols, dst, dst_stride, row0, log_n, s, log_T, scale, out_scale, R);
    }
    ctx->launches++;
  }
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}

int log2_strict(u64 n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1ULL << l) < n) l++;
  return l;
}

// MerkleTree::new over device leaves, in two steps so that callers can put them on different
// streams.  nleaves leaves of `width`, split into subtrees of 2^log_sub leaves whose roots go to
// d_roots; digests in plonky2 layout per subtree.
constexpr unsigned COOP_LOG = 12;  // levels with <= 2^12 nodes: one 16-thread group per node
constexpr unsigned FUSE_LOG = 4;   // <= 2^4 nodes per subtree: all remaining levels in one launch
// (both thresholds swept on B200 in round 1: 2^12 nodes and 16 nodes per subtree are the minima)

int merkle_leaves(vpbs_ctx* ctx, const u64* d_leaves, u64 nleaves, u32 width, unsigned log_sub,
                  u64* d_digests, u64* d_roots, cudaStream_t st) {
  if (nleaves == 0) return VPBS_OK;
  const u64 sub_digests = 2 * (1ULL << log_sub) - 2;
  const int all_cap = log_sub == 0;
  merkle::hash_leaves<<<(unsigned)((nleaves + 127) / 128), 128, 0, st>>>(
      d_leaves, nleaves, width, all_cap ? d_roots : d_digests, log_sub, sub_digests, all_cap);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}

// Levels 1..log_sub above the leaf digests of nleaves >> log_sub subtrees.
int merkle_levels(vpbs_ctx* ctx, u64 nleaves, unsigned log_sub, u64* d_digests, u64* d_roots,
                  cudaStream_t st) {
  if (nleaves == 0 || log_sub == 0) return VPBS_OK;
  const u64 sub_leaves = 1ULL << log_sub;
  const u64 nsub = nleaves >> log_sub;
  const u64 sub_digests = 2 * sub_leaves - 2;
  // Big levels: one thread per node, one launch per level.  Levels with at most 4096 nodes in
  // total use one 16-thread group per node (lower latency); once a subtree is down to 16 nodes
  // the rest of the tree runs in ONE launch (one CTA per subtree, no launch gaps).
  unsigned total_log = 0;
  while ((1ULL << total_log) < nleaves) total_log++;
  const unsigned coop_from = total_log > COOP_LOG ? total_log - COOP_LOG : 1;  // first level with <= 2^COOP_LOG nodes
  unsigned fused_from = log_sub > FUSE_LOG ? log_sub - FUSE_LOG : 1;  // <= 2^FUSE_LOG nodes per subtree
  if (fused_from < coop_from) fused_from = coop_from;
  const bool fuse = nsub <= 4096;
  for (unsigned level = 1; level <= log_sub; level++) {
    const u64 nnodes = nsub << (log_sub - level);
    if (fuse && level >= fused_from) {
      const u64 nodes0 = 1ULL << (log_sub - level);
      unsigned threads = (unsigned)(nodes0 * 16 < 32 ? 32 : nodes0 * 16);
      if (threads > merkle::COOP_THREADS) threads = merkle::COOP_THREADS;
      const size_t smem = (size_t)(threads / 16) * 48 * sizeof(double);
      merkle::reduce_levels_coop<<<(unsigned)nsub, threads, smem, st>>>(d_digests, d_roots, level,
                                                                       log_sub, sub_digests);
      ctx->launches++;
      break;
    }
    if (level >= coop_from)
      merkle::reduce_level_coop<<<(unsigned)((nnodes + 7) / 8), 128, 0, st>>>(
          d_digests, d_roots, level, log_sub, sub_digests, nnodes);
    else
      merkle::reduce_level<<<(unsigned)((nnodes + 127) / 128), 128, 0, st>>>(
          d_digests, d_roots, level, log_sub, sub_digests, nnodes);
    ctx->launches++;
  }
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}

int merkle_build(vpbs_ctx* ctx, const u64* d_leaves, u64 nleaves, u32 width, unsigned log_sub,
                 u64* d_digests, u64* d_roots, cudaEvent_t after_leaves = nullptr) {
  int rc = merkle_leaves(ctx, d_leaves, nleaves, width, log_sub, d_digests, d_roots, ctx->stream);
  if (rc) return rc;
  if (after_leaves) cudaEventRecord(after_leaves, ctx->stream);
  return merkle_levels(ctx, nleaves, log_sub, d_digests, d_roots, ctx->stream);
}

// Points of the commit the host API hangs its overlapped output copies on.
struct Overlap {
  cudaEvent_t coeffs_ready = nullptr;      // after the IFFT
  std::vector<cudaEvent_t> block_ready;    // after LDE block b (its n leaf rows are final)
  // Column-chunked mode (host API, wide batches): the inputs of chunk k arrive on another stream
  // (wait h2d_ready[k]); chunk k's coefficients are final at coeffs_chunk_ready[k]; the leaf
  // matrix is final at lde_done.  chunk_cols == 0: all columns at once (events above).
  u32 chunk_cols = 0;
  bool lde_by_block = false;  // chunked inputs, but LDE over all columns block by block (see commit_core)
  std::vector<cudaEvent_t> h2d_ready, coeffs_chunk_ready;
  cudaEvent_t lde_done = nullptr;
  // Staged upload in flight (pageable host columns): h2d_ready[k] is recorded by the uploader thread,
  // and a stream wait enqueued before the record would wait for nothing — so the enqueueing thread
  // first waits on the host until the record has happened.
  hoststage::Upload* upload = nullptr;
  // Hash every column chunk into the leaf sponges right after its LDE (merkle::hash_leaves_part)
  // instead of hashing whole rows at the end: only worth it when the chunks arrive slowly (staged
  // uploads from pageable memory) — with pinned inputs it measured 0.1 ms slower.
  bool absorb_chunks = false;
};

struct Timer {
  vpbs_ctx* ctx;
  bool on;
  int n = 0;
  bool leaf_event = false;
  Overlap* overlap = nullptr;
  void mark() {
    if (on && n < 4) cudaEventRecord(ctx->ev[n++], ctx->stream);
  }
  float ms(int a, int b) {
    float t = 0;
    if (on && a < n && b < n) cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
    return t;
  }
};

// The commit on device data; shard = LDE blocks [first_leaf/n, +nleaves_shard/n).
int commit_core(vpbs_ctx* ctx, const u64* d_cols, u32 ncols, u32 log_n, u32 rate_bits,
                u32 cap_height, int inputs_are_coeffs, const u64* d_salt, u64 first_leaf,
                u64 nleaves_shard, u64* d_coeffs_out, u64* d_leaves, u64* d_digests, u64* d_roots,
                Timer* tm) {
  if (!d_cols || !d_roots || ncols == 0) return fail(ctx, VPBS_ERR_ARG, "null pointer or ncols == 0");
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  const unsigned log_m = log_n + rate_bits;
  if (cap_height > log_m)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  const u32 width = ncols + (d_salt ? VPBS_SALT_SIZE : 0);
  if (nleaves_shard == 0 || (nleaves_shard & (n - 1)) || (first_leaf & (n - 1)) ||
      first_leaf + nleaves_shard > m)
    return fail(ctx, VPBS_ERR_ARG, "shard must be a whole number of n-row LDE blocks");
  const unsigned log_sub = log_m - cap_height;  // leaves per cap subtree
  const bool whole = nleaves_shard == m;
  if (!whole) {
    const int lg = log2_strict(nleaves_shard);
    if (lg < 0 || (first_leaf & (nleaves_shard - 1)) || (unsigned)lg < log_sub)
      return fail(ctx, VPBS_ERR_ARG,
                  "shard must be an aligned power of two covering whole cap subtrees");
  }
  if (!d_leaves) return fail(ctx, VPBS_ERR_ARG, "leaves buffer required");
  if (log_sub > 0 && !d_digests) return fail(ctx, VPBS_ERR_ARG, "digests buffer required");

  int rc;
  if ((rc = ensure_roots(ctx, log_m)) != VPBS_OK) return rc;
  const u64* coset = nullptr;
  if ((rc = get_coset_table(ctx, log_n, rate_bits, gl::COSET_SHIFT, &coset)) != VPBS_OK) return rc;
  u64 *work = nullptr, *work2 = nullptr;
  if ((rc = arena_get(ctx, "work", (size_t)ncols * n * sizeof(u64), (void**)&work)) != VPBS_OK)
    return rc;
  const bool two_streams = (nleaves_shard >> log_n) > 1;
  if (two_streams &&
      (rc = arena_get(ctx, "work2", (size_t)ncols * n * sizeof(u64), (void**)&work2)) != VPBS_OK)
    return rc;

  tm->mark();  // 0
  const u64 b0 = first_leaf >> log_n, nb = nleaves_shard >> log_n;
  const u64 n_inv = gl::inv(n % gl::P);
  u64* cbuf = d_coeffs_out;
  if (!inputs_are_coeffs && !cbuf &&
      (rc = arena_get(ctx, "coeffs", (size_t)ncols * n * sizeof(u64), (void**)&cbuf)))
    return rc;
  Overlap* ov = tm->overlap;
  const u32 chunk = (ov && ov->chunk_cols && ov->chunk_cols < ncols) ? ov->chunk_cols : ncols;
  const bool chunked = chunk < ncols;
  if (chunked) tm->mark();  // 1: pipelined mode reports all transforms under "FFT + blinding"
  // Chunked inputs, two orders for the LDE: per column chunk right after its IFFT (the GPU starts
  // early: best when little is copied back), or — by_block — once all coefficients are there,
  // block by block over all columns, so that finished leaf rows can stream out while later
  // blocks are still being transformed (best when the leaf matrix is copied to the host).
  const bool by_block = chunked && ov->lde_by_block;
  // chunk-wise hashing needs whole permutations per chunk (chunk a multiple of the rate) and rows
  // without salt columns
  const bool absorb = chunked && !by_block && ov->absorb_chunks && !d_salt && chunk % poseidon::RATE == 0;
  // "FFT + blinding" + "transpose LDEs": one size-n coset transform per LDE block, written as
  // leaf rows (columns c0 .. c0 + nc of the row-major matrix).
  auto lde_blocks = [&](const u64* coeffs, u32 c0, u32 nc, bool block_events) -> int {
    if (two_streams) {  // odd blocks on the auxiliary stream, once the coefficients are there
      CU(ctx, cudaEventRecord(ctx->aux_fork, ctx->stream));
      CU(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
    }
    for (u64 b = 0; b < nb; b++) {
      const bool aux = two_streams && (b & 1);
      cudaStream_t st = aux ? ctx->aux_stream : ctx->stream;
      int rc2 = run_transform<false>(ctx, coeffs, n, nc, log_n, aux ? work2 : work, Out::Leaf,
                                     d_leaves + c0, width, b << log_n, coset + ((b0 + b) << log_n), 1,
                                     st);
      if (rc2 != VPBS_OK) return rc2;
      if (block_events && ov && !d_salt && b < ov->block_ready.size())
        cudaEventRecord(ov->block_ready[b], st);
    }
    if (two_streams) {
      CU(ctx, cudaEventRecord(ctx->aux_join, ctx->aux_stream));
      CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));
    }
    return VPBS_OK;
  };
  for (u32 c0 = 0, k = 0; c0 < ncols; c0 += chunk, k++) {
    const u32 nc = ncols - c0 < chunk ? ncols - c0 : chunk;
    if (chunked && k < ov->h2d_ready.size()) {
      if (ov->upload) ov->upload->wait_recorded(k);
      CU(ctx, cudaStreamWaitEvent(ctx->stream, ov->h2d_ready[k], 0));
    } else if (!chunked && ov) {  // uploads on the H2D stream but a single compute chunk: wait for all
      if (ov->upload && !ov->h2d_ready.empty()) ov->upload->wait_recorded((unsigned)ov->h2d_ready.size() - 1);
      for (cudaEvent_t e : ov->h2d_ready) CU(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
    }
    // "IFFT": values -> coefficients (natural order), scaled by n^-1.
    const u64* coeffs = d_cols + (u64)c0 * n;
    if (!inputs_are_coeffs) {
      if ((rc = run_transform<true>(ctx, d_cols + (u64)c0 * n, n, nc, log_n, work, Out::Natural,
                                    cbuf + (u64)c0 * n, n, 0, nullptr, n_inv)) != VPBS_OK)
        return rc;
      coeffs = cbuf + (u64)c0 * n;
    } else if (d_coeffs_out && d_coeffs_out != d_cols) {
      CU(ctx, cudaMemcpyAsync(d_coeffs_out + (u64)c0 * n, d_cols + (u64)c0 * n,
                              (size_t)nc * n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (!chunked) {
      tm->mark();  // 1
      if (ov && ov->coeffs_ready) cudaEventRecord(ov->coeffs_ready, ctx->stream);
    } else if (k < ov->coeffs_chunk_ready.size()) {
      cudaEventRecord(ov->coeffs_chunk_ready[k], ctx->stream);
    }
    if (!by_block && (rc = lde_blocks(coeffs, c0, nc, !chunked)) != VPBS_OK) return rc;
    if (absorb) {  // this chunk's columns of every leaf row into the sponges
      u64* sponge = nullptr;
      if ((rc = arena_get(ctx, "sponge", (size_t)nleaves_shard * poseidon::WIDTH * sizeof(u64), (void**)&sponge)))
        return rc;
      const bool last = c0 + nc == ncols;
      const u64 sub_digests = 2 * (1ULL << log_sub) - 2;
      merkle::hash_leaves_part<<<(unsigned)((nleaves_shard + 127) / 128), 128, 0, ctx->stream>>>(
          d_leaves, nleaves_shard, width, c0, c0 + nc, sponge, c0 == 0, last,
          log_sub == 0 ? d_roots : d_digests, log_sub, sub_digests, log_sub == 0);
      ctx->launches++;
    }
  }
  if (by_block &&
      (rc = lde_blocks(inputs_are_coeffs ? d_cols : cbuf, 0, ncols, true)) != VPBS_OK)
    return rc;
  if (d_salt) {
    const u64 cnt = nleaves_shard * 4;
    merkle::scatter_salt<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(
        d_salt, m, log_m, first_leaf, nleaves_shard, d_leaves, width, ncols);
    ctx->launches++;
  }
  if (ov && ov->lde_done) cudaEventRecord(ov->lde_done, ctx->stream);
  tm->mark();  // 2
  // "build Merkle tree".  (Hashing every LDE block right after its transform, on the two LDE streams
  // with the tree levels on a third, was built and measured in round 2: 8.30 ms against 7.03 ms for
  // this phase-by-phase order at 2^16 x 128 — a 65,536-leaf launch is less than one wave of 128-thread
  // CTAs, each of which lives ~1 ms (16 sequential permutations), so the per-block kernels fragment
  // the machine)
  if (absorb) {  // leaf digests are final: only the levels above them remain
    if (tm->on) cudaEventRecord(ctx->ev[8], ctx->stream);
    if ((rc = merkle_levels(ctx, nleaves_shard, log_sub, d_digests, d_roots, ctx->stream)) != VPBS_OK) return rc;
  } else if ((rc = merkle_build(ctx, d_leaves, nleaves_shard, width, log_sub, d_digests, d_roots,
                                tm->on ? ctx->ev[8] : nullptr)) != VPBS_OK)
    return rc;
  tm->leaf_event = tm->on;
  tm->mark();  // 3
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}

void fill_stats(vpbs_stats* st, Timer& tm, uint64_t launches) {
  // events: 0 start | 1 after ifft | 2 after fft | 3 after merkle  (host variants add h2d/d2h)
  st->ifft_ms = tm.ms(0, 1);
  st->fft_ms = tm.ms(1, 2);
  st->merkle_ms = tm.ms(2, 3);
  if (tm.leaf_event && tm.n > 2) cudaEventElapsedTime(&st->leaf_hash_ms, tm.ctx->ev[2], tm.ctx->ev[8]);
  st->kernel_launches = launches;
}

bool usable(vpbs_ctx* ctx) { return ctx != nullptr; }

u64 reverse_bits64(u64 x, unsigned bits) {
  u64 r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1ULL) << (bits - 1 - i);
  return r;
}

// The persistent NTT passes use two tile buffers (64-68 KB of dynamic shared memory): opt in, on
// the current device, for every instantiation run_transform launches.
cudaError_t allow_large_smem() {
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* f, size_t bytes) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  };
  set((const void*)ntt::pass_strided_r16p<false, false>, ntt::R16P_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16p<false, true>, ntt::R16P_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16p<true, false>, ntt::R16P_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16p<true, true>, ntt::R16P_STRIDED_SMEM);
  set((const void*)ntt::pass_final_r16p<false, ntt::STORE_LEAF>, ntt::R16P_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16p<false, ntt::STORE_NATURAL>, ntt::R16P_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16p<true, ntt::STORE_LEAF>, ntt::R16P_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16p<true, ntt::STORE_NATURAL>, ntt::R16P_FINAL_SMEM);
  set((const void*)ntt::pass_strided_r16t<false, false>, ntt::R16T_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16t<false, true>, ntt::R16T_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16t<true, false>, ntt::R16T_STRIDED_SMEM);
  set((const void*)ntt::pass_strided_r16t<true, true>, ntt::R16T_STRIDED_SMEM);
  set((const void*)ntt::pass_final_r16t<false, ntt::STORE_LEAF>, ntt::R16T_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16t<true, ntt::STORE_LEAF>, ntt::R16T_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16t<false, ntt::STORE_NATURAL>, ntt::R16T_FINAL_SMEM);
  set((const void*)ntt::pass_final_r16t<true, ntt::STORE_NATURAL>, ntt::R16T_FINAL_SMEM);
  return e;
}

// Copies columns [c0, c1) between per-column host pointers and a column-major device buffer
// (column c at dev + c * n).  Host columns that happen to be adjacent in memory (one allocation, as
// plonky2's flattened buffers or this repo's hosts provide) travel as ONE transfer: 128 separate
// 512 KB copies per direction cost the 2^16 x 128 commit about a millisecond of DMA set-up.
cudaError_t copy_columns(u64* dev, const uint64_t* const* host, u32 c0, u32 c1, u64 n, bool to_device,
                         cudaStream_t st) {
  u32 c = c0;
  while (c < c1) {
    if (!host[c]) {  // skipped output column
      c++;
      continue;
    }
    u32 e = c + 1;
    while (e < c1 && host[e] == host[e - 1] + n) e++;
    const size_t bytes = (size_t)(e - c) * n * sizeof(u64);
    cudaError_t err = to_device
        ? cudaMemcpyAsync(dev + (u64)c * n, host[c], bytes, cudaMemcpyHostToDevice, st)
        : cudaMemcpyAsync(const_cast<uint64_t*>(host[c]), dev + (u64)c * n, bytes, cudaMemcpyDeviceToHost, st);
    if (err != cudaSuccess) return err;
    c = e;
  }
  return cudaSuccess;
}

int bind(vpbs_ctx* ctx) {
  if (!usable(ctx)) return VPBS_ERR_STATE;
  CU(ctx, cudaSetDevice(ctx->device));
  return VPBS_OK;
}

}  // namespace

extern "C" {

int vpbs_abi_version(void) { return VPBS_ABI_VERSION; }

int vpbs_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(nullptr, VPBS_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return VPBS_ERR_CUDA;
  }
  return n;
}

int vpbs_ctx_create(int device, vpbs_ctx** out) {
  if (!out) return fail(nullptr, VPBS_ERR_ARG, "out == NULL");
  *out = nullptr;
  int n = vpbs_device_count();
  if (n < 0) return n;
  if (n == 0) return fail(nullptr, VPBS_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= n) return fail(nullptr, VPBS_ERR_ARG, "device index out of range");
  vpbs_ctx* ctx = new (std::nothrow) vpbs_ctx();
  if (!ctx) return fail(nullptr, VPBS_ERR_OOM, "host allocation failed");
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming);

  for (int i = 0; e == cudaSuccess && i < 10; i++) e = cudaEventCreate(&ctx->ev[i]);
  if (e == cudaSuccess) {
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess && sms > 0) ctx->sms = (unsigned)sms;
  }
  if (e == cudaSuccess) e = allow_large_smem();
  if (e != cudaSuccess) {
    fail(nullptr, VPBS_ERR_CUDA, std::string("context setup: ") + cudaGetErrorString(e));
    delete ctx;
    return VPBS_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return VPBS_OK;
}

void vpbs_ctx_destroy(vpbs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (vpbs_fri* f : ctx->fri_chains) {  // orphan, like the batches below
    cudaFree(f->coeffs);
    cudaFree(f->values);
    cudaFree(f->scratch);
    for (auto& l : f->layers) {
      cudaFree(l.leaves);
      cudaFree(l.digests);
      cudaFree(l.cap);
    }
    f->layers.clear();
    f->coeffs = f->values = f->scratch = nullptr;
    f->ctx = nullptr;
  }
  ctx->fri_chains.clear();
  for (vpbs_sigmas* sg : ctx->sigma_sets) {  // orphan, like the batches below
    cudaFree(sg->sigmas);
    cudaFree(sg->k_is);
    sg->sigmas = sg->k_is = nullptr;
    sg->ctx = nullptr;
  }
  ctx->sigma_sets.clear();
  for (vpbs_batch* b : ctx->batches) {  // orphan: the handle stays valid but holds nothing
    cudaFree(b->coeffs);
    cudaFree(b->leaves);
    cudaFree(b->digests);
    cudaFree(b->cap);
    b->coeffs = b->leaves = b->digests = b->cap = nullptr;
    b->ctx = nullptr;
  }
  ctx->batches.clear();
  for (auto& kv : ctx->arena)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : ctx->coset_tables) cudaFree(kv.second);
  if (ctx->roots) cudaFree(ctx->roots);
  for (int i = 0; i < 10; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (auto& kv : ctx->pool) cudaFree(kv.second);
  for (cudaEvent_t e : ctx->ov) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->peer_ev) cudaEventDestroy(e);
  if (ctx->peer_tail) cudaEventDestroy(ctx->peer_tail);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);

  ctx->ring.release();
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int vpbs_ctx_set_stream(vpbs_ctx* ctx, void* cuda_stream) {
  if (!usable(ctx)) return VPBS_ERR_STATE;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return VPBS_OK;
}

int vpbs_ctx_set_host_threads(vpbs_ctx* ctx, unsigned threads) {
  if (!usable(ctx)) return VPBS_ERR_STATE;
  if (threads > 64) return fail(ctx, VPBS_ERR_ARG, "host_threads > 64");
  ctx->host_threads = threads;
  return VPBS_OK;
}

int vpbs_ctx_set_shard(vpbs_ctx* ctx, uint32_t index, uint32_t count) {
  if (!usable(ctx)) return VPBS_ERR_STATE;
  if (count == 0 || (count & (count - 1)) || index >= count)
    return fail(ctx, VPBS_ERR_ARG, "shard count must be a power of two and index < count");
  ctx->shard_index = index;
  ctx->shard_count = count;
  return VPBS_OK;
}

int vpbs_ctx_sync(vpbs_ctx* ctx) {
  int rc = bind(ctx);
  if (rc) return rc;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

const char* vpbs_last_error(vpbs_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> g(g_err_mu);
  static thread_local std::string copy;
  copy = g_global_err;
  return copy.c_str();
}

uint64_t vpbs_ctx_kernel_launches(vpbs_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* vpbs_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void vpbs_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- single-vector transforms --------------------------------------------------------------------
static int transform_host(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n, bool inverse, bool coset,
                          uint64_t shift) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!inout) return fail(ctx, VPBS_ERR_ARG, "inout == NULL");
  if (log_n > 30) return fail(ctx, VPBS_ERR_ARG, "log_n > 30");
  const u64 n = 1ULL << log_n;
  u64 *a = nullptr, *b = nullptr, *w = nullptr;
  if ((rc = arena_get(ctx, "in", n * sizeof(u64), (void**)&a))) return rc;
  if ((rc = arena_get(ctx, "coeffs", n * sizeof(u64), (void**)&b))) return rc;
  if ((rc = arena_get(ctx, "work", n * sizeof(u64), (void**)&w))) return rc;
  if ((rc = ensure_roots(ctx, log_n))) return rc;
  const u64* scale = nullptr;
  if (coset && (rc = get_coset_table(ctx, log_n, 0, shift, &scale))) return rc;
  CU(ctx, cudaMemcpyAsync(a, inout, n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
  if (inverse)
    rc = run_transform<true>(ctx, a, n, 1, log_n, w, Out::Natural, b, n, 0, nullptr, gl::inv(n % gl::P));
  else
    rc = run_transform<false>(ctx, a, n, 1, log_n, w, Out::Natural, b, n, 0, scale, 1);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(inout, b, n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}
int vpbs_fft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n) {
  return transform_host(ctx, inout, log_n, false, false, 1);
}
int vpbs_ifft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n) {
  return transform_host(ctx, inout, log_n, true, false, 1);
}
int vpbs_coset_fft(vpbs_ctx* ctx, uint64_t* inout, uint32_t log_n, uint64_t shift) {
  return transform_host(ctx, inout, log_n, false, true, shift);
}

// ---- hashing -------------------------------------------------------------------------------------
int vpbs_poseidon_permute(vpbs_ctx* ctx, uint64_t* states, uint64_t count) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!states) return fail(ctx, VPBS_ERR_ARG, "states == NULL");
  u64* d = nullptr;
  const size_t bytes = count * 12 * sizeof(u64);
  if ((rc = arena_get(ctx, "in", bytes, (void**)&d))) return rc;
  CU(ctx, cudaMemcpyAsync(d, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
  merkle::permute_batch<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d, count);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(states, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_hash_or_noop_batch(vpbs_ctx* ctx, const uint64_t* rows, uint64_t count, uint32_t len,
                            uint64_t* hashes_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!hashes_out || (len && !rows)) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  u64 *d = nullptr, *h = nullptr;
  const size_t bytes = count * (size_t)len * sizeof(u64);
  if ((rc = arena_get(ctx, "leaves", bytes, (void**)&d))) return rc;
  if ((rc = arena_get(ctx, "cap", count * 32, (void**)&h))) return rc;
  if (bytes) CU(ctx, cudaMemcpyAsync(d, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
  merkle::hash_leaves<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d, count, len, h, 0,
                                                                               0, 1);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(hashes_out, h, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_two_to_one_batch(vpbs_ctx* ctx, const uint64_t* left, const uint64_t* right,
                          uint64_t count, uint64_t* hashes_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!left || !right || !hashes_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  u64 *l = nullptr, *r = nullptr, *h = nullptr;
  if ((rc = arena_get(ctx, "in", count * 32, (void**)&l))) return rc;
  if ((rc = arena_get(ctx, "coeffs", count * 32, (void**)&r))) return rc;
  if ((rc = arena_get(ctx, "cap", count * 32, (void**)&h))) return rc;
  CU(ctx, cudaMemcpyAsync(l, left, count * 32, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(r, right, count * 32, cudaMemcpyHostToDevice, ctx->stream));
  merkle::two_to_one_batch<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(l, r, count, h);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(hashes_out, h, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// ---- MerkleTree::new -----------------------------------------------------------------------------
int vpbs_merkle_new(vpbs_ctx* ctx, const uint64_t* leaves, uint64_t nleaves, uint32_t leaf_len,
                    uint32_t cap_height, uint64_t* digests_out, uint64_t* cap_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  const int lg = log2_strict(nleaves);
  if (lg < 0) return fail(ctx, VPBS_ERR_ARG, "leaves.len() must be a power of two");
  if ((int)cap_height > lg)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  if (!cap_out || (leaf_len && !leaves)) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const u64 ncap = 1ULL << cap_height, ndig = 2 * (nleaves - ncap);
  if (ndig && !digests_out) return fail(ctx, VPBS_ERR_ARG, "digests_out == NULL");
  u64 *dl = nullptr, *dd = nullptr, *dc = nullptr;
  const size_t lbytes = nleaves * (size_t)leaf_len * sizeof(u64);
  if ((rc = arena_get(ctx, "leaves", lbytes, (void**)&dl))) return rc;
  if ((rc = arena_get(ctx, "digests", ndig * 32, (void**)&dd))) return rc;
  if ((rc = arena_get(ctx, "cap", ncap * 32, (void**)&dc))) return rc;
  if (lbytes) CU(ctx, cudaMemcpyAsync(dl, leaves, lbytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = merkle_build(ctx, dl, nleaves, leaf_len, (unsigned)lg - cap_height, dd, dc))) return rc;
  if (ndig) CU(ctx, cudaMemcpyAsync(digests_out, dd, ndig * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(cap_out, dc, ncap * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_merkle_new_dev(vpbs_ctx* ctx, const uint64_t* d_leaves, uint64_t nleaves, uint32_t leaf_len,
                        uint32_t cap_height, uint64_t* d_digests_out, uint64_t* d_cap_out,
                        vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  const int lg = log2_strict(nleaves);
  if (lg < 0) return fail(ctx, VPBS_ERR_ARG, "leaves.len() must be a power of two");
  if ((int)cap_height > lg)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  if (!d_cap_out || (leaf_len && !d_leaves)) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  if ((1ULL << cap_height) < nleaves && !d_digests_out)
    return fail(ctx, VPBS_ERR_ARG, "digests buffer required");
  const uint64_t l0 = ctx->launches;
  if (stats) cudaEventRecord(ctx->ev[0], ctx->stream);
  if ((rc = merkle_build(ctx, d_leaves, nleaves, leaf_len, (unsigned)lg - cap_height, d_digests_out,
                         d_cap_out, stats ? ctx->ev[8] : nullptr)))
    return rc;
  if (stats) {
    cudaEventRecord(ctx->ev[1], ctx->stream);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memset(stats, 0, sizeof *stats);
    cudaEventElapsedTime(&stats->leaf_hash_ms, ctx->ev[0], ctx->ev[8]);
    cudaEventElapsedTime(&stats->merkle_ms, ctx->ev[0], ctx->ev[1]);
    stats->total_ms = stats->merkle_ms;
    stats->kernel_launches = ctx->launches - l0;
  }
  return VPBS_OK;
}

// ---- PolynomialBatch::lde_values -------------------------------------------------------------------
int vpbs_lde_batch(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                   uint32_t rate_bits, int inputs_are_coeffs, uint64_t* const* coeffs_out,
                   uint64_t* lde_cols_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!cols || ncols == 0 || !lde_cols_out) return fail(ctx, VPBS_ERR_ARG, "null pointer or ncols == 0");
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  const unsigned log_m = log_n + rate_bits;
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  u64 *din = nullptr, *dco = nullptr, *dw = nullptr, *dpad = nullptr, *dout = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)ncols * n * 8, (void**)&din))) return rc;
  if ((rc = arena_get(ctx, "coeffs", (size_t)ncols * n * 8, (void**)&dco))) return rc;
  if ((rc = arena_get(ctx, "work", (size_t)ncols * m * 8, (void**)&dw))) return rc;
  if ((rc = arena_get(ctx, "pad", (size_t)ncols * m * 8, (void**)&dpad))) return rc;
  if ((rc = arena_get(ctx, "leaves", (size_t)ncols * m * 8, (void**)&dout))) return rc;
  if ((rc = ensure_roots(ctx, log_m))) return rc;
  const u64* scale = nullptr;
  if ((rc = get_coset_table(ctx, log_m, 0, gl::COSET_SHIFT, &scale))) return rc;
  for (u32 c = 0; c < ncols; c++) {
    if (!cols[c]) return fail(ctx, VPBS_ERR_ARG, "cols[c] == NULL");
    CU(ctx, cudaMemcpyAsync(din + (u64)c * n, cols[c], n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  const u64* coeffs = din;
  if (!inputs_are_coeffs) {
    if ((rc = run_transform<true>(ctx, din, n, ncols, log_n, dw, Out::Natural, dco, n, 0, nullptr,
                                  gl::inv(n % gl::P))))
      return rc;
    coeffs = dco;
  }
  // PolynomialCoeffs::lde: zero padding made explicit, then one natural-order size-m coset fft.
  CU(ctx, cudaMemsetAsync(dpad, 0, (size_t)ncols * m * 8, ctx->stream));
  CU(ctx, cudaMemcpy2DAsync(dpad, m * 8, coeffs, n * 8, n * 8, ncols, cudaMemcpyDeviceToDevice,
                            ctx->stream));
  if ((rc = run_transform<false>(ctx, dpad, m, ncols, log_m, dw, Out::Natural, dout, m, 0, scale, 1)))
    return rc;
  if (coeffs_out)
    for (u32 c = 0; c < ncols; c++)
      if (coeffs_out[c])
        CU(ctx, cudaMemcpyAsync(coeffs_out[c], coeffs + (u64)c * n, n * 8, cudaMemcpyDeviceToHost,
                                ctx->stream));
  CU(ctx, cudaMemcpyAsync(lde_cols_out, dout, (size_t)ncols * m * 8, cudaMemcpyDeviceToHost,
                          ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// ---- PolynomialBatch::from_values / from_coeffs ------------------------------------------------------
int vpbs_commit_shard_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                          uint64_t first_leaf, uint64_t nleaves_shard, uint64_t* d_coeffs_out,
                          uint64_t* d_leaves_out, uint64_t* d_digests_out, uint64_t* d_roots_out,
                          vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  const uint64_t l0 = ctx->launches;
  Timer tm{ctx, stats != nullptr};
  rc = commit_core(ctx, d_cols, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, nullptr,
                   first_leaf, nleaves_shard, d_coeffs_out, d_leaves_out, d_digests_out,
                   d_roots_out, &tm);
  if (rc) return rc;
  if (stats) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    stats->total_ms = tm.ms(0, 3);
  }
  return VPBS_OK;
}

int vpbs_commit_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                    uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                    const uint64_t* d_salt, uint64_t* d_coeffs_out, uint64_t* d_leaves_out,
                    uint64_t* d_digests_out, uint64_t* d_cap_out, vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  const uint64_t l0 = ctx->launches;
  Timer tm{ctx, stats != nullptr};
  rc = commit_core(ctx, d_cols, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, d_salt, 0,
                   1ULL << (log_n + rate_bits), d_coeffs_out, d_leaves_out, d_digests_out,
                   d_cap_out, &tm);
  if (rc) return rc;
  if (stats) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    stats->total_ms = tm.ms(0, 3);
  }
  return VPBS_OK;
}

}  // extern "C"

namespace {

// Width of the column chunks wide host batches travel in (0: one piece).
u32 host_chunk_cols(u32 ncols, u32 log_n) { return (ncols >= 64 && log_n >= 12) ? 32u : 0u; }

// Host columns in ordinary (pageable) memory, and enough of them to matter: the library stages them
// through its own pinned ring (host_stage.h) instead of leaving it to the driver.
bool want_staging(vpbs_ctx* ctx, const uint64_t* const* cols, u32 ncols, u64 n) {
  if (ctx->host_threads == 0 || (u64)ncols * n * sizeof(u64) < (1u << 20)) return false;
  return hoststage::is_pageable(cols[0]) && hoststage::is_pageable(cols[ncols - 1]);
}
// Starts the uploader thread for columns [0, ncols) in chunks of chunk_cols (0: one chunk); chunk k's
// event is ready[k].
int start_staged_upload(vpbs_ctx* ctx, u64* din, const uint64_t* const* cols, u32 ncols, u64 n,
                        u32 chunk_cols, const std::vector<cudaEvent_t>& ready, cudaStream_t hs,
                        cudaEvent_t done_ev, std::unique_ptr<hoststage::Upload>* out) {
  CU(ctx, ctx->ring.ensure(ctx->host_threads));
  std::unique_ptr<hoststage::Upload> up(new hoststage::Upload());
  up->device = ctx->device;
  up->ring = &ctx->ring;
  up->stream = hs;
  up->dev = din;
  up->cols = cols;
  up->n = n;
  up->ready = ready;
  up->done_ev = done_ev;
  for (u32 c0 = 0; c0 < ncols;) {
    const u32 c1 = (chunk_cols && c0 + chunk_cols < ncols) ? c0 + chunk_cols : ncols;
    up->chunks.push_back({c0, c1});
    c0 = c1;
  }
  up->start();
  *out = std::move(up);
  return VPBS_OK;
}

// State between commit_host_enqueue and commit_host_finish.
struct HostRun {
  Timer tm{nullptr, false};
  bool chunked = false;
  uint64_t l0 = 0;
  std::unique_ptr<hoststage::Upload> upload;  // staged upload in flight (joined by commit_host_finish)
};

// One commit — or the row-range shard [first_leaf, first_leaf + nleaves_shard) of one — from host
// buffers: uploads, kernels and downloads are all ENQUEUED here and nothing is waited for, so that
// a caller can start several devices before finishing any (vpbs_commit_multi).  This device
// downloads coefficient columns [cc0, cc1), its leaf rows, and the digests / cap entries of the cap
// subtrees it owns; the *_out pointers address the shard's own part of the caller's buffers.
int commit_host_enqueue(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                        uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                        const uint64_t* const* salt_cols, u64 first_leaf, u64 nleaves_shard,
                        uint64_t* const* coeffs_out, u32 cc0, u32 cc1, uint64_t* leaves_out,
                        uint64_t* digests_out, uint64_t* roots_out, bool want_stats, HostRun* run,
                        const std::vector<cudaEvent_t>* delivered = nullptr) {
  int rc;
  const unsigned log_m = log_n + rate_bits;
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  const u32 width = ncols + (salt_cols ? VPBS_SALT_SIZE : 0);
  const unsigned log_sub = log_m - cap_height;
  const u64 nroots = (nleaves_shard >> log_sub) ? (nleaves_shard >> log_sub) : 1;
  const u64 ndig = 2 * (nleaves_shard - nroots);
  u64 *din = nullptr, *dco = nullptr, *dle = nullptr, *ddi = nullptr, *dca = nullptr, *dsa = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)ncols * n * 8, (void**)&din))) return rc;
  if ((rc = arena_get(ctx, "coeffs", (size_t)ncols * n * 8, (void**)&dco))) return rc;
  if ((rc = arena_get(ctx, "leaves", (size_t)nleaves_shard * width * 8, (void**)&dle))) return rc;
  if ((rc = arena_get(ctx, "digests", ndig * 32, (void**)&ddi))) return rc;
  if ((rc = arena_get(ctx, "cap", nroots * 32, (void**)&dca))) return rc;
  if (salt_cols && (rc = arena_get(ctx, "salt", (size_t)4 * m * 8, (void**)&dsa))) return rc;

  run->l0 = ctx->launches;
  cudaEvent_t e0 = ctx->ev[4], e1 = ctx->ev[5], e2 = ctx->ev[6], e3 = ctx->ev[7];
  for (u32 c = 0; c < ncols; c++)
    if (!cols[c]) return fail(ctx, VPBS_ERR_ARG, "cols[c] == NULL");
  // Wide batches are pipelined by column chunk: chunk k's inputs travel on the H2D stream while
  // chunk k-1 is already being transformed.
  // chunk width swept with tools/e2e_sweep.py (2^16 x 128, e2e ms): 8 -> 12.21, 16 -> 12.19,
  // 32 -> 12.08, 64 -> 12.35
  const u32 chunk_cols = host_chunk_cols(ncols, log_n);
  const u32 nchunks = chunk_cols ? (ncols + chunk_cols - 1) / chunk_cols : 0;
  const u64 nblocks = nleaves_shard >> log_n;
  while (ctx->ov.size() < nblocks + 2 * (u64)nchunks + 4) {
    cudaEvent_t e;
    CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ov.push_back(e);
  }
  Overlap ovl;
  size_t evi = 0;
  ovl.coeffs_ready = ctx->ov[evi++];
  cudaEvent_t all_done = ctx->ov[evi++];
  ovl.lde_done = ctx->ov[evi++];
  ovl.block_ready.assign(ctx->ov.begin() + evi, ctx->ov.begin() + evi + nblocks);
  evi += nblocks;
  if (nchunks) {
    ovl.chunk_cols = chunk_cols;
    ovl.lde_by_block = leaves_out != nullptr;
    ovl.h2d_ready.assign(ctx->ov.begin() + evi, ctx->ov.begin() + evi + nchunks);
    evi += nchunks;
    ovl.coeffs_chunk_ready.assign(ctx->ov.begin() + evi, ctx->ov.begin() + evi + nchunks);
  }
  // vpbs_commit_multi: the columns are being delivered into `din` by other streams (one upload per
  // chunk by its owner GPU, then peer copies); `delivered` holds one event per chunk (or a single
  // one) in place of this context's own uploads
  if (delivered) ovl.h2d_ready = *delivered;
  const bool staged = !delivered && want_staging(ctx, cols, ncols, n);
  if (staged && !nchunks) ovl.h2d_ready.assign(1, ctx->ov[evi]);  // one upload event, one compute chunk
  run->chunked = nchunks != 0 || staged;
  cudaStream_t hs = run->chunked ? ctx->h2d_stream : ctx->stream;
  if (want_stats) cudaEventRecord(e0, ctx->stream);
  if (run->chunked) {  // the H2D stream starts after whatever the caller queued on the compute stream
    CU(ctx, cudaEventRecord(all_done, ctx->stream));
    CU(ctx, cudaStreamWaitEvent(hs, all_done, 0));
  }
  if (salt_cols)
    for (int s = 0; s < VPBS_SALT_SIZE; s++) {
      if (!salt_cols[s]) return fail(ctx, VPBS_ERR_ARG, "salt_cols[s] == NULL");
      CU(ctx, cudaMemcpyAsync(dsa + (u64)s * m, salt_cols[s], m * 8, cudaMemcpyHostToDevice,
                              ctx->stream));
    }
  if (staged) {
    if ((rc = start_staged_upload(ctx, din, cols, ncols, n, chunk_cols, ovl.h2d_ready, hs,
                                  want_stats ? e1 : nullptr, &run->upload)))
      return rc;
    ovl.upload = run->upload.get();
    ovl.absorb_chunks = !ovl.lde_by_block;
  }
  for (u32 c0 = 0, k = 0; c0 < ncols && !delivered && !staged; k++) {
    const u32 c1 = (nchunks && c0 + chunk_cols < ncols) ? c0 + chunk_cols : ncols;
    CU(ctx, copy_columns(din, cols, c0, c1, n, true, hs));
    if (nchunks) CU(ctx, cudaEventRecord(ovl.h2d_ready[k], hs));
    c0 = c1;
  }
  if (want_stats && !staged) cudaEventRecord(e1, hs);
  // Output copies run on the copy stream as soon as their data is final: coefficients after the
  // IFFT, leaf rows after their last NTT pass, digests and cap after the tree.  All kernels are
  // enqueued first, so the copies overlap the remaining NTT passes and the hashing.
  run->tm = Timer{ctx, want_stats};
  run->tm.overlap = &ovl;
  rc = commit_core(ctx, din, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, dsa, first_leaf,
                   nleaves_shard, inputs_are_coeffs ? nullptr : dco, dle, ddi, dca, &run->tm);
  run->tm.overlap = nullptr;
  if (rc) {
    if (run->upload) run->upload->join();
    cudaStreamSynchronize(hs);
    return rc;
  }
  CU(ctx, cudaEventRecord(all_done, ctx->stream));
  if (want_stats) cudaEventRecord(e2, ctx->stream);
  cudaStream_t cs = ctx->copy_stream;
  if (coeffs_out && cc0 < cc1) {
    u64* csrc = inputs_are_coeffs ? din : dco;
    if (!nchunks) CU(ctx, cudaStreamWaitEvent(cs, ovl.coeffs_ready, 0));
    for (u32 c0 = 0, k = 0; c0 < ncols; k++) {
      const u32 c1 = (nchunks && c0 + chunk_cols < ncols) ? c0 + chunk_cols : ncols;
      const u32 lo = c0 > cc0 ? c0 : cc0, hi = c1 < cc1 ? c1 : cc1;
      if (lo < hi) {
        if (nchunks) CU(ctx, cudaStreamWaitEvent(cs, ovl.coeffs_chunk_ready[k], 0));
        CU(ctx, copy_columns(csrc, coeffs_out, lo, hi, n, false, cs));
      }
      c0 = c1;
    }
  }
  if (leaves_out) {
    const size_t block_elems = (size_t)n * width;
    const bool per_block = (!nchunks || ovl.lde_by_block) && !salt_cols;
    if (!per_block)  // rows are final only after the last column chunk / the salt scatter
      CU(ctx, cudaStreamWaitEvent(cs, ovl.lde_done, 0));
    for (u64 blk = 0; blk < nblocks; blk++) {
      if (per_block) CU(ctx, cudaStreamWaitEvent(cs, ovl.block_ready[blk], 0));
      CU(ctx, cudaMemcpyAsync(leaves_out + blk * block_elems, dle + blk * block_elems,
                              block_elems * 8, cudaMemcpyDeviceToHost, cs));
    }
  }
  CU(ctx, cudaStreamWaitEvent(cs, all_done, 0));
  if (digests_out && ndig)
    CU(ctx, cudaMemcpyAsync(digests_out, ddi, ndig * 32, cudaMemcpyDeviceToHost, cs));
  CU(ctx, cudaMemcpyAsync(roots_out, dca, nroots * 32, cudaMemcpyDeviceToHost, cs));
  if (want_stats) cudaEventRecord(e3, cs);
  return VPBS_OK;
}

int commit_host_finish(vpbs_ctx* ctx, HostRun* run, vpbs_stats* stats) {
  if (run->upload) {
    const cudaError_t ue = run->upload->join();
    if (ue != cudaSuccess) {
      cudaStreamSynchronize(ctx->stream);
      return fail(ctx, VPBS_ERR_CUDA, std::string("staged upload: ") + cudaGetErrorString(ue));
    }
  }
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (run->chunked) CU(ctx, cudaStreamSynchronize(ctx->h2d_stream));
  if (stats) {
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, run->tm, ctx->launches - run->l0);
    cudaEventElapsedTime(&stats->h2d_ms, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&stats->d2h_ms, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&stats->total_ms, ctx->ev[4], ctx->ev[7]);
  }
  return VPBS_OK;
}

int check_commit_args(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                      uint32_t rate_bits, uint32_t cap_height, const uint64_t* cap_out) {
  if (!cols || ncols == 0 || !cap_out) return fail(ctx, VPBS_ERR_ARG, "null pointer or ncols == 0");
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  if (cap_height > log_n + rate_bits)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  return VPBS_OK;
}

}  // namespace

extern "C" {

int vpbs_commit(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                const uint64_t* const* salt_cols, uint64_t* const* coeffs_out, uint64_t* leaves_out,
                uint64_t* digests_out, uint64_t* cap_out, vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  if ((rc = check_commit_args(ctx, cols, ncols, log_n, rate_bits, cap_height, cap_out))) return rc;
  HostRun run;
  rc = commit_host_enqueue(ctx, cols, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, salt_cols,
                           0, 1ULL << (log_n + rate_bits), coeffs_out, 0, ncols, leaves_out,
                           digests_out, cap_out, stats != nullptr, &run);
  if (rc) return rc;
  return commit_host_finish(ctx, &run, stats);
}

int vpbs_commit_multi(vpbs_ctx* const* ctxs, int nctx, const uint64_t* const* cols, uint32_t ncols,
                      uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                      const uint64_t* const* salt_cols, uint64_t* const* coeffs_out,
                      uint64_t* leaves_out, uint64_t* digests_out, uint64_t* cap_out,
                      vpbs_stats* stats) {
  if (!ctxs || nctx < 1 || !ctxs[0]) return fail(nullptr, VPBS_ERR_ARG, "no contexts");
  vpbs_ctx* c0 = ctxs[0];
  int rc;
  if ((rc = check_commit_args(c0, cols, ncols, log_n, rate_bits, cap_height, cap_out))) return rc;
  if (nctx & (nctx - 1)) return fail(c0, VPBS_ERR_ARG, "nctx must be a power of two");
  if ((u64)nctx > (1ULL << rate_bits) || (u64)nctx > (1ULL << cap_height))
    return fail(c0, VPBS_ERR_ARG,
                "nctx must not exceed 2^rate_bits (whole LDE blocks per GPU) or 2^cap_height "
                "(whole cap subtrees per GPU)");
  for (int g = 0; g < nctx; g++) {
    if (!ctxs[g]) return fail(c0, VPBS_ERR_ARG, "ctxs[g] == NULL");
    for (int k = 0; k < g; k++)
      if (ctxs[k] == ctxs[g]) return fail(c0, VPBS_ERR_ARG, "the same context twice");
  }
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  const u32 width = ncols + (salt_cols ? VPBS_SALT_SIZE : 0);
  const u64 ncap = 1ULL << cap_height;
  const u64 shard = m / (u64)nctx, cap_per = ncap / (u64)nctx, dig_per = 2 * (shard - cap_per);
  std::vector<HostRun> runs((size_t)nctx);
  std::vector<vpbs_stats> st((size_t)nctx);
  // Inputs: every column chunk crosses PCIe ONCE, into the GPU that owns it (chunk k -> GPU k mod
  // nctx, all host links in parallel), and reaches the other GPUs by peer copies over NVLink.
  std::vector<cudaEvent_t> delivered;
  if (nctx > 1) {
    for (u32 c = 0; c < ncols; c++)
      if (!cols[c]) return fail(c0, VPBS_ERR_ARG, "cols[c] == NULL");
    std::vector<u64*> din((size_t)nctx, nullptr);
    for (int g = 0; g < nctx; g++) {  // buffers, peer access, and where each compute stream stands now
      vpbs_ctx* ctx = ctxs[g];
      if ((rc = bind(ctx))) return rc;
      if ((rc = arena_get(ctx, "in", (size_t)ncols * n * 8, (void**)&din[g]))) return rc;
      if (!ctx->peer_tail) CU(ctx, cudaEventCreateWithFlags(&ctx->peer_tail, cudaEventDisableTiming));
      CU(ctx, cudaEventRecord(ctx->peer_tail, ctx->stream));
      for (int o = 0; o < nctx; o++) {
        if (o == g || ctx->peers_enabled.count(ctxs[o]->device)) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[o]->device) == cudaSuccess && can) {
          cudaError_t pe = cudaDeviceEnablePeerAccess(ctxs[o]->device, 0);
          if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
            return fail(ctx, VPBS_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
        }
        cudaGetLastError();
        ctx->peers_enabled.insert(ctxs[o]->device);
      }
    }
    const u32 chunk_cols = host_chunk_cols(ncols, log_n);
    const u32 step = chunk_cols ? chunk_cols : ncols;
    const u32 nchunks = (ncols + step - 1) / step;
    delivered.resize(nchunks);
    std::vector<bool> started((size_t)nctx, false);
    for (u32 k = 0, col0 = 0; k < nchunks; k++, col0 += step) {
      const u32 col1 = col0 + step < ncols ? col0 + step : ncols;
      const int owner = (int)(k % (u32)nctx);
      vpbs_ctx* oc = ctxs[owner];
      if ((rc = bind(oc))) return rc;
      while (oc->peer_ev.size() < nchunks) {
        cudaEvent_t e;
        CU(oc, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        oc->peer_ev.push_back(e);
      }
      if (!started[owner]) {  // nobody's `in` buffer is overwritten while its previous commit reads it
        for (int g = 0; g < nctx; g++) CU(oc, cudaStreamWaitEvent(oc->h2d_stream, ctxs[g]->peer_tail, 0));
        started[owner] = true;
      }
      CU(oc, copy_columns(din[owner], cols, col0, col1, n, true, oc->h2d_stream));
      const size_t off = (size_t)col0 * n, bytes = (size_t)(col1 - col0) * n * sizeof(u64);
      for (int g = 0; g < nctx; g++)
        if (g != owner)
          CU(oc, cudaMemcpyPeerAsync(din[g] + off, ctxs[g]->device, din[owner] + off, oc->device, bytes,
                                     oc->h2d_stream));
      CU(oc, cudaEventRecord(oc->peer_ev[k], oc->h2d_stream));
      delivered[k] = oc->peer_ev[k];
    }
  }
  int enq = 0;
  for (; enq < nctx; enq++) {  // start every device ...
    vpbs_ctx* ctx = ctxs[enq];
    if ((rc = bind(ctx))) break;
    const u32 cc0 = (u32)((u64)ncols * enq / nctx), cc1 = (u32)((u64)ncols * (enq + 1) / nctx);
    rc = commit_host_enqueue(ctx, cols, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs,
                             salt_cols, shard * enq, shard, coeffs_out, cc0, cc1,
                             leaves_out ? leaves_out + shard * enq * width : nullptr,
                             digests_out ? digests_out + dig_per * enq * 4 : nullptr,
                             cap_out + cap_per * enq * 4, stats != nullptr, &runs[enq],
                             nctx > 1 ? &delivered : nullptr);
    if (rc) break;
  }
  int first_err = rc;
  vpbs_ctx* err_ctx = rc ? ctxs[enq] : nullptr;
  for (int g = 0; g < enq; g++) {  // ... then wait for all of them
    vpbs_ctx* ctx = ctxs[g];
    int r2 = bind(ctx);
    if (!r2 && nctx > 1 && cudaStreamSynchronize(ctx->h2d_stream) != cudaSuccess)
      r2 = fail(ctx, VPBS_ERR_CUDA, "input delivery failed");
    if (!r2) r2 = commit_host_finish(ctx, &runs[g], stats ? &st[g] : nullptr);
    if (r2 && !first_err) {
      first_err = r2;
      err_ctx = ctx;
    }
  }
  if (first_err) {
    if (err_ctx && err_ctx != c0) c0->err = err_ctx->err;  // vpbs_last_error(ctxs[0]) explains
    return first_err;
  }
  if (stats) {  // the slowest device bounds the call; launches add up
    *stats = st[0];
    for (int g = 1; g < nctx; g++) {
      stats->kernel_launches += st[g].kernel_launches;
      if (st[g].total_ms > stats->total_ms) {
        const uint64_t l = stats->kernel_launches;
        *stats = st[g];
        stats->kernel_launches = l;
      }
    }
  }
  return VPBS_OK;
}

// ---- openings ------------------------------------------------------------------------------------------
static int eval_ext2_device(vpbs_ctx* ctx, const u64* d_coeffs, u32 ncols, u32 log_n,
                            const uint64_t* points, u32 npoints, uint64_t* out) {
  int rc;
  u64 *d_pts = nullptr, *d_out = nullptr;
  if ((rc = arena_get(ctx, "idx", (size_t)npoints * 16, (void**)&d_pts))) return rc;
  if ((rc = arena_get(ctx, "rows", (size_t)npoints * ncols * 16, (void**)&d_out))) return rc;
  CU(ctx, cudaMemcpyAsync(d_pts, points, (size_t)npoints * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (log_n >= 14)
    ntt::eval_ext2<1024><<<dim3(ncols, npoints), 1024, 0, ctx->stream>>>(d_coeffs, 1ULL << log_n, log_n, d_pts,
                                                               d_out, ncols);
  else
    ntt::eval_ext2<256><<<dim3(ncols, npoints), 256, 0, ctx->stream>>>(d_coeffs, 1ULL << log_n, log_n, d_pts,
                                                               d_out, ncols);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(out, d_out, (size_t)npoints * ncols * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_eval_ext2(vpbs_ctx* ctx, const uint64_t* const* coeff_cols, uint32_t ncols, uint32_t log_n,
                   const uint64_t* points, uint32_t npoints, uint64_t* out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (npoints == 0 || ncols == 0) return VPBS_OK;
  if (!coeff_cols || !points || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  if (log_n > 30 || npoints > 65535) return fail(ctx, VPBS_ERR_ARG, "log_n > 30 or too many points");
  const u64 n = 1ULL << log_n;
  u64* din = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)ncols * n * 8, (void**)&din))) return rc;
  for (u32 c = 0; c < ncols; c++) {
    if (!coeff_cols[c]) return fail(ctx, VPBS_ERR_ARG, "coeff_cols[c] == NULL");
    CU(ctx, cudaMemcpyAsync(din + (u64)c * n, coeff_cols[c], n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  return eval_ext2_device(ctx, din, ncols, log_n, points, npoints, out);
}

int vpbs_batch_eval_ext2(vpbs_batch* b, const uint64_t* points, uint32_t npoints, uint64_t* out) {
  if (!b) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = b->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (npoints == 0) return VPBS_OK;
  if (!points || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  if (npoints > 65535) return fail(ctx, VPBS_ERR_ARG, "too many points");
  return eval_ext2_device(ctx, b->coeffs, b->ncols, b->log_n, points, npoints, out);
}

// [P2] plonk/proof.rs OpeningSet::new evaluates the polynomials of all four oracles at the same points:
// one upload of the points, one kernel per batch, one synchronisation (four separate
// vpbs_batch_eval_ext2 calls cost the N=1024 step 0.69 ms, almost all of it round trips).
int vpbs_batches_eval_ext2(vpbs_batch* const* batches, uint32_t nbatches, const uint64_t* points,
                           uint32_t npoints, uint64_t* const* outs) {
  if (!batches || nbatches == 0 || !batches[0]) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = batches[0]->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (npoints == 0) return VPBS_OK;
  if (!points || !outs) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  if (npoints > 65535) return fail(ctx, VPBS_ERR_ARG, "too many points");
  size_t total = 0;
  for (u32 k = 0; k < nbatches; k++) {
    if (!batches[k] || batches[k]->ctx != ctx) return fail(ctx, VPBS_ERR_STATE, "batches of different contexts");
    if (!outs[k]) return fail(ctx, VPBS_ERR_ARG, "outs[k] == NULL");
    total += (size_t)npoints * batches[k]->ncols * 16;
  }
  u64 *d_pts = nullptr, *d_out = nullptr;
  if ((rc = arena_get(ctx, "idx", (size_t)npoints * 16, (void**)&d_pts))) return rc;
  if ((rc = arena_get(ctx, "rows", total, (void**)&d_out))) return rc;
  CU(ctx, cudaMemcpyAsync(d_pts, points, (size_t)npoints * 16, cudaMemcpyHostToDevice, ctx->stream));
  size_t off = 0;
  for (u32 k = 0; k < nbatches; k++) {
    const vpbs_batch* b = batches[k];
    u64* o = d_out + off / 8;
    if (b->log_n >= 14)
      ntt::eval_ext2<1024><<<dim3(b->ncols, npoints), 1024, 0, ctx->stream>>>(b->coeffs, 1ULL << b->log_n,
                                                                            b->log_n, d_pts, o, b->ncols);
    else
      ntt::eval_ext2<256><<<dim3(b->ncols, npoints), 256, 0, ctx->stream>>>(b->coeffs, 1ULL << b->log_n,
                                                                          b->log_n, d_pts, o, b->ncols);
    ctx->launches++;
    const size_t bytes = (size_t)npoints * b->ncols * 16;
    CU(ctx, cudaMemcpyAsync(outs[k], o, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    off += bytes;
  }
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// ---- FRI commit phase --------------------------------------------------------------------------------
int vpbs_fri_layer_commit(vpbs_ctx* ctx, const uint64_t* values_ext, uint64_t len,
                          uint32_t arity_bits, uint32_t cap_height, uint64_t* leaves_out,
                          uint64_t* digests_out, uint64_t* cap_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  const int lg = log2_strict(len);
  if (lg < 0 || lg > 30) return fail(ctx, VPBS_ERR_ARG, "values.len() must be a power of two (<= 2^30)");
  if ((int)arity_bits > lg) return fail(ctx, VPBS_ERR_ARG, "arity larger than the vector");
  const u64 nleaves = len >> arity_bits;
  const unsigned log_leaves = (unsigned)lg - arity_bits;
  if (cap_height > log_leaves)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  if (!values_ext || !cap_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const u64 ncap = 1ULL << cap_height, ndig = 2 * (nleaves - ncap);
  if (ndig && !digests_out) return fail(ctx, VPBS_ERR_ARG, "digests_out == NULL");
  u64 *dv = nullptr, *dl = nullptr, *dd = nullptr, *dc = nullptr;
  if ((rc = arena_get(ctx, "in", len * 16, (void**)&dv))) return rc;
  if ((rc = arena_get(ctx, "leaves", len * 16, (void**)&dl))) return rc;
  if ((rc = arena_get(ctx, "digests", ndig * 32, (void**)&dd))) return rc;
  if ((rc = arena_get(ctx, "cap", ncap * 32, (void**)&dc))) return rc;
  CU(ctx, cudaMemcpyAsync(dv, values_ext, len * 16, cudaMemcpyHostToDevice, ctx->stream));
  ntt::fri_gather_leaves<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(
      (const ulonglong2*)dv, (unsigned)lg, (ulonglong2*)dl);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  if ((rc = merkle_build(ctx, dl, nleaves, 2u << arity_bits, log_leaves - cap_height, dd, dc))) return rc;
  if (leaves_out) CU(ctx, cudaMemcpyAsync(leaves_out, dl, len * 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (ndig) CU(ctx, cudaMemcpyAsync(digests_out, dd, ndig * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(cap_out, dc, ncap * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_fri_fold(vpbs_ctx* ctx, const uint64_t* coeffs_ext, uint64_t len, uint32_t arity_bits,
                  const uint64_t beta[2], uint64_t shift_next, uint64_t* coeffs_out,
                  uint64_t* values_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  const int lg = log2_strict(len);
  if (lg < 0 || lg > 30) return fail(ctx, VPBS_ERR_ARG, "coeffs.len() must be a power of two (<= 2^30)");
  if ((int)arity_bits > lg) return fail(ctx, VPBS_ERR_ARG, "arity larger than the vector");
  if (!coeffs_ext || !beta || !coeffs_out || !values_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const u64 out_len = len >> arity_bits;
  const unsigned log_out = (unsigned)lg - arity_bits;
  u64 *dc = nullptr, *df = nullptr, *dp = nullptr, *dw = nullptr, *dq = nullptr, *dv = nullptr;
  if ((rc = arena_get(ctx, "in", len * 16, (void**)&dc))) return rc;
  if ((rc = arena_get(ctx, "coeffs", out_len * 16, (void**)&df))) return rc;
  if ((rc = arena_get(ctx, "pad", out_len * 16, (void**)&dp))) return rc;
  if ((rc = arena_get(ctx, "work", out_len * 16, (void**)&dw))) return rc;
  if ((rc = arena_get(ctx, "leaves", out_len * 16, (void**)&dq))) return rc;
  if ((rc = arena_get(ctx, "rows", out_len * 16, (void**)&dv))) return rc;
  if ((rc = ensure_roots(ctx, log_out))) return rc;
  const u64* scale = nullptr;
  if ((rc = get_coset_table(ctx, log_out, 0, shift_next, &scale))) return rc;
  CU(ctx, cudaMemcpyAsync(dc, coeffs_ext, len * 16, cudaMemcpyHostToDevice, ctx->stream));
  ntt::fri_fold<<<(unsigned)((out_len + 127) / 128), 128, 0, ctx->stream>>>(
      (const ulonglong2*)dc, out_len, arity_bits, gl::canon(beta[0]), gl::canon(beta[1]),
      (ulonglong2*)df, dp);
  ctx->launches++;
  if ((rc = run_transform<false>(ctx, dp, out_len, 2, log_out, dw, Out::Natural, dq, out_len, 0, scale, 1)))
    return rc;
  ntt::interleave2<<<(unsigned)((out_len + 255) / 256), 256, 0, ctx->stream>>>(dq, out_len, (ulonglong2*)dv);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(coeffs_out, df, out_len * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(values_out, dv, out_len * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}


// ---- FRI commit phase as one device-resident chain ------------------------------------------------------
void vpbs_fri_destroy(vpbs_fri* f) {
  if (!f) return;
  if (f->ctx) {
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    f->ctx->fri_chains.erase(f);
    // back to the context's pool: a prover runs one commit phase of the same shape per proof, and
    // cudaMalloc / cudaFree of these buffers cost several times the phase's kernels (6.98 -> 1.43 ms)
    pool_free(f->ctx, f->coeffs, f->cap_len * 16);
    pool_free(f->ctx, f->values, f->cap_len * 16);
    pool_free(f->ctx, f->scratch, f->cap_len * 16 * 3);
    for (auto& l : f->layers) {
      pool_free(f->ctx, l.leaves, l.leaves_bytes);
      pool_free(f->ctx, l.digests, l.digests_bytes);
      pool_free(f->ctx, l.cap, l.cap_bytes);
    }
  }
  delete f;
}

}  // extern "C"

namespace {
// values = coeffs.coset_fft(shift) for `len` extension elements held interleaved in f->coeffs
// (natural order in and out), through the planar two-column transform.
int fri_evaluate(vpbs_fri* f) {
  vpbs_ctx* ctx = f->ctx;
  const int lg = log2_strict(f->len);
  int rc;
  if ((rc = ensure_roots(ctx, (unsigned)lg))) return rc;
  const u64* scale = nullptr;
  if ((rc = get_coset_table(ctx, (unsigned)lg, 0, f->shift, &scale))) return rc;
  u64* planar = f->scratch;                 // 2 * len
  u64* work = f->scratch + 2 * f->cap_len;  // 2 * len
  u64* outp = f->scratch + 4 * f->cap_len;  // 2 * len
  ntt::deinterleave2_pad<<<(unsigned)((f->len + 255) / 256), 256, 0, ctx->stream>>>(
      (const ulonglong2*)f->coeffs, f->len, f->len, planar);
  ctx->launches++;
  if ((rc = run_transform<false>(ctx, planar, f->len, 2, (unsigned)lg, work, Out::Natural, outp, f->len, 0,
                                 scale, 1)))
    return rc;
  ntt::interleave2<<<(unsigned)((f->len + 255) / 256), 256, 0, ctx->stream>>>(outp, f->len,
                                                                            (ulonglong2*)f->values);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}
}  // namespace

extern "C" {

int vpbs_fri_begin(vpbs_ctx* ctx, const uint64_t* final_poly_coeffs_ext, uint64_t ncoeffs,
                   uint32_t rate_bits, vpbs_fri** out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!final_poly_coeffs_ext || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  *out = nullptr;
  const int lg = log2_strict(ncoeffs);
  if (lg < 0 || lg + (int)rate_bits > 30)
    return fail(ctx, VPBS_ERR_ARG, "coeffs.len() must be a power of two with len << rate_bits <= 2^30");
  vpbs_fri* f = new (std::nothrow) vpbs_fri();
  if (!f) return fail(ctx, VPBS_ERR_OOM, "host allocation failed");
  f->ctx = ctx;
  f->len = f->cap_len = ncoeffs << rate_bits;
  f->shift = gl::COSET_SHIFT;
  cudaError_t e = pool_alloc(ctx, f->cap_len * 16, &f->coeffs);
  if (e == cudaSuccess) e = pool_alloc(ctx, f->cap_len * 16, &f->values);
  if (e == cudaSuccess) e = pool_alloc(ctx, f->cap_len * 16 * 3, &f->scratch);
  // [P2] PolynomialCoeffs::lde: the coefficient vector zero-padded to len << rate_bits
  if (e == cudaSuccess) e = cudaMemsetAsync(f->coeffs, 0, f->cap_len * 16, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(f->coeffs, final_poly_coeffs_ext, ncoeffs * 16, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    pool_free(ctx, f->coeffs, f->cap_len * 16);
    pool_free(ctx, f->values, f->cap_len * 16);
    pool_free(ctx, f->scratch, f->cap_len * 16 * 3);
    delete f;
    return fail(ctx, e == cudaErrorMemoryAllocation ? VPBS_ERR_OOM : VPBS_ERR_CUDA,
                std::string("fri begin: ") + cudaGetErrorString(e));
  }
  ctx->fri_chains.insert(f);
  if ((rc = fri_evaluate(f))) {  // lde_final_values = lde_final_poly.coset_fft(F::coset_shift())
    vpbs_fri_destroy(f);
    return rc;
  }
  *out = f;
  return VPBS_OK;
}

// [P2] fri/oracle.rs prove_openings up to lde_final_values, from resident batches (openings.cuh).
int vpbs_fri_begin_openings(vpbs_ctx* ctx, vpbs_batch* const* oracles, uint32_t noracles,
                            const uint32_t* batch_sizes, uint32_t nbatches, const uint32_t* poly_refs,
                            const uint64_t* points, const uint64_t alpha[2], uint32_t rate_bits,
                            vpbs_fri** out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!oracles || !batch_sizes || !poly_refs || !points || !alpha || !out || noracles == 0 || nbatches == 0)
    return fail(ctx, VPBS_ERR_ARG, "null pointer, no oracles or no batches");
  *out = nullptr;
  for (u32 o = 0; o < noracles; o++) {
    if (!oracles[o] || oracles[o]->ctx != ctx)
      return fail(ctx, VPBS_ERR_STATE, "oracle batch is NULL or belongs to another context");
    if (oracles[o]->log_n != oracles[0]->log_n)
      return fail(ctx, VPBS_ERR_ARG, "all oracles must have the same degree");
  }
  const u32 log_n = oracles[0]->log_n;
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  const u64 n = 1ULL << log_n;
  u64 total = 0;
  u32 longest = 0;
  for (u32 b = 0; b < nbatches; b++) {
    if (batch_sizes[b] == 0) return fail(ctx, VPBS_ERR_ARG, "empty FRI batch");
    total += batch_sizes[b];
    if (batch_sizes[b] > longest) longest = batch_sizes[b];
  }
  std::vector<const u64*> ptrs(total);
  for (u64 k = 0; k < total; k++) {
    const u32 o = poly_refs[2 * k], pi = poly_refs[2 * k + 1];
    if (o >= noracles || pi >= oracles[o]->ncols)
      return fail(ctx, VPBS_ERR_ARG, "FriPolynomialInfo out of range");
    ptrs[k] = oracles[o]->coeffs + (u64)pi * n;
  }
  // alpha^j, j <= longest (alpha^len_b is the shift of batch b), on the host
  auto emul = [](const u64 a[2], const u64 b[2], u64 r[2]) {
    const u64 t = gl::mul(a[1], b[1]);
    const u64 re = gl::add(gl::mul(a[0], b[0]), gl::mul(7, t));
    const u64 im = gl::add(gl::mul(a[0], b[1]), gl::mul(a[1], b[0]));
    r[0] = re;
    r[1] = im;
  };
  const u64 al[2] = {gl::canon(alpha[0]), gl::canon(alpha[1])};
  std::vector<u64> pows(2 * ((size_t)longest + 1));
  pows[0] = 1;
  pows[1] = 0;
  for (u32 j = 1; j <= longest; j++) emul(&pows[2 * (j - 1)], al, &pows[2 * j]);

  vpbs_fri* f = new (std::nothrow) vpbs_fri();
  if (!f) return fail(ctx, VPBS_ERR_OOM, "host allocation failed");
  f->ctx = ctx;
  f->len = f->cap_len = n << rate_bits;
  f->shift = gl::COSET_SHIFT;
  cudaError_t e = pool_alloc(ctx, f->cap_len * 16, &f->coeffs);
  if (e == cudaSuccess) e = pool_alloc(ctx, f->cap_len * 16, &f->values);
  if (e == cudaSuccess) e = pool_alloc(ctx, f->cap_len * 16 * 3, &f->scratch);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->coeffs, 0, f->cap_len * 16, ctx->stream);  // lde(): zero padding
  if (e != cudaSuccess) {
    pool_free(ctx, f->coeffs, f->cap_len * 16);
    pool_free(ctx, f->values, f->cap_len * 16);
    pool_free(ctx, f->scratch, f->cap_len * 16 * 3);
    delete f;
    return fail(ctx, e == cudaErrorMemoryAllocation ? VPBS_ERR_OOM : VPBS_ERR_CUDA,
                std::string("fri begin: ") + cudaGetErrorString(e));
  }
  ctx->fri_chains.insert(f);
  auto bail = [&](int code) {
    vpbs_fri_destroy(f);
    return code;
  };
  const u64 nblocks = (n + openings::DIV_BLOCK - 1) / openings::DIV_BLOCK;
  u64 *d_ptrs = nullptr, *d_pows = nullptr, *d_tot = nullptr;
  if ((rc = arena_get(ctx, "open_ptrs", total * 8, (void**)&d_ptrs))) return bail(rc);
  if ((rc = arena_get(ctx, "open_pows", pows.size() * 8, (void**)&d_pows))) return bail(rc);
  if ((rc = arena_get(ctx, "open_tot", nblocks * 32, (void**)&d_tot))) return bail(rc);
  ulonglong2* comp = (ulonglong2*)f->scratch;  // n extension elements; free until fri_evaluate
  ulonglong2* totals = (ulonglong2*)d_tot;
  ulonglong2* carries = totals + nblocks;
  if (cudaMemcpyAsync(d_ptrs, ptrs.data(), total * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(d_pows, pows.data(), pows.size() * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return bail(fail(ctx, VPBS_ERR_CUDA, "openings: upload failed"));
  u64 off = 0;
  for (u32 b = 0; b < nbatches; b++) {
    const u32 len = batch_sizes[b];
    const u64 z0 = gl::canon(points[2 * b]), z1 = gl::canon(points[2 * b + 1]);
    openings::reduce_polys<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        (const u64* const*)d_ptrs + off, len, (const ulonglong2*)d_pows, n, comp);
    openings::divide_pass1<<<(unsigned)nblocks, openings::DIV_THREADS, 0, ctx->stream>>>(comp, n, z0, z1, totals);
    openings::divide_carries<<<1, 32, 0, ctx->stream>>>(totals, nblocks, z0, z1, carries);
    openings::divide_pass2<<<(unsigned)nblocks, openings::DIV_THREADS, 0, ctx->stream>>>(
        comp, n, z0, z1, carries, pows[2 * (size_t)len], pows[2 * (size_t)len + 1], b == 0,
        (ulonglong2*)f->coeffs);
    ctx->launches += 4;
    off += len;
  }
  if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, VPBS_ERR_CUDA, "openings: launch failed"));
  if ((rc = fri_evaluate(f))) return bail(rc);  // lde_final_values ("perform final FFT")
  // the pointer / power tables above are host vectors: make sure the copies have left them
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return bail(fail(ctx, VPBS_ERR_CUDA, "openings: synchronize failed"));
  *out = f;
  return VPBS_OK;
}

int vpbs_fri_commit_layer(vpbs_fri* f, uint32_t arity_bits, uint32_t cap_height, uint64_t* cap_out) {
  if (!f) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = f->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (!cap_out) return fail(ctx, VPBS_ERR_ARG, "cap_out == NULL");
  if (f->layer_open) return fail(ctx, VPBS_ERR_STATE, "the previous layer has not been folded yet");
  const int lg = log2_strict(f->len);
  if ((int)arity_bits > lg) return fail(ctx, VPBS_ERR_ARG, "arity larger than the vector");
  const unsigned log_leaves = (unsigned)lg - arity_bits;
  if (cap_height > log_leaves)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  vpbs_fri::Layer L;
  L.nleaves = f->len >> arity_bits;
  L.leaf_len = 2u << arity_bits;
  L.cap_height = cap_height;
  const u64 ncap = 1ULL << cap_height, ndig = 2 * (L.nleaves - ncap);
  L.leaves_bytes = f->len * 16;
  L.digests_bytes = ndig ? ndig * 32 : 32;
  L.cap_bytes = ncap * 32;
  cudaError_t e = pool_alloc(ctx, L.leaves_bytes, &L.leaves);
  if (e == cudaSuccess) e = pool_alloc(ctx, L.digests_bytes, &L.digests);
  if (e == cudaSuccess) e = pool_alloc(ctx, L.cap_bytes, &L.cap);
  if (e != cudaSuccess) {
    pool_free(ctx, L.leaves, L.leaves_bytes);
    pool_free(ctx, L.digests, L.digests_bytes);
    pool_free(ctx, L.cap, L.cap_bytes);
    return fail(ctx, VPBS_ERR_OOM, std::string("fri layer allocation: ") + cudaGetErrorString(e));
  }
  f->layers.push_back(L);
  // reverse_index_bits_in_place(values) + chunking = a pure re-indexing (ntt::fri_gather_leaves)
  ntt::fri_gather_leaves<<<(unsigned)((f->len + 255) / 256), 256, 0, ctx->stream>>>(
      (const ulonglong2*)f->values, (unsigned)lg, (ulonglong2*)L.leaves);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  if ((rc = merkle_build(ctx, L.leaves, L.nleaves, L.leaf_len, log_leaves - cap_height, L.digests, L.cap)))
    return rc;
  CU(ctx, cudaMemcpyAsync(cap_out, L.cap, ncap * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));  // the challenger needs the cap before beta exists
  f->pending_arity_bits = arity_bits;
  f->layer_open = true;
  return VPBS_OK;
}

int vpbs_fri_fold_layer(vpbs_fri* f, const uint64_t beta[2]) {
  if (!f) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = f->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (!beta) return fail(ctx, VPBS_ERR_ARG, "beta == NULL");
  if (!f->layer_open) return fail(ctx, VPBS_ERR_STATE, "no committed layer to fold");
  const u32 ab = f->pending_arity_bits;
  const u64 out_len = f->len >> ab;
  // coeffs' = chunks_exact(arity).map(|c| reduce_with_powers(c, beta)); written to the value buffer
  // (free until the next evaluation) and swapped in
  ntt::fri_fold<<<(unsigned)((out_len + 127) / 128), 128, 0, ctx->stream>>>(
      (const ulonglong2*)f->coeffs, out_len, ab, gl::canon(beta[0]), gl::canon(beta[1]),
      (ulonglong2*)f->values, f->scratch /* planar copy, unused here */);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  std::swap(f->coeffs, f->values);
  f->len = out_len;
  f->shift = gl::pow(f->shift, 1ULL << ab);  // shift = shift.exp_u64(arity)
  f->layer_open = false;
  return fri_evaluate(f);                    // values = coeffs.coset_fft(shift)
}

int vpbs_fri_final_poly(vpbs_fri* f, uint32_t rate_bits, uint64_t* coeffs_out) {
  if (!f) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = f->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (!coeffs_out) return fail(ctx, VPBS_ERR_ARG, "coeffs_out == NULL");
  if (f->layer_open) return fail(ctx, VPBS_ERR_STATE, "the last layer has not been folded yet");
  const u64 keep = f->len >> rate_bits;  // coeffs.truncate(len >> rate_bits)
  if (keep == 0) return fail(ctx, VPBS_ERR_ARG, "rate_bits exceeds the polynomial's length");
  CU(ctx, cudaMemcpyAsync(coeffs_out, f->coeffs, keep * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_fri_query_layer(vpbs_fri* f, uint32_t layer, const uint64_t* leaf_indices, uint64_t count,
                         uint64_t* rows_out, uint64_t* siblings_out) {
  if (!f) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = f->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (layer >= f->layers.size()) return fail(ctx, VPBS_ERR_ARG, "no such layer");
  if (count == 0) return VPBS_OK;
  if (!leaf_indices || !rows_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const vpbs_fri::Layer& L = f->layers[layer];
  for (uint64_t i = 0; i < count; i++)
    if (leaf_indices[i] >= L.nleaves) return fail(ctx, VPBS_ERR_ARG, "leaf index out of range");
  const int lg = log2_strict(L.nleaves);
  const unsigned num_layers = (unsigned)lg - L.cap_height;
  u64 *d_idx = nullptr, *d_rows = nullptr, *d_sib = nullptr;
  if ((rc = arena_get(ctx, "idx", count * 8, (void**)&d_idx))) return rc;
  if ((rc = arena_get(ctx, "rows", count * (size_t)L.leaf_len * 8, (void**)&d_rows))) return rc;
  CU(ctx, cudaMemcpyAsync(d_idx, leaf_indices, count * 8, cudaMemcpyHostToDevice, ctx->stream));
  merkle::gather_rows<<<(unsigned)count, 128, 0, ctx->stream>>>(L.leaves, L.leaf_len, d_idx, count, d_rows);
  ctx->launches++;
  CU(ctx, cudaMemcpyAsync(rows_out, d_rows, count * (size_t)L.leaf_len * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (num_layers && siblings_out) {
    if ((rc = arena_get(ctx, "sibs", count * (size_t)num_layers * 32, (void**)&d_sib))) return rc;
    const u64 total = count * num_layers;
    merkle::gather_siblings<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(
        L.digests, d_idx, count, num_layers, 2 * (1ULL << num_layers) - 2, d_sib);
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(siblings_out, d_sib, count * (size_t)num_layers * 32, cudaMemcpyDeviceToHost,
                            ctx->stream));
  }
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// ---- FRI proof of work -----------------------------------------------------------------------------
int vpbs_pow_grind(vpbs_ctx* ctx, const uint64_t state[12], uint32_t witness_pos,
                   uint32_t response_lane, uint32_t min_leading_zeros, uint64_t first_candidate,
                   uint64_t count, uint64_t* witness_out, int* found) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!state || !witness_out || !found) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  if (witness_pos >= 12 || response_lane >= 12 || min_leading_zeros > 64)
    return fail(ctx, VPBS_ERR_ARG, "witness_pos / response_lane / min_leading_zeros out of range");
  *found = 0;
  u64* d = nullptr;
  if ((rc = arena_get(ctx, "pow", 13 * 8, (void**)&d))) return rc;
  const unsigned long long none = ~0ULL;
  CU(ctx, cudaMemcpyAsync(d, state, 12 * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d + 12, &none, 8, cudaMemcpyHostToDevice, ctx->stream));
  // candidates per launch; stop at the first chunk with a hit.  A witness is expected after
  // 2^min_leading_zeros candidates: four times that per launch finds it in the first one with
  // probability 1 - e^-4 without hashing millions of candidates beyond it (a 2^22 chunk cost the
  // 16-bit grind of the N=1024 step 0.52 ms; 2^18: see DESIGN.md §9.2)
  u64 chunk = min_leading_zeros < 20 ? (4ULL << min_leading_zeros) : (1ULL << 22);
  if (chunk < (1ULL << 17)) chunk = 1ULL << 17;
  for (u64 off = 0; off < count; off += chunk) {
    const u64 cnt = count - off < chunk ? count - off : chunk;
    merkle::pow_grind<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(
        d, witness_pos, response_lane, min_leading_zeros, first_candidate + off, cnt,
        (unsigned long long*)(d + 12));
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    unsigned long long best = none;
    CU(ctx, cudaMemcpyAsync(&best, d + 12, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (best != none) {
      *witness_out = best;
      *found = 1;
      break;
    }
  }
  return VPBS_OK;
}

// ---- device-resident batches ---------------------------------------------------------------------------
void vpbs_batch_destroy(vpbs_batch* b) {
  if (!b) return;
  if (!b->ctx) {  // the context went first and already released the device buffers
    delete b;
    return;
  }
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  b->ctx->batches.erase(b);
  pool_free(b->ctx, b->coeffs, b->coeffs_bytes);
  pool_free(b->ctx, b->leaves, b->leaves_bytes);
  pool_free(b->ctx, b->digests, b->digests_bytes);
  pool_free(b->ctx, b->cap, b->cap_bytes);
  delete b;
}

int vpbs_batch_commit(vpbs_ctx* ctx, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n,
                      uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                      const uint64_t* const* salt_cols, uint64_t* cap_out, vpbs_batch** out,
                      vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!cols || ncols == 0 || !cap_out || !out)
    return fail(ctx, VPBS_ERR_ARG, "null pointer or ncols == 0");
  *out = nullptr;
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  const unsigned log_m = log_n + rate_bits;
  if (cap_height > log_m)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  const u64 ncap = 1ULL << cap_height;
  vpbs_batch* b = nullptr;
  if ((rc = batch_alloc(ctx, ncols, log_n, rate_bits, cap_height, salt_cols != nullptr,
                        inputs_are_coeffs != 0, &b)))
    return rc;
  u64 *din = nullptr, *dsa = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)ncols * n * 8, (void**)&din)) ||
      (salt_cols && (rc = arena_get(ctx, "salt", (size_t)4 * m * 8, (void**)&dsa)))) {
    vpbs_batch_destroy(b);
    return rc;
  }
  const uint64_t l0 = ctx->launches;
  cudaEvent_t e0 = ctx->ev[4], e1 = ctx->ev[5], e2 = ctx->ev[6], e3 = ctx->ev[7];
  if (stats) cudaEventRecord(e0, ctx->stream);
  cudaError_t ce = cudaSuccess;
  // Wide batches: inputs travel in column chunks on the H2D stream and chunk k is transformed
  // (IFFT + its columns of every LDE block) while chunk k+1 is still on the bus.
  const u32 chunk_cols = host_chunk_cols(ncols, log_n);
  const u32 nchunks = chunk_cols ? (ncols + chunk_cols - 1) / chunk_cols : 0;
  for (u32 c = 0; c < ncols; c++)
    if (!cols[c]) {
      vpbs_batch_destroy(b);
      return fail(ctx, VPBS_ERR_ARG, "cols[c] == NULL");
    }
  // Pageable host columns go through the context's pinned ring (host_stage.h), also when the batch
  // is too narrow for chunked compute (one upload event then).
  const bool staged = want_staging(ctx, cols, ncols, n);
  const u32 nup = nchunks ? nchunks : (staged ? 1u : 0u);  // upload events
  Overlap ovl;
  std::unique_ptr<hoststage::Upload> upload;
  cudaStream_t hs = nup ? ctx->h2d_stream : ctx->stream;
  if (nup) {
    while (ctx->ov.size() < (size_t)nup + 1 && ce == cudaSuccess) {
      cudaEvent_t e;
      ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (ce == cudaSuccess) ctx->ov.push_back(e);
    }
    if (ce == cudaSuccess) {
      ovl.chunk_cols = chunk_cols;
      ovl.h2d_ready.assign(ctx->ov.begin() + 1, ctx->ov.begin() + 1 + nup);
      // the H2D stream starts after whatever the caller queued on the compute stream
      ce = cudaEventRecord(ctx->ov[0], ctx->stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(hs, ctx->ov[0], 0);
    }
  }
  if (staged && ce == cudaSuccess) {
    if ((rc = start_staged_upload(ctx, din, cols, ncols, n, chunk_cols, ovl.h2d_ready, hs,
                                  stats ? e1 : nullptr, &upload))) {
      vpbs_batch_destroy(b);
      return rc;
    }
    ovl.upload = upload.get();
    ovl.absorb_chunks = true;  // slow arrivals: hash chunk k while chunk k + 1 is still being staged
  }
  for (u32 c0 = 0, k = 0; c0 < ncols && ce == cudaSuccess && !staged; k++) {
    const u32 c1 = (nchunks && c0 + chunk_cols < ncols) ? c0 + chunk_cols : ncols;
    ce = copy_columns(din, cols, c0, c1, n, true, hs);
    if (ce == cudaSuccess && nchunks) ce = cudaEventRecord(ovl.h2d_ready[k], hs);
    c0 = c1;
  }
  if (salt_cols)
    for (int s = 0; s < VPBS_SALT_SIZE && ce == cudaSuccess; s++) {
      if (!salt_cols[s]) { ce = cudaErrorInvalidValue; break; }
      ce = cudaMemcpyAsync(dsa + (u64)s * m, salt_cols[s], m * 8, cudaMemcpyHostToDevice, ctx->stream);
    }
  if (ce != cudaSuccess) {
    if (upload) upload->join();
    if (nup) cudaStreamSynchronize(hs);
    vpbs_batch_destroy(b);
    return fail(ctx, ce == cudaErrorInvalidValue ? VPBS_ERR_ARG : VPBS_ERR_CUDA,
                std::string("batch input copy: ") + cudaGetErrorString(ce));
  }
  if (stats && !staged) cudaEventRecord(e1, hs);
  Timer tm{ctx, stats != nullptr};
  if (nup) tm.overlap = &ovl;
  // coefficients always end up in the batch (from_coeffs: a device copy of the inputs)
  rc = commit_core(ctx, din, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, dsa, b->first_leaf,
                   b->nleaves, b->coeffs, b->leaves, b->digests, b->own_roots(), &tm);
  const cudaError_t ue = upload ? upload->join() : cudaSuccess;
  if (rc == VPBS_OK && ue != cudaSuccess)
    rc = fail(ctx, VPBS_ERR_CUDA, std::string("staged upload: ") + cudaGetErrorString(ue));
  if (rc) {
    if (nup) cudaStreamSynchronize(hs);
    cudaStreamSynchronize(ctx->stream);
    vpbs_batch_destroy(b);
    return rc;
  }
  if (stats) cudaEventRecord(e2, ctx->stream);
  ce = cudaMemcpyAsync(cap_out, b->cap, ncap * 32, cudaMemcpyDeviceToHost, ctx->stream);
  if (stats) cudaEventRecord(e3, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_CUDA, std::string("batch commit: ") + cudaGetErrorString(ce));
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    cudaEventElapsedTime(&stats->h2d_ms, e0, e1);
    cudaEventElapsedTime(&stats->d2h_ms, e2, e3);
    cudaEventElapsedTime(&stats->total_ms, e0, e3);
  }
  *out = b;
  return VPBS_OK;
}

}  // extern "C"

namespace {
// A new batch handle with its four device buffers (from the context's pool).
int batch_alloc(vpbs_ctx* ctx, u32 ncols, u32 log_n, u32 rate_bits, u32 cap_height, bool salted,
                bool coeff_inputs, vpbs_batch** out) {
  const u64 n = 1ULL << log_n, m = n << rate_bits;
  const u32 width = ncols + (salted ? VPBS_SALT_SIZE : 0);
  const u64 ncap = 1ULL << cap_height;
  // the shard of this context: whole n-row LDE blocks covering whole cap subtrees
  u64 first = 0, nl = m;
  if (ctx->shard_count > 1) {
    nl = m / ctx->shard_count;
    if (nl < n || nl < (m >> cap_height) || nl * ctx->shard_count != m)
      return fail(ctx, VPBS_ERR_ARG,
                  "shard count too large for this commit (a shard is whole LDE blocks and whole cap subtrees)");
    first = nl * ctx->shard_index;
  }
  const u64 nroots = nl >> (log_n + rate_bits - cap_height), ndig = 2 * (nl - nroots);
  vpbs_batch* b = new (std::nothrow) vpbs_batch();
  if (!b) return fail(ctx, VPBS_ERR_OOM, "host allocation failed");
  b->ctx = ctx;
  b->ncols = ncols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
  b->width = width;
  b->coeff_inputs = coeff_inputs;
  b->first_leaf = first;
  b->nleaves = nl;
  b->coeffs_bytes = (size_t)ncols * n * 8;
  b->leaves_bytes = (size_t)nl * width * 8;
  b->digests_bytes = ndig ? ndig * 32 : 32;
  b->cap_bytes = ncap * 32;
  ctx->batches.insert(b);
  cudaError_t e = pool_alloc(ctx, b->coeffs_bytes, &b->coeffs);
  if (e == cudaSuccess) e = pool_alloc(ctx, b->leaves_bytes, &b->leaves);
  if (e == cudaSuccess) e = pool_alloc(ctx, b->digests_bytes, &b->digests);
  if (e == cudaSuccess) e = pool_alloc(ctx, b->cap_bytes, &b->cap);
  if (e != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_OOM, std::string("batch allocation: ") + cudaGetErrorString(e));
  }
  if (b->sharded() && (e = cudaMemsetAsync(b->cap, 0, b->cap_bytes, ctx->stream)) != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_CUDA, std::string("batch allocation: ") + cudaGetErrorString(e));
  }
  *out = b;
  return VPBS_OK;
}

// Inclusive prefix product of data[0..count) in place (perm::scan_blocks recursion).
int prefix_product(vpbs_ctx* ctx, u64* data, u64 count, u64* scratch) {
  const u64 nblocks = (count + perm::SCAN_THREADS - 1) / perm::SCAN_THREADS;
  perm::scan_blocks<<<(unsigned)nblocks, perm::SCAN_THREADS, 0, ctx->stream>>>(
      data, count, nblocks > 1 ? scratch : nullptr);
  ctx->launches++;
  if (nblocks > 1) {
    int rc = prefix_product(ctx, scratch, nblocks, scratch + nblocks);
    if (rc) return rc;
    perm::scan_apply<<<(unsigned)nblocks, perm::SCAN_THREADS, 0, ctx->stream>>>(data, count, scratch);
    ctx->launches++;
  }
  CU(ctx, cudaGetLastError());
  return VPBS_OK;
}

// Z and partial products of every challenge on device data, written in commit order into d_out
// (num_challenges * K columns of n): Z_0 .. Z_{c-1}, then the K - 1 partial products of each
// challenge.  d_wires / d_sigmas column-major with the given strides.
int zs_core(vpbs_ctx* ctx, const u64* d_wires, u64 wires_stride, const u64* d_sigmas, const u64* d_kis,
            u32 num_routed, u32 log_n, u32 max_degree, const uint64_t* betas, const uint64_t* gammas,
            u32 num_challenges, u64* d_out) {
  const u64 n = 1ULL << log_n;
  const u32 K = (num_routed + max_degree - 1) / max_degree;
  int rc;
  u64 *quot = nullptr, *rowprod = nullptr;
  int* flag = nullptr;
  if ((rc = arena_get(ctx, "zs_quot", (size_t)K * n * 8, (void**)&quot))) return rc;
  // row products + the scan's block totals of every level (n/256 + n/256^2 + ... < n/128)
  if ((rc = arena_get(ctx, "zs_rows", (size_t)(n + n / 128 + 512) * 8, (void**)&rowprod))) return rc;
  if ((rc = arena_get(ctx, "zs_flag", sizeof(int), (void**)&flag))) return rc;
  if ((rc = ensure_roots(ctx, log_n))) return rc;
  const ntt::Roots R{ctx->roots, ctx->roots_log};
  CU(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  const unsigned grid = (unsigned)((n + 127) / 128);
  for (u32 c = 0; c < num_challenges; c++) {
    perm::chunk_quotients<<<grid, 128, 0, ctx->stream>>>(
        d_wires, wires_stride, d_sigmas, n, d_kis, num_routed, max_degree, log_n,
        gl::canon(betas[c]), gl::canon(gammas[c]), R, quot, rowprod, flag);
    ctx->launches++;
    if ((rc = prefix_product(ctx, rowprod, n, rowprod + n))) return rc;
    perm::finish_rows<<<grid, 128, 0, ctx->stream>>>(
        quot, rowprod, log_n, K, d_out + (u64)c * n,
        d_out + ((u64)num_challenges + (u64)c * (K - 1)) * n);
    ctx->launches++;
  }
  CU(ctx, cudaGetLastError());
  int h_flag = 0;
  CU(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flag)
    return fail(ctx, VPBS_ERR_ARG,
                "a permutation denominator is zero (plonky2's batch_multiplicative_inverse panics)");
  return VPBS_OK;
}

int check_zs_args(vpbs_ctx* ctx, u32 num_routed, u32 log_n, u32 max_degree, u32 num_challenges,
                  const void* betas, const void* gammas) {
  if (!betas || !gammas || num_routed == 0 || num_challenges == 0)
    return fail(ctx, VPBS_ERR_ARG, "null pointer, num_routed == 0 or num_challenges == 0");
  if (max_degree < 2) return fail(ctx, VPBS_ERR_ARG, "max_degree must be at least 2");
  if (log_n > 30) return fail(ctx, VPBS_ERR_ARG, "log_n > 30");
  if ((num_routed + max_degree - 1) / max_degree > (u32)perm::MAX_CHUNKS)
    return fail(ctx, VPBS_ERR_ARG, "more than 32 chunks of routed wires per row");
  return VPBS_OK;
}
}  // namespace

extern "C" {

// ---- permutation argument: Z and partial products ------------------------------------------------------
int vpbs_sigmas_upload(vpbs_ctx* ctx, const uint64_t* const* sigma_cols, const uint64_t* k_is,
                       uint32_t num_routed, uint32_t log_n, vpbs_sigmas** out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!sigma_cols || !k_is || !out || num_routed == 0)
    return fail(ctx, VPBS_ERR_ARG, "null pointer or num_routed == 0");
  if (log_n > 30) return fail(ctx, VPBS_ERR_ARG, "log_n > 30");
  *out = nullptr;
  const u64 n = 1ULL << log_n;
  vpbs_sigmas* sg = new (std::nothrow) vpbs_sigmas();
  if (!sg) return fail(ctx, VPBS_ERR_OOM, "host allocation failed");
  sg->ctx = ctx;
  sg->num_routed = num_routed;
  sg->log_n = log_n;
  cudaError_t e = cudaMalloc(&sg->sigmas, (size_t)num_routed * n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&sg->k_is, (size_t)num_routed * 8);
  std::vector<u64> kc(num_routed);
  for (u32 j = 0; j < num_routed; j++) kc[j] = gl::canon(k_is[j]);
  for (u32 j = 0; j < num_routed && e == cudaSuccess; j++) {
    if (!sigma_cols[j]) e = cudaErrorInvalidValue;
    else e = cudaMemcpyAsync(sg->sigmas + (u64)j * n, sigma_cols[j], n * 8, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(sg->k_is, kc.data(), (size_t)num_routed * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(sg->sigmas);
    cudaFree(sg->k_is);
    delete sg;
    return fail(ctx, e == cudaErrorInvalidValue ? VPBS_ERR_ARG : e == cudaErrorMemoryAllocation ? VPBS_ERR_OOM : VPBS_ERR_CUDA,
                std::string("sigmas upload: ") + cudaGetErrorString(e));
  }
  ctx->sigma_sets.insert(sg);
  *out = sg;
  return VPBS_OK;
}

void vpbs_sigmas_destroy(vpbs_sigmas* sg) {
  if (!sg) return;
  if (sg->ctx) {
    cudaSetDevice(sg->ctx->device);
    cudaStreamSynchronize(sg->ctx->stream);
    sg->ctx->sigma_sets.erase(sg);
    cudaFree(sg->sigmas);
    cudaFree(sg->k_is);
  }
  delete sg;
}

int vpbs_zs_partial_products(vpbs_ctx* ctx, const uint64_t* const* wire_cols,
                             const vpbs_sigmas* sigmas, uint32_t max_degree, const uint64_t* betas,
                             const uint64_t* gammas, uint32_t num_challenges,
                             uint64_t* const* cols_out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!sigmas || sigmas->ctx != ctx) return fail(ctx, VPBS_ERR_STATE, "sigmas handle does not belong to this context");
  if (!wire_cols || !cols_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const u32 nr = sigmas->num_routed, log_n = sigmas->log_n;
  if ((rc = check_zs_args(ctx, nr, log_n, max_degree, num_challenges, betas, gammas))) return rc;
  const u64 n = 1ULL << log_n;
  const u32 K = (nr + max_degree - 1) / max_degree, ncols_out = num_challenges * K;
  u64 *dw = nullptr, *dout = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)nr * n * 8, (void**)&dw))) return rc;
  if ((rc = arena_get(ctx, "zs_out", (size_t)ncols_out * n * 8, (void**)&dout))) return rc;
  for (u32 j = 0; j < nr; j++) {
    if (!wire_cols[j]) return fail(ctx, VPBS_ERR_ARG, "wire_cols[j] == NULL");
    CU(ctx, cudaMemcpyAsync(dw + (u64)j * n, wire_cols[j], n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = zs_core(ctx, dw, n, sigmas->sigmas, sigmas->k_is, nr, log_n, max_degree, betas, gammas,
                    num_challenges, dout)))
    return rc;
  for (u32 c = 0; c < ncols_out; c++)
    if (cols_out[c])
      CU(ctx, cudaMemcpyAsync(cols_out[c], dout + (u64)c * n, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_batch_zs_partial_products(vpbs_batch* wires, const vpbs_sigmas* sigmas, uint32_t max_degree,
                                   const uint64_t* betas, const uint64_t* gammas,
                                   uint32_t num_challenges, uint32_t rate_bits, uint32_t cap_height,
                                   uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats) {
  if (!wires) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = wires->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (!sigmas || sigmas->ctx != ctx) return fail(ctx, VPBS_ERR_STATE, "sigmas handle does not belong to this context");
  if (!cap_out || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  *out = nullptr;
  const u32 nr = sigmas->num_routed, log_n = sigmas->log_n;
  if (wires->log_n != log_n || wires->ncols < nr)
    return fail(ctx, VPBS_ERR_ARG, "wires batch does not match the sigmas (degree or routed wires)");
  if ((rc = check_zs_args(ctx, nr, log_n, max_degree, num_challenges, betas, gammas))) return rc;
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  if (cap_height > log_n + rate_bits)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  const u64 n = 1ULL << log_n;
  const u32 K = (nr + max_degree - 1) / max_degree, ncols_out = num_challenges * K;
  const uint64_t l0 = ctx->launches;
  cudaEvent_t e0 = ctx->ev[4], e3 = ctx->ev[7];
  if (stats) cudaEventRecord(e0, ctx->stream);
  // routed wire VALUES over the subgroup = forward transform of the batch's coefficients (exact)
  u64 *dvals = nullptr, *dwork = nullptr, *dout = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)nr * n * 8, (void**)&dvals))) return rc;
  if ((rc = arena_get(ctx, "work", (size_t)nr * n * 8, (void**)&dwork))) return rc;
  if ((rc = arena_get(ctx, "zs_out", (size_t)ncols_out * n * 8, (void**)&dout))) return rc;
  if ((rc = ensure_roots(ctx, log_n + rate_bits))) return rc;
  if ((rc = run_transform<false>(ctx, wires->coeffs, n, nr, log_n, dwork, Out::Natural, dvals, n, 0,
                                 nullptr, 1)))
    return rc;
  if ((rc = zs_core(ctx, dvals, n, sigmas->sigmas, sigmas->k_is, nr, log_n, max_degree, betas, gammas,
                    num_challenges, dout)))
    return rc;
  vpbs_batch* b = nullptr;
  if ((rc = batch_alloc(ctx, ncols_out, log_n, rate_bits, cap_height, false, false, &b))) return rc;
  Timer tm{ctx, stats != nullptr};
  rc = commit_core(ctx, dout, ncols_out, log_n, rate_bits, cap_height, 0, nullptr, b->first_leaf,
                   b->nleaves, b->coeffs, b->leaves, b->digests, b->own_roots(), &tm);
  if (rc) {
    vpbs_batch_destroy(b);
    return rc;
  }
  cudaError_t ce = cudaMemcpyAsync(cap_out, b->cap, b->cap_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (stats) cudaEventRecord(e3, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_CUDA, std::string("zs batch commit: ") + cudaGetErrorString(ce));
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    cudaEventElapsedTime(&stats->total_ms, e0, e3);
  }
  *out = b;
  return VPBS_OK;
}

// The gate constraints of a circuit as a program (perm::gate_program_eval): validated here, once.
int vpbs_gate_program_upload(vpbs_ctx* ctx, const uint64_t* code, uint32_t ncode, const uint64_t* imms,
                             uint32_t nimm, uint32_t nregs, uint32_t num_constraints,
                             vpbs_gate_program** out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!out || (!code && ncode) || (!imms && nimm)) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  *out = nullptr;
  if (nregs == 0 || nregs > perm::MAX_PROG_REGS || num_constraints == 0 || num_constraints > 4096)
    return fail(ctx, VPBS_ERR_ARG, "gate program: too many registers (shared-memory register file) or constraints");
  u32 max_wire = 0, max_const = 0;
  std::vector<bool> written(nregs, false);  // straight-line code: a register must be written before it is read
  for (u32 pc = 0; pc < ncode; pc++) {
    const u64 ins = code[pc];
    const unsigned op = (unsigned)(ins & 0xff), dst = (unsigned)((ins >> 8) & 0xff);
    const unsigned kind[2] = {(unsigned)((ins >> 16) & 0xf), (unsigned)((ins >> 20) & 0xf)};
    const unsigned idx[2] = {(unsigned)((ins >> 24) & 0xffff), (unsigned)((ins >> 40) & 0xffff)};
    if (op > perm::OP_MAD || (ins >> 56)) return fail(ctx, VPBS_ERR_ARG, "gate program: bad opcode");
    const bool binary = op <= perm::OP_MUL || op == perm::OP_MAD;
    if (binary && dst >= nregs) return fail(ctx, VPBS_ERR_ARG, "gate program: bad destination register");
    if (op == perm::OP_EMIT && idx[1] >= num_constraints)
      return fail(ctx, VPBS_ERR_ARG, "gate program: constraint index out of range");
    const int nops = binary ? 2 : 1;
    for (int o = 0; o < nops; o++) {
      switch (kind[o]) {
        case perm::K_REG:
          if (idx[o] >= nregs) return fail(ctx, VPBS_ERR_ARG, "gate program: bad register");
          if (!written[idx[o]]) return fail(ctx, VPBS_ERR_ARG, "gate program: register read before it is written");
          break;
        case perm::K_WIRE: if (idx[o] + 1 > max_wire) max_wire = idx[o] + 1; break;
        case perm::K_CONST: if (idx[o] + 1 > max_const) max_const = idx[o] + 1; break;
        case perm::K_IMM: if (idx[o] >= nimm) return fail(ctx, VPBS_ERR_ARG, "gate program: bad immediate"); break;
        case perm::K_PIH: if (idx[o] >= 4) return fail(ctx, VPBS_ERR_ARG, "gate program: bad public-input index"); break;
        default: return fail(ctx, VPBS_ERR_ARG, "gate program: bad operand kind");
      }
    }
    if (op == perm::OP_MAD && !written[dst])
      return fail(ctx, VPBS_ERR_ARG, "gate program: MAD accumulates into a register that was never written");
    if (binary) written[dst] = true;
  }
  vpbs_gate_program* g = new (std::nothrow) vpbs_gate_program();
  if (!g) return fail(ctx, VPBS_ERR_OOM, "host allocation failed");
  g->device = ctx->device;
  g->ncode = ncode; g->nimm = nimm; g->nregs = nregs; g->num_constraints = num_constraints;
  g->max_wire = max_wire; g->max_const = max_const;
  std::vector<u64> ic(imms, imms + nimm);
  for (u64& v : ic) v = gl::canon(v);
  cudaError_t e = cudaMalloc((void**)&g->code, (size_t)(ncode ? ncode : 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&g->imm, (size_t)(nimm ? nimm : 1) * 8);
  if (e == cudaSuccess && ncode) e = cudaMemcpy(g->code, code, (size_t)ncode * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nimm) e = cudaMemcpy(g->imm, ic.data(), (size_t)nimm * 8, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(g->code);
    cudaFree(g->imm);
    delete g;
    return fail(ctx, e == cudaErrorMemoryAllocation ? VPBS_ERR_OOM : VPBS_ERR_CUDA,
                std::string("gate program upload: ") + cudaGetErrorString(e));
  }
  *out = g;
  return VPBS_OK;
}

void vpbs_gate_program_destroy(vpbs_gate_program* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaFree(g->code);
  cudaFree(g->imm);
  delete g;
}

}  // extern "C"

namespace {
// compute_quotient_polys up to the quotient VALUES: d_vals[c * q + i] for the leaves the batches hold
// (all of the quotient domain, or — sharded batches — the rank's row range, the rest zeroed).
int quotient_values_core(vpbs_batch* constants_sigmas, uint32_t sigmas_first_col, vpbs_batch* wires,
                         vpbs_batch* zs_pp, const uint64_t* k_is, uint32_t num_routed, uint32_t max_degree,
                         uint32_t qdb, const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas,
                         uint32_t nc, const uint64_t* const* gate_terms, const vpbs_gate_program* program,
                         const uint64_t* public_inputs_hash, u64* d_vals) {
  vpbs_ctx* ctx = wires->ctx;
  int rc;
  if (constants_sigmas->ctx != ctx || zs_pp->ctx != ctx)
    return fail(ctx, VPBS_ERR_STATE, "batches of different contexts");
  if (!k_is || !betas || !gammas || !alphas || !d_vals) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const u32 log_n = wires->log_n;
  if (nc == 0 || nc > 4) return fail(ctx, VPBS_ERR_ARG, "num_challenges must be 1..4");
  if (num_routed == 0 || max_degree < 2) return fail(ctx, VPBS_ERR_ARG, "num_routed == 0 or max_degree < 2");
  const u32 K = (num_routed + max_degree - 1) / max_degree;
  if (K > (u32)perm::MAX_CHUNKS) return fail(ctx, VPBS_ERR_ARG, "too many partial-product chunks");
  if (wires->ncols < num_routed || constants_sigmas->ncols < sigmas_first_col + num_routed || zs_pp->ncols != nc * K)
    return fail(ctx, VPBS_ERR_ARG, "batch widths do not match num_routed / the chunk count");
  if (constants_sigmas->log_n != log_n || zs_pp->log_n != log_n)
    return fail(ctx, VPBS_ERR_ARG, "batches of different degree");
  const bool sharded = wires->sharded();
  for (const vpbs_batch* b : {wires, constants_sigmas, zs_pp}) {
    if (qdb > b->rate_bits || qdb >= 5)
      return fail(ctx, VPBS_ERR_ARG, "quotient_degree_bits exceeds the batches' rate_bits");
    if (b->sharded() != sharded || b->first_leaf != wires->first_leaf || b->nleaves != wires->nleaves)
      return fail(ctx, VPBS_ERR_ARG, "the three batches must hold the same rows");
    if (sharded && b->rate_bits != qdb)
      return fail(ctx, VPBS_ERR_ARG, "sharded batches: the quotient domain must be the whole LDE (rate_bits == quotient_degree_bits)");
  }
  if (program) {
    if (gate_terms) return fail(ctx, VPBS_ERR_ARG, "gate_terms and a gate program are mutually exclusive");
    if (program->device != ctx->device) return fail(ctx, VPBS_ERR_STATE, "gate program lives on another device");
    if (program->max_wire > wires->width || program->max_const > constants_sigmas->width)
      return fail(ctx, VPBS_ERR_ARG, "gate program reads columns the batches do not have");
  }
  const unsigned log_q = log_n + qdb;
  const u64 n = 1ULL << log_n, q = 1ULL << log_q;
  const u64 k0 = sharded ? wires->first_leaf : 0, kcount = sharded ? wires->nleaves : q;

  perm::QuotientParams qp;
  memset(&qp, 0, sizeof qp);
  const u64 g_pow_n = gl::pow(gl::COSET_SHIFT, n), wr = gl::primitive_root_of_unity(qdb);
  u64 xr = 1;
  for (u32 k = 0; k < (1u << qdb); k++, xr = gl::mul(xr, wr)) {
    qp.zh[k] = gl::sub(gl::mul(g_pow_n, xr), 1);
    qp.zh_inv[k] = gl::inv(qp.zh[k]);
  }
  const u32 nterms = nc + nc * K;
  for (u32 c = 0; c < nc; c++) {
    qp.beta[c] = gl::canon(betas[c]);
    qp.gamma[c] = gl::canon(gammas[c]);
    const u64 a = gl::canon(alphas[c]);
    u64 pw = 1;
    for (u32 j = 0; j < nterms; j++, pw = gl::mul(pw, a)) qp.apow[c][j] = pw;
    qp.agate[c] = pw;
  }
  qp.n_canon = n % gl::P;

  u64 *d_k = nullptr, *d_gate = nullptr;
  if ((rc = arena_get(ctx, "idx", (size_t)num_routed * 8, (void**)&d_k))) return rc;
  if ((rc = ensure_roots(ctx, log_q))) return rc;
  std::vector<u64> kc(k_is, k_is + num_routed);
  for (u64& v : kc) v = gl::canon(v);
  CU(ctx, cudaMemcpyAsync(d_k, kc.data(), (size_t)num_routed * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (sharded) CU(ctx, cudaMemsetAsync(d_vals, 0, (size_t)nc * q * 8, ctx->stream));
  if (gate_terms) {
    if ((rc = arena_get(ctx, "gate_terms", (size_t)nc * q * 8, (void**)&d_gate))) return rc;
    for (u32 c = 0; c < nc; c++) {
      if (!gate_terms[c]) return fail(ctx, VPBS_ERR_ARG, "gate_terms[c] == NULL");
      CU(ctx, cudaMemcpyAsync(d_gate + (u64)c * q, gate_terms[c], q * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if (program) {  // evaluate_gate_constraints_base_batch, alpha-reduced, at every point of the quotient domain
    const u32 ng = program->num_constraints;
    std::vector<u64> ap((size_t)nc * ng + 4);  // alpha powers, then public_inputs_hash
    for (u32 c = 0; c < nc; c++) {
      const u64 a = gl::canon(alphas[c]);
      u64 pw = 1;
      for (u32 j = 0; j < ng; j++, pw = gl::mul(pw, a)) ap[(size_t)c * ng + j] = pw;
    }
    for (int k = 0; k < 4; k++) ap[(size_t)nc * ng + k] = public_inputs_hash ? gl::canon(public_inputs_hash[k]) : 0;
    u64* d_ap = nullptr;
    if ((rc = arena_get(ctx, "gate_apow", ap.size() * 8, (void**)&d_ap))) return rc;
    if ((rc = arena_get(ctx, "gate_terms", (size_t)nc * q * 8, (void**)&d_gate))) return rc;
    CU(ctx, cudaMemcpyAsync(d_ap, ap.data(), ap.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = (size_t)program->nregs * perm::PROG_THREADS * perm::PROG_POINTS * sizeof(u64);
    CU(ctx, cudaFuncSetAttribute((const void*)perm::gate_program_eval,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(ctx, cudaFuncSetAttribute((const void*)perm::gate_program_eval,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const u64 per_cta = (u64)perm::PROG_THREADS * perm::PROG_POINTS;
    perm::gate_program_eval<<<(unsigned)((kcount + per_cta - 1) / per_cta),
                              perm::PROG_THREADS, smem, ctx->stream>>>(
        program->code, program->ncode, program->imm, d_ap, ng, wires->leaves, wires->width,
        constants_sigmas->leaves, constants_sigmas->width, nc, log_q, k0, kcount, d_gate);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // `ap` dies with this scope
  }
  const ntt::Roots R{ctx->roots, ctx->roots_log};
  perm::quotient_permutation_terms<<<(unsigned)((kcount + 127) / 128), 128, 0, ctx->stream>>>(
      wires->leaves, wires->width, constants_sigmas->leaves, constants_sigmas->width, sigmas_first_col,
      zs_pp->leaves, zs_pp->width, d_k, num_routed, max_degree, K, nc, log_q, qdb, qp, R, d_gate, k0, kcount,
      d_vals);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  // `kc` must outlive its copy; the kernel above is ordered after it on the same stream
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// The tail of compute_quotient_polys + prove() step 7: coset_ifft(7) of the complete quotient values,
// chunks of n coefficients, committed from coefficients (the context's shard of the rows, if it shards).
int quotient_commit_core(vpbs_ctx* ctx, const u64* d_vals, u32 nc, u32 log_n, u32 qdb, u32 rate_bits,
                         u32 cap_height, uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats, uint64_t l0) {
  int rc;
  if (log_n + rate_bits > 30 || qdb >= 5) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  if (cap_height > log_n + rate_bits)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  const unsigned log_q = log_n + qdb;
  const u64 q = 1ULL << log_q;
  cudaEvent_t e0 = ctx->ev[4], e3 = ctx->ev[7];
  if (stats) cudaEventRecord(e0, ctx->stream);
  u64 *d_work = nullptr, *d_coef = nullptr;
  if ((rc = arena_get(ctx, "work", (size_t)nc * q * 8, (void**)&d_work))) return rc;
  if ((rc = arena_get(ctx, "zs_out", (size_t)nc * q * 8, (void**)&d_coef))) return rc;
  if ((rc = ensure_roots(ctx, log_n + (rate_bits > qdb ? rate_bits : qdb)))) return rc;
  // coset_ifft(7): inverse transform (natural order in and out, scaled by 1/q), then 7^-j
  if ((rc = run_transform<true>(ctx, d_vals, q, nc, log_q, d_work, Out::Natural, d_coef, q, 0, nullptr,
                                gl::inv(q % gl::P))))
    return rc;
  perm::coset_unscale<<<(unsigned)((q + 255) / 256), 256, 0, ctx->stream>>>(d_coef, q, nc, gl::inv(gl::COSET_SHIFT));
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  // chunks of n coefficients are contiguous: nc * 2^qdb columns, committed from coefficients
  const u32 ncols_out = nc << qdb;
  vpbs_batch* b = nullptr;
  if ((rc = batch_alloc(ctx, ncols_out, log_n, rate_bits, cap_height, false, true, &b))) return rc;
  Timer tm{ctx, stats != nullptr};
  rc = commit_core(ctx, d_coef, ncols_out, log_n, rate_bits, cap_height, 1, nullptr, b->first_leaf,
                   b->nleaves, b->coeffs, b->leaves, b->digests, b->own_roots(), &tm);
  if (rc) {
    vpbs_batch_destroy(b);
    return rc;
  }
  cudaError_t ce = cudaMemcpyAsync(cap_out, b->cap, b->cap_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (stats) cudaEventRecord(e3, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_CUDA, std::string("quotient batch commit: ") + cudaGetErrorString(ce));
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    cudaEventElapsedTime(&stats->total_ms, e0, e3);
  }
  *out = b;
  return VPBS_OK;
}
}  // namespace

extern "C" {

// [P2] plonk/prover.rs compute_quotient_polys (see perm::quotient_permutation_terms) + step 7's commit.
int vpbs_batch_quotient_polys(vpbs_batch* constants_sigmas, uint32_t sigmas_first_col, vpbs_batch* wires,
                              vpbs_batch* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                              uint32_t max_degree, uint32_t quotient_degree_bits, const uint64_t* betas,
                              const uint64_t* gammas, const uint64_t* alphas, uint32_t num_challenges,
                              const uint64_t* const* gate_terms, const vpbs_gate_program* program,
                              const uint64_t* public_inputs_hash, uint32_t rate_bits, uint32_t cap_height,
                              uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats) {
  if (!wires || !constants_sigmas || !zs_pp) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = wires->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (!cap_out || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  *out = nullptr;
  if (wires->sharded())
    return fail(ctx, VPBS_ERR_ARG, "sharded batches: use vpbs_batch_quotient_values, exchange the values, "
                                   "then vpbs_quotient_commit_values");
  if (num_challenges == 0 || num_challenges > 4) return fail(ctx, VPBS_ERR_ARG, "num_challenges must be 1..4");
  const uint64_t l0 = ctx->launches;
  const u64 q = 1ULL << (wires->log_n + quotient_degree_bits);
  if (quotient_degree_bits >= 5 || wires->log_n + quotient_degree_bits > 30)
    return fail(ctx, VPBS_ERR_ARG, "quotient_degree_bits exceeds the batches' rate_bits");
  u64* d_vals = nullptr;
  if ((rc = arena_get(ctx, "in", (size_t)num_challenges * q * 8, (void**)&d_vals))) return rc;
  if ((rc = quotient_values_core(constants_sigmas, sigmas_first_col, wires, zs_pp, k_is, num_routed, max_degree,
                                 quotient_degree_bits, betas, gammas, alphas, num_challenges, gate_terms,
                                 program, public_inputs_hash, d_vals)))
    return rc;
  return quotient_commit_core(ctx, d_vals, num_challenges, wires->log_n, quotient_degree_bits, rate_bits,
                              cap_height, cap_out, out, stats, l0);
}

// The two halves of the call above for a proof whose batches are sharded by row range: every rank
// computes the quotient values of ITS leaves (the rest of d_vals_out is zeroed), the ranks add their
// buffers up (an all-reduce: a value and zeros), and every rank commits its shard of the quotient batch.
int vpbs_batch_quotient_values(vpbs_batch* constants_sigmas, uint32_t sigmas_first_col, vpbs_batch* wires,
                               vpbs_batch* zs_pp, const uint64_t* k_is, uint32_t num_routed,
                               uint32_t max_degree, uint32_t quotient_degree_bits, const uint64_t* betas,
                               const uint64_t* gammas, const uint64_t* alphas, uint32_t num_challenges,
                               const uint64_t* const* gate_terms, const vpbs_gate_program* program,
                               const uint64_t* public_inputs_hash, uint64_t* d_vals_out) {
  if (!wires || !constants_sigmas || !zs_pp) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = wires->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  return quotient_values_core(constants_sigmas, sigmas_first_col, wires, zs_pp, k_is, num_routed, max_degree,
                              quotient_degree_bits, betas, gammas, alphas, num_challenges, gate_terms, program,
                              public_inputs_hash, d_vals_out);
}

int vpbs_quotient_commit_values(vpbs_ctx* ctx, const uint64_t* d_vals, uint32_t num_challenges, uint32_t log_n,
                                uint32_t quotient_degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!d_vals || !cap_out || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  *out = nullptr;
  if (num_challenges == 0 || num_challenges > 4) return fail(ctx, VPBS_ERR_ARG, "num_challenges must be 1..4");
  return quotient_commit_core(ctx, d_vals, num_challenges, log_n, quotient_degree_bits, rate_bits, cap_height,
                              cap_out, out, stats, ctx->launches);
}

int vpbs_batch_commit_dev(vpbs_ctx* ctx, const uint64_t* d_cols, uint32_t ncols, uint32_t log_n,
                          uint32_t rate_bits, uint32_t cap_height, int inputs_are_coeffs,
                          uint64_t* cap_out, vpbs_batch** out, vpbs_stats* stats) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!d_cols || ncols == 0 || !cap_out || !out) return fail(ctx, VPBS_ERR_ARG, "null pointer or ncols == 0");
  *out = nullptr;
  if (log_n + rate_bits > 30) return fail(ctx, VPBS_ERR_ARG, "log_n + rate_bits > 30");
  if (cap_height > log_n + rate_bits)
    return fail(ctx, VPBS_ERR_ARG, "cap_height should be at most log2(leaves.len())");
  vpbs_batch* b = nullptr;
  if ((rc = batch_alloc(ctx, ncols, log_n, rate_bits, cap_height, false, inputs_are_coeffs != 0, &b)))
    return rc;
  const uint64_t l0 = ctx->launches;
  Timer tm{ctx, stats != nullptr};
  rc = commit_core(ctx, d_cols, ncols, log_n, rate_bits, cap_height, inputs_are_coeffs, nullptr,
                   b->first_leaf, b->nleaves, b->coeffs, b->leaves, b->digests, b->own_roots(), &tm);
  if (rc) {
    vpbs_batch_destroy(b);
    return rc;
  }
  cudaError_t ce = cudaMemcpyAsync(cap_out, b->cap, b->cap_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    vpbs_batch_destroy(b);
    return fail(ctx, VPBS_ERR_CUDA, std::string("batch commit: ") + cudaGetErrorString(ce));
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    fill_stats(stats, tm, ctx->launches - l0);
    stats->total_ms = tm.ms(0, 3);
  }
  *out = b;
  return VPBS_OK;
}

int vpbs_batch_get_lde_rows(vpbs_batch* b, uint64_t first_index, uint64_t step, uint64_t count,
                            uint64_t* rows_out) {
  if (!b) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = b->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!rows_out) return fail(ctx, VPBS_ERR_ARG, "rows_out == NULL");
  const unsigned log_m = b->log_n + b->rate_bits;
  const u64 m = 1ULL << log_m;
  if (step == 0 || first_index >= m || (count - 1) > (m - 1 - first_index) / step)
    return fail(ctx, VPBS_ERR_ARG, "LDE index range out of bounds");
  if (b->sharded())  // every requested row must live in this shard
    for (u64 r = 0; r < count; r++) {
      const u64 leaf = reverse_bits64(first_index + r * step, log_m);
      if (leaf < b->first_leaf || leaf >= b->first_leaf + b->nleaves)
        return fail(ctx, VPBS_ERR_ARG, "LDE row outside this shard");
    }
  u64* d_rows = nullptr;
  const size_t bytes = count * (size_t)b->ncols * 8;
  // pulled in pieces of at most 64 Ki rows through the context's staging buffer
  const u64 piece = 1ULL << 16;
  if ((rc = arena_get(ctx, "rows", (size_t)(count < piece ? count : piece) * b->ncols * 8, (void**)&d_rows)))
    return rc;
  (void)bytes;
  for (u64 off = 0; off < count; off += piece) {
    const u64 cnt = count - off < piece ? count - off : piece;
    perm::gather_lde_rows<<<(unsigned)cnt, 128, 0, ctx->stream>>>(
        b->leaves, b->width, b->ncols, log_m, first_index + off * step, step, cnt, b->first_leaf, d_rows);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(rows_out + off * b->ncols, d_rows, (size_t)cnt * b->ncols * 8,
                            cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VPBS_OK;
}

int vpbs_batch_shape(vpbs_batch* b, uint32_t* ncols, uint32_t* log_n, uint32_t* rate_bits,
                     uint32_t* cap_height, uint32_t* width) {
  if (!b) return VPBS_ERR_STATE;
  if (ncols) *ncols = b->ncols;
  if (log_n) *log_n = b->log_n;
  if (rate_bits) *rate_bits = b->rate_bits;
  if (cap_height) *cap_height = b->cap_height;
  if (width) *width = b->width;
  return VPBS_OK;
}

int vpbs_batch_shard(vpbs_batch* b, uint64_t* first_leaf, uint64_t* nleaves) {
  if (!b) return VPBS_ERR_STATE;
  if (first_leaf) *first_leaf = b->first_leaf;
  if (nleaves) *nleaves = b->nleaves;
  return VPBS_OK;
}

static int batch_indices(vpbs_batch* b, const uint64_t* idx, uint64_t count, u64** d_idx) {
  vpbs_ctx* ctx = b->ctx;
  const u64 m = 1ULL << (b->log_n + b->rate_bits);
  for (uint64_t i = 0; i < count; i++) {
    if (idx[i] >= m) return fail(ctx, VPBS_ERR_ARG, "leaf index out of range");
    if (idx[i] < b->first_leaf || idx[i] >= b->first_leaf + b->nleaves)
      return fail(ctx, VPBS_ERR_ARG, "leaf index outside this shard");
  }
  int rc;
  if ((rc = arena_get(ctx, "idx", count * 8, (void**)d_idx))) return rc;
  if (!b->sharded()) {
    CU(ctx, cudaMemcpyAsync(*d_idx, idx, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    return VPBS_OK;
  }
  // the kernels index the shard's own buffers: leaf - first_leaf (the shard starts at a subtree boundary)
  std::vector<u64> local(idx, idx + count);
  for (u64& v : local) v -= b->first_leaf;
  CU(ctx, cudaMemcpyAsync(*d_idx, local.data(), count * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));  // `local` dies with this frame
  return VPBS_OK;
}

int vpbs_batch_get_leaves(vpbs_batch* b, const uint64_t* leaf_indices, uint64_t count,
                          uint64_t* rows_out) {
  if (!b) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = b->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!leaf_indices || !rows_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  u64 *d_idx = nullptr, *d_rows = nullptr;
  if ((rc = batch_indices(b, leaf_indices, count, &d_idx))) return rc;
  if ((rc = arena_get(ctx, "rows", count * (size_t)b->width * 8, (void**)&d_rows))) return rc;
  merkle::gather_rows<<<(unsigned)count, 128, 0, ctx->stream>>>(b->leaves, b->width, d_idx, count, d_rows);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(rows_out, d_rows, count * (size_t)b->width * 8, cudaMemcpyDeviceToHost,
                          ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_batch_prove(vpbs_batch* b, const uint64_t* leaf_indices, uint64_t count,
                     uint64_t* siblings_out) {
  if (!b) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = b->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  const unsigned num_layers = b->log_n + b->rate_bits - b->cap_height;
  if (count == 0 || num_layers == 0) return VPBS_OK;
  if (!leaf_indices || !siblings_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  u64 *d_idx = nullptr, *d_sib = nullptr;
  if ((rc = batch_indices(b, leaf_indices, count, &d_idx))) return rc;
  const size_t bytes = count * (size_t)num_layers * 32;
  if ((rc = arena_get(ctx, "rows", bytes, (void**)&d_sib))) return rc;
  const u64 sub_digests = 2 * (1ULL << num_layers) - 2;
  const u64 total = count * num_layers;
  merkle::gather_siblings<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(
      b->digests, d_idx, count, num_layers, sub_digests, d_sib);
  ctx->launches++;
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(siblings_out, d_sib, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

// [P2] fri/prover.rs fri_prover_query_round -> initial_trees_proof: for every oracle the leaf row and
// the Merkle path at the same query indices.  One index upload, two gathers per batch, one
// synchronisation for all batches (eight separate get_leaves / prove calls cost the step 0.4 ms).
int vpbs_batches_open(vpbs_batch* const* batches, uint32_t nbatches, const uint64_t* leaf_indices,
                      uint64_t count, uint64_t* const* rows_out, uint64_t* const* siblings_out) {
  if (!batches || nbatches == 0 || !batches[0]) return VPBS_ERR_STATE;
  vpbs_batch* b0 = batches[0];
  vpbs_ctx* ctx = b0->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (count == 0) return VPBS_OK;
  if (!leaf_indices || !rows_out || !siblings_out) return fail(ctx, VPBS_ERR_ARG, "null pointer");
  const unsigned num_layers = b0->log_n + b0->rate_bits - b0->cap_height;
  size_t total = 0;
  for (u32 k = 0; k < nbatches; k++) {
    const vpbs_batch* b = batches[k];
    if (!b || b->ctx != ctx) return fail(ctx, VPBS_ERR_STATE, "batches of different contexts");
    if (b->log_n != b0->log_n || b->rate_bits != b0->rate_bits || b->cap_height != b0->cap_height ||
        b->first_leaf != b0->first_leaf || b->nleaves != b0->nleaves)
      return fail(ctx, VPBS_ERR_ARG, "batches opened together must have the same tree shape and shard");
    if (!rows_out[k] || (num_layers && !siblings_out[k])) return fail(ctx, VPBS_ERR_ARG, "null output pointer");
    total += count * ((size_t)b->width * 8 + (size_t)num_layers * 32) + 32;  // + alignment slack per region
  }
  u64 *d_idx = nullptr, *d_buf = nullptr;
  if ((rc = batch_indices(b0, leaf_indices, count, &d_idx))) return rc;
  if ((rc = arena_get(ctx, "rows", total, (void**)&d_buf))) return rc;
  const u64 sub_digests = 2 * (1ULL << num_layers) - 2;
  size_t off = 0;
  for (u32 k = 0; k < nbatches; k++) {
    const vpbs_batch* b = batches[k];
    u64* d_rows = d_buf + off / 8;
    const size_t row_bytes = count * (size_t)b->width * 8, sib_bytes = count * (size_t)num_layers * 32;
    merkle::gather_rows<<<(unsigned)count, 128, 0, ctx->stream>>>(b->leaves, b->width, d_idx, count, d_rows);
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(rows_out[k], d_rows, row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    off += (row_bytes + 15) & ~(size_t)15;  // gather_siblings moves hashes as 16-byte halves
    if (num_layers) {
      u64* d_sib = d_buf + off / 8;
      const u64 tot = count * num_layers;
      merkle::gather_siblings<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(
          b->digests, d_idx, count, num_layers, sub_digests, d_sib);
      ctx->launches++;
      CU(ctx, cudaMemcpyAsync(siblings_out[k], d_sib, sib_bytes, cudaMemcpyDeviceToHost, ctx->stream));
      off += (sib_bytes + 15) & ~(size_t)15;
    }
  }
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

int vpbs_batch_download(vpbs_batch* b, uint64_t* const* coeffs_out, uint64_t* leaves_out,
                        uint64_t* digests_out) {
  if (!b) return VPBS_ERR_STATE;
  vpbs_ctx* ctx = b->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  const u64 n = 1ULL << b->log_n, m = b->nleaves;  // a sharded batch downloads its own rows / digests
  const u64 ndig = 2 * (m - (m >> (b->log_n + b->rate_bits - b->cap_height)));
  if (coeffs_out)
    for (u32 c = 0; c < b->ncols; c++)
      if (coeffs_out[c])
        CU(ctx, cudaMemcpyAsync(coeffs_out[c], b->coeffs + (u64)c * n, n * 8, cudaMemcpyDeviceToHost,
                                ctx->stream));
  if (leaves_out)
    CU(ctx, cudaMemcpyAsync(leaves_out, b->leaves, (size_t)m * b->width * 8, cudaMemcpyDeviceToHost,
                            ctx->stream));
  if (digests_out && ndig)
    CU(ctx, cudaMemcpyAsync(digests_out, b->digests, ndig * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return VPBS_OK;
}

}  // extern "C"
