// ntt_bench.cu — standalone timing of the 2^16-row commit's transform kernels (developer tool for
// kernel iterations; bench.py is the number of record): IFFT of C columns, then the 2^rate_bits
// coset LDE blocks written as leaf rows — the launch sequence of run_transform / commit_core in
// csrc/vpbs_commit.cu (persistent passes, twiddle and scale split across the two passes, LDE blocks
// alternating between two streams).  Prints the time of each phase and a checksum of coefficients
// and leaves so that variants can be compared bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../csrc -o ntt_bench ntt_bench.cu
//   [-DNTT_BENCH_ONE_STREAM] [-DNTT_BENCH_CTAS_PER_SM=k]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "ntt.cuh"

using gl::u64;

#ifndef NTT_BENCH_CTAS_PER_SM
#define NTT_BENCH_CTAS_PER_SM 3
#endif

__global__ void fill(u64* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    u64 z = i * 0x9E3779B97F4A7C15ULL + 12345;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    p[i] = gl::canon(z ^ (z >> 31));
  }
}
__global__ void checksum(const u64* p, size_t n, u64* out) {
  u64 acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += p[i] * (2 * i + 1);
  atomicAdd((unsigned long long*)out, (unsigned long long)acc);
}

template <typename K>
static void big_smem(K k, size_t bytes) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int main(int argc, char** argv) {
  const unsigned C = argc > 1 ? atoi(argv[1]) : 128, log_n = 16, r = 3;
  const u64 n = 1ULL << log_n, m = n << r;
  u64 *cols, *coeffs, *work, *work2, *leaves, *roots, *coset, *sum;
  cudaMalloc(&cols, C * n * 8);
  cudaMalloc(&coeffs, C * n * 8);
  cudaMalloc(&work, C * n * 8);
  cudaMalloc(&work2, C * n * 8);
  cudaMalloc(&leaves, C * m * 8);
  cudaMalloc(&roots, (m / 2) * 8);
  cudaMalloc(&coset, m * 8);
  cudaMalloc(&sum, 16);
  cudaMemset(sum, 0, 16);
  fill<<<(unsigned)((C * n + 255) / 256), 256>>>(cols, C * n);
  ntt::fill_roots<<<(unsigned)((m / 2 + 255) / 256), 256>>>(roots, log_n + r);
  ntt::fill_coset_powers<<<(unsigned)((m + 255) / 256), 256>>>(coset, log_n, r, gl::COSET_SHIFT);
  const ntt::Roots R{roots, log_n + r};
  const u64 n_inv = gl::inv(n);
  big_smem(ntt::pass_strided_r16p<true, true>, ntt::R16P_STRIDED_SMEM);
  big_smem(ntt::pass_strided_r16p<false, false>, ntt::R16P_STRIDED_SMEM);
  big_smem(ntt::pass_final_r16p<true, ntt::STORE_NATURAL>, ntt::R16P_FINAL_SMEM);
  big_smem(ntt::pass_final_r16p<false, ntt::STORE_LEAF>, ntt::R16P_FINAL_SMEM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const unsigned cap = (unsigned)sms * NTT_BENCH_CTAS_PER_SM;
  cudaEvent_t e[3], fork, join[2];
  for (auto& x : e) cudaEventCreate(&x);
  cudaStream_t st[2];
  for (auto& x : st) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  for (auto& x : join) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
  const unsigned tx1 = (unsigned)(n >> 12), nt1 = tx1 * C;                 // strided: 16 x C tiles
  const unsigned txn = (unsigned)(n >> 12), ntn = txn * C;                 // final, natural store
  const unsigned txl = (unsigned)(n >> 8), ntl = txl * ((C + 15) / 16);    // final, leaf store
  auto grid = [&](unsigned nt) { return nt < cap ? nt : cap; };
  float best_i = 1e9f, best_f = 1e9f;
  for (int it = 0; it < 12; it++) {
    cudaEventRecord(e[0]);
    ntt::pass_strided_r16p<true, true><<<grid(nt1), ntt::THREADS, ntt::R16P_STRIDED_SMEM>>>(
        cols, n, work, n, log_n, nullptr, R, tx1, nt1);
    ntt::pass_final_r16p<true, ntt::STORE_NATURAL><<<grid(ntn), ntt::THREADS, ntt::R16P_FINAL_SMEM>>>(
        work, n, C, coeffs, n, 0, log_n, nullptr, n_inv, R, 0u, nullptr, txn, ntn);
    cudaEventRecord(e[1]);
    cudaEventRecord(fork, 0);
    cudaStreamWaitEvent(st[0], fork, 0);
    cudaStreamWaitEvent(st[1], fork, 0);
    for (u64 b = 0; b < (1u << r); b++) {
#ifdef NTT_BENCH_ONE_STREAM
      cudaStream_t q = st[0];
      u64* wk = work;
#else
      cudaStream_t q = st[b & 1];
      u64* wk = (b & 1) ? work2 : work;
#endif
      ntt::pass_strided_r16p<false, false><<<grid(nt1), ntt::THREADS, ntt::R16P_STRIDED_SMEM, q>>>(
          coeffs, n, wk, n, log_n, coset + (b << log_n), R, tx1, nt1);
      ntt::pass_final_r16p<false, ntt::STORE_LEAF><<<grid(ntl), ntt::THREADS, ntt::R16P_FINAL_SMEM, q>>>(
          wk, n, C, leaves, C, b << log_n, log_n, nullptr, 1, R, log_n, coset + (b << log_n), txl, ntl);
    }
    cudaEventRecord(join[0], st[0]);
    cudaEventRecord(join[1], st[1]);
    cudaStreamWaitEvent(0, join[0], 0);
    cudaStreamWaitEvent(0, join[1], 0);
    cudaEventRecord(e[2]);
    cudaEventSynchronize(e[2]);
    float a, b;
    cudaEventElapsedTime(&a, e[0], e[1]);
    cudaEventElapsedTime(&b, e[1], e[2]);
    if (a < best_i) best_i = a;
    if (b < best_f) best_f = b;
  }
  checksum<<<592, 256>>>(coeffs, C * n, sum);
  checksum<<<592, 256>>>(leaves, C * m, sum + 1);
  u64 h[2];
  cudaMemcpy(h, sum, 16, cudaMemcpyDeviceToHost);
  printf("C=%u  ifft %.3f ms  lde %.3f ms  coeffs=%016llx leaves=%016llx  err=%s\n", C, best_i, best_f,
         (unsigned long long)h[0], (unsigned long long)h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
