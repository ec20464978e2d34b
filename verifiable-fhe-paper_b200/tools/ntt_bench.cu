// ntt_bench.cu — standalone timing of the 2^16-row commit's transform kernels (developer tool for
// kernel iterations; bench.py is the number of record): IFFT of C columns, then the 2^rate_bits
// coset LDE blocks written as leaf rows, exactly the launch sequence of run_transform in
// csrc/vpbs_commit.cu.  Prints the time of each phase and a checksum of coefficients and leaves so
// that variants can be compared bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../csrc -o ntt_bench ntt_bench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "ntt.cuh"

using gl::u64;

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

int main(int argc, char** argv) {
  const unsigned C = argc > 1 ? atoi(argv[1]) : 128, log_n = 16, r = 3;
  const u64 n = 1ULL << log_n, m = n << r;
  u64 *cols, *coeffs, *work, *leaves, *roots, *coset, *sum;
  cudaMalloc(&cols, C * n * 8);
  cudaMalloc(&coeffs, C * n * 8);
  cudaMalloc(&work, C * n * 8);
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
  cudaEvent_t e[3];
  for (auto& x : e) cudaEventCreate(&x);
  cudaFuncSetAttribute(ntt::pass_strided_r16p<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)ntt::R16P_STRIDED_SMEM);
  cudaFuncSetAttribute(ntt::pass_final_r16p<false, ntt::STORE_LEAF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)ntt::R16P_FINAL_SMEM);
  u64* work2;
  cudaMalloc(&work2, C * n * 8);
  cudaStream_t st[2];
  cudaEvent_t fork, join[2];
  for (auto& x : st) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
  for (auto& x : join) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
  (void)work2; (void)st; (void)fork; (void)join;
  float best_i = 1e9f, best_f = 1e9f;
  for (int it = 0; it < 12; it++) {
    cudaEventRecord(e[0]);
    {
      dim3 g1((unsigned)(n >> 12), C), g2((unsigned)(n >> 12), C);
      ntt::pass_strided_r16<true><<<g1, ntt::THREADS>>>(cols, n, work, n, log_n, nullptr, R);
      ntt::pass_final_r16<true, ntt::STORE_NATURAL><<<g2, ntt::THREADS>>>(work, n, C, coeffs, n, 0, log_n,
                                                                        nullptr, n_inv, R);
    }
    cudaEventRecord(e[1]);
#ifdef NTT_BENCH_TWO_STREAMS
    // blocks alternate between two streams and two work buffers: the tail wave of one kernel
    // (2048 CTAs = 3.46 waves of 592) is filled by the other stream's CTAs
    cudaEventRecord(fork, 0);
    cudaStreamWaitEvent(st[0], fork, 0);
    cudaStreamWaitEvent(st[1], fork, 0);
    for (u64 b = 0; b < (1u << r); b++) {
      dim3 g1((unsigned)(n >> 12), C), g2((unsigned)(n >> 8), (C + 15) / 16);
      cudaStream_t q = st[b & 1];
      u64* wk = (b & 1) ? work2 : work;
#ifdef NTT_BENCH_PERSIST
      const unsigned nt1 = g1.x * g1.y, nt2 = g2.x * g2.y, cap = 148 * NTT_BENCH_PERSIST;
      ntt::pass_strided_r16p<false, false><<<nt1 < cap ? nt1 : cap, ntt::THREADS, ntt::R16P_STRIDED_SMEM, q>>>(
          coeffs, n, wk, n, log_n, coset + (b << log_n), R, g1.x, nt1);
      ntt::pass_final_r16p<false, ntt::STORE_LEAF><<<nt2 < cap ? nt2 : cap, ntt::THREADS, ntt::R16P_FINAL_SMEM, q>>>(
          wk, n, C, leaves, C, b << log_n, log_n, nullptr, 1, R, log_n, coset + (b << log_n), g2.x, nt2);
#else
      ntt::pass_strided_r16<false, false><<<g1, ntt::THREADS, 0, q>>>(coeffs, n, wk, n, log_n,
                                                                      coset + (b << log_n), R);
      ntt::pass_final_r16<false, ntt::STORE_LEAF><<<g2, ntt::THREADS, 0, q>>>(wk, n, C, leaves, C, b << log_n,
                                                                            log_n, nullptr, 1, R, log_n, coset + (b << log_n));
#endif
    }
    cudaEventRecord(join[0], st[0]);
    cudaEventRecord(join[1], st[1]);
    cudaStreamWaitEvent(0, join[0], 0);
    cudaStreamWaitEvent(0, join[1], 0);
    if (0)
#endif
    for (u64 b = 0; b < (1u << r); b++) {
      dim3 g1((unsigned)(n >> 12), C), g2((unsigned)(n >> 8), (C + 15) / 16);
#ifdef NTT_BENCH_TW_AT_STORE
      ntt::pass_strided_r16<false><<<g1, ntt::THREADS>>>(coeffs, n, work, n, log_n, coset + (b << log_n), R);
      ntt::pass_final_r16<false, ntt::STORE_LEAF><<<g2, ntt::THREADS>>>(work, n, C, leaves, C, b << log_n,
                                                                      log_n, nullptr, 1, R);
#else
      ntt::pass_strided_r16<false, false><<<g1, ntt::THREADS>>>(coeffs, n, work, n, log_n,
                                                                coset + (b << log_n), R);
      ntt::pass_final_r16<false, ntt::STORE_LEAF><<<g2, ntt::THREADS>>>(work, n, C, leaves, C, b << log_n,
                                                                      log_n, nullptr, 1, R, log_n, coset + (b << log_n));
#endif
    }
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
