// dmma_bench.cu — does mma.sync.m8n8k4.f64 (DMMA) give the MDS layer cheaper issue slots than DFMA?
// Measures, per SM sub-partition, cycles per warp instruction for: DFMA alone, DMMA alone, the
// Goldilocks multiply alone, and the multiply interleaved with DFMA / DMMA (does the FP64 work hide
// behind the integer issue stream?).  Developer tool; prints one line per experiment.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../csrc -o dmma_bench dmma_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "gl64.cuh"
using gl::u64;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void dmma16x8x8(double (&c)[16], int o, double a, double b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[o]), "+d"(c[o + 1]), "+d"(c[o + 2]), "+d"(c[o + 3])
               : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
}
__device__ __forceinline__ void dmma16x8x16(double (&c)[16], int o, double a, double b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[o]), "+d"(c[o + 1]), "+d"(c[o + 2]), "+d"(c[o + 3])
               : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
}

template <int MODE>
__global__ void k(u64* out, int iters, long long* cyc) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  double a = 1.0 + (tid & 3), b = 2.0 + (tid & 7);
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = 4503599627370496.0 + i;
  u64 x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = 0x123456789abcdef1ULL * (tid + i + 1);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // 16 DFMA
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(a, b, c[i]);
    } else if (MODE == 1) {  // 8 DMMA (= 64 DFMA worth of MACs per lane... 8 x 256 MACs per warp)
#pragma unroll
      for (int i = 0; i < 8; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
    } else if (MODE == 2) {  // 8 modular multiplies (4 chains)
#pragma unroll
      for (int i = 0; i < 8; i++) x[i & 3] = gl::mul_lazy(x[i & 3], x[(i + 1) & 3]);
    } else if (MODE == 3) {  // 8 multiplies + 16 DFMA
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x[i & 3] = gl::mul_lazy(x[i & 3], x[(i + 1) & 3]);
        c[2 * i] = fma(a, b, c[2 * i]);
        c[2 * i + 1] = fma(a, b, c[2 * i + 1]);
      }
    } else if (MODE == 4) {  // 8 multiplies + 2 DMMA (= 16 DFMA-equivalents per lane: 2 x 256 MACs / 32)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x[i & 3] = gl::mul_lazy(x[i & 3], x[(i + 1) & 3]);
        if (i == 2) dmma(c[0], c[1], a, b);
        if (i == 6) dmma(c[2], c[3], a, b);
      }
    } else if (MODE == 6) {  // 4 DMMA m16n8k8 (1024 MACs each = 128 DFMA of MACs in total)
#pragma unroll
      for (int i = 0; i < 4; i++) dmma16x8x8(c, 4 * i, a, b);
    } else if (MODE == 7) {  // 4 DMMA m16n8k16 (2048 MACs each = 256 DFMA of MACs in total)
#pragma unroll
      for (int i = 0; i < 4; i++) dmma16x8x16(c, 4 * i, a, b);
    } else if (MODE == 8) {  // 8 multiplies + 1 DMMA m16n8k16 (= 64 DFMA of MACs)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x[i & 3] = gl::mul_lazy(x[i & 3], x[(i + 1) & 3]);
        if (i == 3) dmma16x8x16(c, 0, a, b);
      }
    } else if (MODE == 5) {  // 8 multiplies + 8 DMMA
#pragma unroll
      for (int i = 0; i < 8; i++) {
        x[i & 3] = gl::mul_lazy(x[i & 3], x[(i + 1) & 3]);
        dmma(c[2 * i], c[2 * i + 1], a, b);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[tid] = x[0] ^ x[1] ^ x[2] ^ x[3] ^ (u64)__double_as_longlong(s);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm, int per_iter_instr) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 128, blocks = sms * warps_per_sm * 32 / threads, iters = 4000;
  u64* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaMalloc(&cyc, (size_t)blocks * 8);
  k<MODE><<<blocks, threads>>>(out, 10, cyc);
  k<MODE><<<blocks, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, (size_t)blocks * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; i++) mean += h[i];
  mean /= blocks;
  const double warps_per_smsp = warps_per_sm / 4.0;
  // cycles an SMSP spends per loop iteration of ONE warp = elapsed / iters / warps sharing the SMSP
  printf("%-34s warps/SM %2d  cycles/iter/warp-slot %8.1f  (%s)  err=%s\n", name, warps_per_sm,
         mean / iters / warps_per_smsp, "SMSP cycles per iteration of one warp",
         cudaGetErrorString(cudaGetLastError()));
  (void)per_iter_instr;
  cudaFree(out);
  cudaFree(cyc);
  delete[] h;
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("16 DFMA", w, 16);
    run<1>("8 DMMA m8n8k4 (=64 DFMA of MACs)", w, 8);
    run<2>("8 modmul", w, 8 * 16);
    run<3>("8 modmul + 16 DFMA", w, 0);
    run<4>("8 modmul + 2 DMMA (=16 DFMA MACs)", w, 0);
    run<5>("8 modmul + 8 DMMA (=64 DFMA MACs)", w, 0);
    run<6>("4 DMMA m16n8k8 (=128 DFMA MACs)", w, 0);
    run<7>("4 DMMA m16n8k16 (=256 DFMA MACs)", w, 0);
    run<8>("8 modmul + 1 DMMA m16n8k16 (=64)", w, 0);
  }
  return 0;
}
