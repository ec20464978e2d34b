// pipe_overlap.cu — how far do the B200 math pipes that Poseidon uses overlap?
// The leaf-hashing kernel keeps three pipes 40-60 % busy each (ncu: ALU, FMA-heavy, FP64) while its
// issue slots are only 65 % used, and its time does not move when work is shifted between the pipes
// (DESIGN.md §4.2).  This tool measures, per SM sub-partition, the cycles one loop iteration takes for
// every combination of four instruction streams — 16 independent instructions of each kind per
// iteration, 8 chains per kind so that latency is covered — with 4 and 5 warps per sub-partition:
//   F = DFMA (FP64 pipe)          A = LOP3 (ALU pipe)
//   W = IMAD.WIDE.U32 with a 64-bit addend (FMA-heavy pipe)   M = IMAD (32-bit, FMA pipe)
// If the pipes overlapped freely a combination would cost the maximum of its parts; if one shared
// resource (dispatch port / register-file bandwidth) serialised them, the sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipe_overlap pipe_overlap.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

constexpr int ITER = 2048;

template <int MASK>
__global__ void k(unsigned long long* out, long long* cyc, unsigned seed) {
  const unsigned x = threadIdx.x * 2654435761u + seed;
  double d[8];
  unsigned a[8], m[8];
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    d[i] = 1.0 + i + (x & 3);
    a[i] = x + i;
    m[i] = x ^ (i * 77);
    w[i] = (unsigned long long)x * (i + 3);
  }
  const double c1 = 1.0 + 1e-9 * seed, c2 = 1e-7 * seed;
  const unsigned y = ~seed, z = seed | 5;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MASK & 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));
        if (MASK & 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(y), "r"(z));
        if (MASK & 4)
          asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(w[i]) : "r"(z));
        if (MASK & 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[i]) : "r"(z), "r"(y));
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  unsigned long long acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= (unsigned long long)__double_as_longlong(d[i]) ^ a[i] ^ m[i] ^ w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MASK>
double run(int threads, unsigned long long* out, long long* cyc, int sms) {
  k<MASK><<<sms, threads>>>(out, cyc, 12345u);
  cudaDeviceSynchronize();
  k<MASK><<<sms, threads>>>(out, cyc, 12345u);
  cudaDeviceSynchronize();
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long v : h) mean += (double)v;
  return mean / sms / ITER;  // cycles per iteration (every sub-partition runs threads/128 warps)
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)sms * 1024 * 8);
  cudaMalloc(&cyc, sms * sizeof(long long));
  const char* names[16] = {"-", "F", "A", "F+A", "W", "F+W", "A+W", "F+A+W", "M", "F+M", "A+M", "F+A+M",
                           "W+M", "F+W+M", "A+W+M", "F+A+W+M"};
  for (int threads : {512, 640}) {
    const int warps = threads / 128;  // per sub-partition
    double c[16] = {};
#define RUN(M) c[M] = run<M>(threads, out, cyc, sms);
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
#undef RUN
    for (int mk = 1; mk < 16; mk++) {
      double mx = 0, sum = 0;
      int kinds = 0;
      for (int b = 0; b < 4; b++)
        if (mk & (1 << b)) {
          mx = c[1 << b] > mx ? c[1 << b] : mx;
          sum += c[1 << b];
          kinds++;
        }
      // 16 instructions of each kind per iteration and warp
      printf("{\"bench\": \"pipe_overlap\", \"warps_per_smsp\": %d, \"streams\": \"%s\", \"cycles_per_iter\": %.1f, "
             "\"cycles_per_warp_instr\": %.3f, \"max_of_parts\": %.1f, \"sum_of_parts\": %.1f}\n",
             warps, names[mk], c[mk], c[mk] / (16.0 * kinds * warps), mx, sum);
    }
  }
  printf("{\"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
