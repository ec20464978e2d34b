// poseidon_bench.cu — standalone timing of merkle::hash_leaves on 2^19 x 128 synthetic leaves
// (developer tool for kernel iterations; bench.py is the number of record).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../csrc [-DVPBS_HASH_MIN_BLOCKS=k]
//        -o poseidon_bench poseidon_bench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "merkle.cuh"

__global__ void fill(unsigned long long* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    unsigned long long z = i * 0x9E3779B97F4A7C15ULL + 12345;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    p[i] = z ^ (z >> 31);
  }
}

int main(int argc, char** argv) {
  const unsigned width = argc > 1 ? atoi(argv[1]) : 128;
  const size_t nleaves = 1u << 19;
  gl::u64 *leaves, *dig;
  cudaMalloc(&leaves, nleaves * width * 8);
  cudaMalloc(&dig, nleaves * 32 * 2);
  fill<<<(unsigned)((nleaves * width + 255) / 256), 256>>>((unsigned long long*)leaves, nleaves * width);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int T = VPBS_HASH_THREADS;
  float best = 1e9;
  for (int it = 0; it < 6; it++) {
    cudaEventRecord(e0);
    merkle::hash_leaves<<<(unsigned)((nleaves + T - 1) / T), T>>>(leaves, nleaves, width, dig, 15,
                                                                  2 * (1u << 15) - 2, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  std::vector<gl::u64> h(8);
  cudaMemcpy(h.data(), dig, 64, cudaMemcpyDeviceToHost);
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, merkle::hash_leaves);
  printf("threads=%d minblocks=%d regs=%d local=%zu  best %.3f ms  (%.1f ns/perm)  dig0=%016llx err=%s\n", T,
         VPBS_HASH_MIN_BLOCKS, a.numRegs, a.localSizeBytes, best,
         best * 1e6 / (nleaves * ((width + 7) / 8)), (unsigned long long)h[0],
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
