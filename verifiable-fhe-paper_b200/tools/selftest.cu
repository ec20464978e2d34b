// selftest.cu — device-vs-host unit checks of the hand-written PTX field primitives.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../csrc -o selftest selftest.cu && ./selftest
// Host side uses the portable C++ paths of gl64.cuh (unsigned __int128 free) as the reference.
#include <cuda_runtime.h>

#include <cstdio>
#include <random>
#include <vector>

#include "poseidon.cuh"

using gl::u64;
typedef unsigned __int128 u128;

__global__ void k_mul(const u64* a, const u64* b, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = gl::mul_lazy(a[i], b[i]);
}
__global__ void k_red96(const u64* a, const u64* b, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = poseidon::reduce96(a[i], b[i]);
}
__global__ void k_sbox(const u64* a, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = poseidon::sbox7(a[i]);
}
__global__ void k_mds(const u64* a, u64* out, int n) {  // n states of 12; rc_next = 0
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 s[12];
  for (int k = 0; k < 12; k++) s[k] = a[i * 12 + k];
  poseidon::mds_add_rc(s, 30);
  for (int k = 0; k < 12; k++) out[i * 12 + k] = s[k];
}

static u64 modp(u128 x) { return (u64)(x % (u128)gl::P); }
static const u64 HOST_RC[360] = {
#include "poseidon_rc.inc"
};
static void host_permute(u64* s) {
  const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  for (int i = 0; i < 12; i++) s[i] = modp(s[i]);
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = modp((u128)s[i] + HOST_RC[12 * r + i]);
    for (int i = 0; i < ((r < 4 || r >= 26) ? 12 : 1); i++) {
      u64 x = s[i], x2 = modp((u128)x * x), x4 = modp((u128)x2 * x2), x3 = modp((u128)x2 * x);
      s[i] = modp((u128)x3 * x4);
    }
    u64 o[12];
    for (int q = 0; q < 12; q++) {
      u128 acc = 0;
      for (int k = 0; k < 12; k++) acc += (u128)s[(k + q) % 12] * C[k];
      if (q == 0) acc += (u128)s[0] * 8;
      o[q] = modp(acc);
    }
    for (int i = 0; i < 12; i++) s[i] = o[i];
  }
}
__global__ void k_perm(u64* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 s[12];
  for (int k = 0; k < 12; k++) s[k] = a[i * 12 + k];
  poseidon::permute_lazy(s);
  for (int k = 0; k < 12; k++) a[i * 12 + k] = gl::canon(s[k]);
}

int main() {
  const int n = 1 << 16;
  std::mt19937_64 rng(1);
  const u64 edge[] = {0, 1, 2, gl::P - 1, gl::P, gl::P + 1, ~0ULL, 0xFFFFFFFFULL, 0x100000000ULL,
                      0xFFFFFFFF00000000ULL, 1ULL << 63};
  std::vector<u64> a(n * 12), b(n * 12), out(n * 12);
  for (auto& x : a) x = (rng() % 4 == 0) ? edge[rng() % 11] : rng();
  for (auto& x : b) x = (rng() % 4 == 0) ? edge[rng() % 11] : rng();
  u64 *da, *db, *dout;
  cudaMalloc(&da, n * 12 * 8); cudaMalloc(&db, n * 12 * 8); cudaMalloc(&dout, n * 12 * 8);
  cudaMemcpy(da, a.data(), n * 12 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), n * 12 * 8, cudaMemcpyHostToDevice);
  int bad = 0;

  k_mul<<<n / 256, 256>>>(da, db, dout, n);
  cudaMemcpy(out.data(), dout, n * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++)
    if (gl::canon(out[i]) != modp((u128)a[i] * b[i])) {
      if (bad++ < 5) printf("mul mismatch a=%016llx b=%016llx got=%016llx want=%016llx\n",
                            (unsigned long long)a[i], (unsigned long long)b[i],
                            (unsigned long long)out[i], (unsigned long long)modp((u128)a[i] * b[i]));
    }
  printf("mul_lazy: %d bad\n", bad);

  int bad2 = 0;
  std::vector<u64> lo(n), hi(n);
  for (int i = 0; i < n; i++) { lo[i] = rng() >> 21; hi[i] = rng() >> 21; if (i % 7 == 0) hi[i] |= 0xFFFFFFFFULL; }
  cudaMemcpy(da, lo.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hi.data(), n * 8, cudaMemcpyHostToDevice);
  k_red96<<<n / 256, 256>>>(da, db, dout, n);
  cudaMemcpy(out.data(), dout, n * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++)
    if (gl::canon(out[i]) != modp((u128)lo[i] + ((u128)hi[i] << 32))) {
      if (bad2++ < 5) printf("reduce96 mismatch lo=%016llx hi=%016llx got=%016llx\n",
                             (unsigned long long)lo[i], (unsigned long long)hi[i], (unsigned long long)out[i]);
    }
  printf("reduce96: %d bad\n", bad2);

  int bad3 = 0;
  cudaMemcpy(da, a.data(), n * 12 * 8, cudaMemcpyHostToDevice);
  k_sbox<<<n / 256, 256>>>(da, dout, n);
  cudaMemcpy(out.data(), dout, n * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++) {
    u64 x = modp(a[i]), x2 = modp((u128)x * x), x4 = modp((u128)x2 * x2), x3 = modp((u128)x2 * x);
    if (gl::canon(out[i]) != modp((u128)x3 * x4)) bad3++;
  }
  printf("sbox7: %d bad\n", bad3);

  int bad4 = 0;
  const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  k_mds<<<n / 256, 256>>>(da, dout, n);
  cudaMemcpy(out.data(), dout, n * 12 * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++)
    for (int r = 0; r < 12; r++) {
      u128 acc = 0;
      for (int k = 0; k < 12; k++) acc += (u128)modp(a[i * 12 + (k + r) % 12]) * C[k];
      if (r == 0) acc += (u128)modp(a[i * 12]) * 8;
      if (gl::canon(out[i * 12 + r]) != modp(acc)) {
        if (bad4++ < 5) printf("mds mismatch state %d lane %d\n", i, r);
      }
    }
  printf("mds: %d bad\n", bad4);
  int bad5 = 0;
  cudaMemcpy(da, a.data(), n * 12 * 8, cudaMemcpyHostToDevice);
  k_perm<<<n / 256, 256>>>(da, n);
  cudaMemcpy(out.data(), da, n * 12 * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 4096; i++) {
    u64 h[12];
    for (int k = 0; k < 12; k++) h[k] = a[i * 12 + k];
    host_permute(h);
    for (int k = 0; k < 12; k++)
      if (h[k] != out[i * 12 + k]) { if (bad5++ < 3) printf("permutation mismatch state %d lane %d\n", i, k); }
  }
  printf("permutation: %d bad\n", bad5);
  cudaError_t e = cudaDeviceSynchronize();
  printf("cuda: %s\n", cudaGetErrorString(e));
  return bad + bad2 + bad3 + bad4 + bad5 ? 1 : 0;
}
