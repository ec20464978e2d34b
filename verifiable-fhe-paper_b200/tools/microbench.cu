// microbench.cu — integer-pipe issue rates on B200 that size the Poseidon roofline (SURVEY.md
// §8(d): "re-measure the IMAD.WIDE rate first").  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
// Each kernel runs ITER x UNROLL independent-chain instructions per thread, one full wave
// (148 SMs x 2048 threads), and reports lane-ops per clock per SM from in-kernel clock64().
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CHECK(x)                                                              \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                               \
    }                                                                         \
  } while (0)

constexpr int ITER = 4096;

struct Clk {
  long long t0, t1;
};

#define KERNEL_PROLOGUE                                  \
  unsigned x = threadIdx.x * 2654435761u + seed;         \
  unsigned long long a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7; \
  __syncthreads();                                       \
  long long t0 = clock64();

#define KERNEL_EPILOGUE                                  \
  long long t1 = clock64();                              \
  if (threadIdx.x == 0) clk[blockIdx.x] = Clk{t0, t1};   \
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;

// 8 independent IMAD.WIDE.U32 accumulate chains
__global__ void k_imad_wide(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(a) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %2, %0;}" : "+l"(a) : "r"(x), "r"(seed));
    OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#undef OP
  }
  KERNEL_EPILOGUE
}
// 8 independent IMAD.WIDE.U32 with a small immediate multiplier (the MDS form)
__global__ void k_imad_wide_imm(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(a) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, 41, %0;}" : "+l"(a) : "r"(x));
    OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#undef OP
  }
  KERNEL_EPILOGUE
}
// 8 independent 32-bit IMAD chains
__global__ void k_imad32(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned b0 = a0, b1 = a1, b2 = a2, b3 = a3, b4 = a4, b5 = a5, b6 = a6, b7 = a7;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(b) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(x), "r"(seed));
    OP(b0) OP(b1) OP(b2) OP(b3) OP(b4) OP(b5) OP(b6) OP(b7)
#undef OP
  }
  a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
  KERNEL_EPILOGUE
}
// 8 independent IADD3 chains
__global__ void k_iadd3(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned b0 = a0, b1 = a1, b2 = a2, b3 = a3, b4 = a4, b5 = a5, b6 = a6, b7 = a7;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(b) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(b) : "r"(x), "r"(seed));
    OP(b0) OP(b1) OP(b2) OP(b3) OP(b4) OP(b5) OP(b6) OP(b7)
#undef OP
  }
  a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
  KERNEL_EPILOGUE
}
// 64-bit add with carry (IADD3 + IADD3.X), 8 chains: counts 2 instructions per op
__global__ void k_add64(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned long long y = ((unsigned long long)seed << 32) | x;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(a) asm volatile("add.u64 %0, %0, %1;" : "+l"(a) : "l"(y));
    OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#undef OP
  }
  KERNEL_EPILOGUE
}
// mixed: 4 IMAD.WIDE chains + 4 LOP3 (alu pipe) chains
__global__ void k_mix_wide_alu(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned b4 = a4, b5 = a5, b6 = a6, b7 = a7;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OPW(a) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %2, %0;}" : "+l"(a) : "r"(x), "r"(seed));
#define OPA(b) asm volatile("{.reg .u32 t; xor.b32 t, %0, %1; and.b32 %0, t, %2;}" : "+r"(b) : "r"(x), "r"(~seed));
    OPW(a0) OPA(b4) OPW(a1) OPA(b5) OPW(a2) OPA(b6) OPW(a3) OPA(b7)
#undef OPW
#undef OPA
  }
  a4 = b4; a5 = b5; a6 = b6; a7 = b7;
  KERNEL_EPILOGUE
}
// 8 independent DFMA chains
__global__ void k_dfma(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, d6 = a6, d7 = a7, m = 1.0 + 1e-9 * seed;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(d) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(m));
    OP(d0) OP(d1) OP(d2) OP(d3) OP(d4) OP(d5) OP(d6) OP(d7)
#undef OP
  }
  a0 = __double_as_longlong(d0); a1 = __double_as_longlong(d1); a2 = __double_as_longlong(d2);
  a3 = __double_as_longlong(d3); a4 = __double_as_longlong(d4); a5 = __double_as_longlong(d5);
  a6 = __double_as_longlong(d6); a7 = __double_as_longlong(d7);
  KERNEL_EPILOGUE
}
// mixed: 4 IMAD.WIDE chains + 4 DFMA chains
__global__ void k_mix_wide_dfma(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  double d4 = a4, d5 = a5, d6 = a6, d7 = a7, m = 1.0 + 1e-9 * seed;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OPW(a) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %2, %0;}" : "+l"(a) : "r"(x), "r"(seed));
#define OPD(d) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(m));
    OPW(a0) OPD(d4) OPW(a1) OPD(d5) OPW(a2) OPD(d6) OPW(a3) OPD(d7)
#undef OPW
#undef OPD
  }
  a4 = __double_as_longlong(d4); a5 = __double_as_longlong(d5);
  a6 = __double_as_longlong(d6); a7 = __double_as_longlong(d7);
  KERNEL_EPILOGUE
}
// 8 independent FFMA chains (reference point: the fp32 fma pipe)
__global__ void k_ffma(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  float f0 = a0, f1 = a1, f2 = a2, f3 = a3, f4 = a4, f5 = a5, f6 = a6, f7 = a7, m = 1.0f + 1e-6f * seed;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(f) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(m));
    OP(f0) OP(f1) OP(f2) OP(f3) OP(f4) OP(f5) OP(f6) OP(f7)
#undef OP
  }
  a0 = __float_as_uint(f0); a1 = __float_as_uint(f1); a2 = __float_as_uint(f2); a3 = __float_as_uint(f3);
  a4 = __float_as_uint(f4); a5 = __float_as_uint(f5); a6 = __float_as_uint(f6); a7 = __float_as_uint(f7);
  KERNEL_EPILOGUE
}


// 8 independent chains of IMAD.WIDE.U32 with RZ addend (product only)
__global__ void k_mul_wide(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(a) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mul.wide.u32 %0, l, %1;}" : "+l"(a) : "r"(seed));
    OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#undef OP
  }
  KERNEL_EPILOGUE
}
// accumulate form: a_k += y * c_k with y changing every iteration (not hoistable)
__global__ void k_mad_wide_acc2(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned y = x;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
    y = y * 1664525u + 1013904223u;
#define OP(a, c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a) : "r"(y), "n"(c));
    OP(a0, 17) OP(a1, 15) OP(a2, 41) OP(a3, 13) OP(a4, 28) OP(a5, 39) OP(a6, 18) OP(a7, 34)
#undef OP
  }
  KERNEL_EPILOGUE
}

// 8 independent chains of I2F.F64.U32 (+1 IADD to close the chain; only the conversion is counted)
__global__ void k_i2f_f64(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  unsigned b0 = a0, b1 = a1, b2 = a2, b3 = a3, b4 = a4, b5 = a5, b6 = a6, b7 = a7;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OP(b) asm volatile("{.reg .f64 d; .reg .u32 l, h; cvt.rn.f64.u32 d, %0; mov.b64 {l, h}, d; add.u32 %0, l, h;}" : "+r"(b));
    OP(b0) OP(b1) OP(b2) OP(b3) OP(b4) OP(b5) OP(b6) OP(b7)
#undef OP
  }
  a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
  KERNEL_EPILOGUE
}
// mixed: 6 DFMA chains + 2 I2F.F64.U32 chains (does the conversion share the FP64 pipe?)
__global__ void k_mix_dfma_i2f(unsigned long long* out, Clk* clk, unsigned seed) {
  KERNEL_PROLOGUE
  double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, m = 1.0 + 1e-9 * seed;
  unsigned b6 = a6, b7 = a7;
#pragma unroll 1
  for (int i = 0; i < ITER; i++) {
#define OPD(d) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(m));
#define OPC(b) asm volatile("{.reg .f64 d; .reg .u32 l, h; cvt.rn.f64.u32 d, %0; mov.b64 {l, h}, d; add.u32 %0, l, h;}" : "+r"(b));
    OPD(d0) OPD(d1) OPD(d2) OPC(b6) OPD(d3) OPD(d4) OPD(d5) OPC(b7)
#undef OPD
#undef OPC
  }
  a0 = __double_as_longlong(d0); a1 = __double_as_longlong(d1); a2 = __double_as_longlong(d2);
  a3 = __double_as_longlong(d3); a4 = __double_as_longlong(d4); a5 = __double_as_longlong(d5);
  a6 = b6; a7 = b7;
  KERNEL_EPILOGUE
}

typedef void (*kern_t)(unsigned long long*, Clk*, unsigned);

int run(const char* name, kern_t k, double instr_per_op, int sms, int threads, int blocks_per_sm) {
  const int blocks = sms * blocks_per_sm;
  unsigned long long* out;
  Clk* clk;
  CHECK(cudaMalloc(&out, sizeof(unsigned long long) * blocks * threads));
  CHECK(cudaMalloc(&clk, sizeof(Clk) * blocks));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, clk, 12345u);  // warm-up
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k<<<blocks, threads>>>(out, clk, 12345u);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<Clk> h(blocks);
  CHECK(cudaMemcpy(h.data(), clk, sizeof(Clk) * blocks, cudaMemcpyDeviceToHost));
  double sum = 0;
  long long mx = 0;
  for (auto& c : h) {
    sum += double(c.t1 - c.t0);
    if (c.t1 - c.t0 > mx) mx = c.t1 - c.t0;
  }
  const double mean_cyc = sum / blocks;
  const double lane_ops_per_sm = double(blocks_per_sm) * threads * ITER * 8.0 * instr_per_op;
  printf("{\"bench\": \"%s\", \"threads_per_sm\": %d, \"instr_lanes_per_clk_per_sm\": %.2f, "
         "\"cycles_per_iter\": %.2f, \"mean_cycles\": %.0f, \"max_cycles\": %lld, \"ms\": %.4f, \"implied_sm_mhz\": %.0f}\n",
         name, blocks_per_sm * threads, lane_ops_per_sm / mean_cyc, mean_cyc / ITER, mean_cyc, mx, ms,
         mx / (ms * 1e3));
  cudaFree(out);
  cudaFree(clk);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", p.name, p.multiProcessorCount,
         p.major, p.minor);
  const int sms = p.multiProcessorCount;
  for (int bps : {2, 8}) {
    run("imad_wide_u32_acc", k_imad_wide, 1, sms, 256, bps);
    run("imad_wide_u32_imm", k_imad_wide_imm, 1, sms, 256, bps);
    run("mul_wide_u32", k_mul_wide, 1, sms, 256, bps);
    run("mad_wide_acc2", k_mad_wide_acc2, 1, sms, 256, bps);
    run("imad_lo_u32", k_imad32, 1, sms, 256, bps);
    run("iadd3_x2", k_iadd3, 2, sms, 256, bps);
    run("add_u64 (2 instr)", k_add64, 2, sms, 256, bps);
    run("mix 4 imad_wide + 8 lop3", k_mix_wide_alu, 1.5, sms, 256, bps);
    run("dfma", k_dfma, 1, sms, 256, bps);
    run("mix imad_wide + dfma", k_mix_wide_dfma, 1, sms, 256, bps);
    run("ffma", k_ffma, 1, sms, 256, bps);
    run("i2f_f64_u32", k_i2f_f64, 1, sms, 256, bps);
    run("mix 6 dfma + 2 i2f_f64", k_mix_dfma_i2f, 1, sms, 256, bps);
  }
  return 0;
}
