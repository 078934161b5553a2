// Microbenchmark: the attention softmax inner loop's instruction mix (64 x FFMA -> MUFU.EX2, 32 x F2FP pack per row
// tile) without TMEM / barriers, at 1, 2 and 4 warps per SM sub-partition: can this mix saturate the MUFU pipe?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/softmax_mix scripts/microbench/softmax_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) mix(uint32_t* out, int iters, float sc, float m) {
  float s[64];
  for (int i = 0; i < 64; ++i) s[i] = -0.01f * ((threadIdx.x * 7 + i) & 255);
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float p0, p1;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(s[2 * i], sc, -m)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(s[2 * i + 1], sc, -m)));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(p1), "f"(p0));
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    m += 1e-6f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s[i] += __uint_as_float(acc & 1u);  // keep the loop body live, values ~unchanged
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 512 * 4);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    const int threads = 128 * warps_per_smsp, iters = 4000;
    mix<<<148, threads>>>(out, 10, 0.18f, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix<<<148, threads>>>(out, iters, 0.18f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double exps = 148.0 * threads * iters * 64.0;
    printf("%d warp(s)/SMSP: %.3f ms, %.3e exp/s, %.2f exp/clk/SM at the max clock %d MHz\n", warps_per_smsp, ms, exps / (ms * 1e-3),
           exps / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  }
  return 0;
}
