// Microbenchmark: MUFU exp2 throughput per SM for f32 / f16x2 / bf16x2 operands (sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/ex2_rate scripts/microbench/ex2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i);
    u[i] = 0xBC00BC00u + i;  // small negative halves / bf16s
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, int per_instr) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 4, 512>>>(out, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 512>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double exps = 148.0 * 4 * 512 * iters * 8.0 * per_instr;
  printf("%-10s %8.3f ms  %.3e exp/s  (%.2f exp/clk/SM at 1.9 GHz)\n", name, ms, exps / (ms * 1e-3), exps / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
  return exps / (ms * 1e-3);
}

__global__ void acc(const float* x, float* y32, float* ybf, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i], r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  y32[i] = r;
  uint32_t p, q;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v), "f"(v));
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(q) : "r"(p));
  ybf[i] = __uint_as_float(q << 16);
}

int main() {
  run<0>("f32", 1);
  run<1>("f16x2", 2);
  run<2>("bf16x2", 2);
  const int n = 4096;
  float *x, *a, *b;
  cudaMallocManaged(&x, n * 4);
  cudaMallocManaged(&a, n * 4);
  cudaMallocManaged(&b, n * 4);
  for (int i = 0; i < n; ++i) x[i] = 8.0f - 24.0f * i / n;
  acc<<<(n + 255) / 256, 256>>>(x, a, b, n);
  cudaDeviceSynchronize();
  double worst = 0, worst_w = 0;
  for (int i = 0; i < n; ++i) {
    double rel = fabs(b[i] - a[i]) / a[i];
    if (rel > worst) worst = rel;
    double w = rel * (a[i] < 1 ? a[i] : 1.0);  // error weighted by the probability's size relative to the row max
    if (x[i] <= 0 && w > worst_w) worst_w = w;
  }
  printf("bf16x2 ex2 vs f32 ex2 on [-16, 8]: worst relative %.4f, worst p-weighted (x <= 0) %.5f\n", worst, worst_w);
  return 0;
}
