// Microbenchmark: do tcgen05.ld (TMEM -> registers) and MUFU.EX2 overlap, or do they queue behind each other?
// One CTA per SM; each of `warps` warps (1 or 2 per SM sub-partition) repeats, per iteration, the per-thread work of one
// 128 x 64 fp32 attention score tile: mode 0 = two tcgen05.ld.32x32b.x32 (8 KB per warp), mode 1 = 64 ex2.approx,
// mode 2 = both (load, wait, then the exponentials of the loaded values - the softmax loop's shape),
// mode 3 = both with the next load issued before the exponentials (software pipelined).
// Prints cycles per iteration per warp-slot: if mode 2 ~ mode 0 + mode 1 the two share a queue.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/tmem_mufu_mix scripts/microbench/tmem_mufu_mix.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE>
__global__ void __launch_bounds__(256) mix(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = *(volatile uint32_t*)&slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t a[32], b[32];
  float acc = 0.f;
  for (int i = 0; i < 32; ++i) a[i] = b[i] = 0x3c000000u + threadIdx.x + i;
  if (MODE == 3) {
    ld32(base, a);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      ld32(base, a);
      ld32(base + 32, b);
      wait_ld();
      acc += __uint_as_float(a[it & 31] ^ b[(it + 7) & 31]) * 1e-30f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(a[i]) * 0.1f - acc * 1e-9f) + ex2(__uint_as_float(b[i]) * 0.1f - acc * 1e-9f);
    } else if (MODE == 2) {
      ld32(base, a);
      ld32(base + 32, b);
      wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-20f) + ex2(__uint_as_float(b[i]) * 1e-20f);
    } else {
      wait_ld();
      ld32(base + 32, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(a[i]) * 1e-20f);
      wait_ld();
      ld32(base, a);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(b[i]) * 1e-20f);
    }
  }
  wait_ld();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc) ^ a[3] ^ b[5];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*(volatile uint32_t*)&slot), "r"(128u) : "memory");
}

template <int MODE>
void run(const char* name, uint32_t* out, long long* cyc, int threads) {
  const int iters = 2000;
  mix<MODE><<<148, threads>>>(out, cyc, 10);
  mix<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-34s %d warps/SM (%d per sub-partition): %.0f cycles per tile-iteration of each warp\n", name, threads / 32,
         threads / 128, s / 148 / iters);
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int threads = 128; threads <= 256; threads *= 2) {
    run<0>("tcgen05.ld 2 x x32 (8 KB/warp)", out, cyc, threads);
    run<1>("64 x MUFU.EX2", out, cyc, threads);
    run<2>("ld, wait, then 64 x EX2", out, cyc, threads);
    run<3>("pipelined: next ld under the EX2s", out, cyc, threads);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
