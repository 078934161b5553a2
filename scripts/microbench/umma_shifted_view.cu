// Feasibility check for halo reuse of the conv A operand: can tcgen05.mma read a K-major, 128-byte-swizzled
// A tile that starts at an arbitrary 128-byte row of a larger shared-memory buffer written by ONE TMA box?
// One CTA loads a [272 rows][64 bf16] buffer (two TMA boxes, SWIZZLE_128B) and a [64][64] weight tile, then for every shift s
// computes D_s = X[s : s + 128, :] . W^T (M = 128, N = 64, K = 64) with the A descriptor's start address advanced by
// s * 128 bytes, in two modes: matrix-descriptor base offset 0, or base offset = (address >> 7) & 7 (the swizzle phase of
// the first row). The host checks every D_s exactly (small-integer data).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/umma_shifted_view scripts/microbench/umma_shifted_view.cu
#include <cstdlib>
#include <vector>

#include "../../autodiffusion_b200/csrc/common.cuh"

using namespace adb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int XROWS = 272, NSHIFT = 12;
__constant__ int c_shifts[NSHIFT];

struct P {
  CUtensorMap tmX;  // box {64, 136}
  CUtensorMap tmW;  // box {64, 64}
  float* out;       // [mode][NSHIFT][128][64]
};

__global__ void __launch_bounds__(128, 1) shifted_view(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_smem = base, w_smem = base + 35 * 1024;
  const uint32_t bar_ld = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot_s), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_ld, XROWS * 128 + 64 * 128);
    tma_load_2d(x_smem, &p.tmX, bar_ld, 0, 0);
    tma_load_2d(x_smem + 136 * 128, &p.tmX, bar_ld, 0, 136);
    tma_load_2d(w_smem, &p.tmW, bar_ld, 0, 0);
  }
  mbar_wait(bar_ld, 0);
  tc_fence_after();
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int si = 0; si < NSHIFT; ++si) {
      if (threadIdx.x == 0) {
        const uint32_t a_addr = x_smem + (uint32_t)c_shifts[si] * 128u;
        uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
        if (mode == 1) a_desc |= (uint64_t)((a_addr >> 7) & 7u) << 49;  // matrix-descriptor base offset
        const uint64_t b_desc = umma_desc_kmajor_sw128(w_smem);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base, a_desc + 2u * kk, b_desc + 2u * kk, idesc, kk != 0);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1u;
      tc_fence_after();
      uint32_t v[32];
      float* orow = p.out + (((size_t)mode * NSHIFT + si) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 64; c += 32) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) orow[c + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map(EncodeFn fn, CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
  cuuint64_t gdim[2] = {64, rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  const int shifts[NSHIFT] = {0, 8, 16, 1, 2, 3, 7, 9, 65, 66, 67, 133};
  CK(cudaMemcpyToSymbol(c_shifts, shifts, sizeof(shifts)));
  std::vector<__nv_bfloat16> hx(XROWS * 64), hw(64 * 64);
  std::vector<float> fx(XROWS * 64), fw(64 * 64);
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 5 - 2); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 5 - 2); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dout;
  const size_t nout = (size_t)2 * NSHIFT * 128 * 64;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dout, nout * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, nout * 4));
  P p;
  make_map((EncodeFn)fnp, &p.tmX, dx, XROWS, 136);
  make_map((EncodeFn)fnp, &p.tmW, dw, 64, 64);
  p.out = dout;
  CK(cudaFuncSetAttribute(shifted_view, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  shifted_view<<<1, 128, 45 * 1024>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(nout);
  CK(cudaMemcpy(ho.data(), dout, nout * 4, cudaMemcpyDeviceToHost));
  for (int mode = 0; mode < 2; ++mode)
    for (int si = 0; si < NSHIFT; ++si) {
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += fx[(size_t)(shifts[si] + r) * 64 + k] * fw[(size_t)n * 64 + k];
          if (ho[(((size_t)mode * NSHIFT + si) * 128 + r) * 64 + n] != ref) ++bad;
        }
      printf("base offset %s  shift %3d rows (start %% 1024 = %4d B): %s (%d of 8192 elements differ)\n", mode ? "(addr>>7)&7" : "0          ",
             shifts[si], (shifts[si] * 128) % 1024, bad ? "WRONG" : "exact", bad);
    }
  return 0;
}
