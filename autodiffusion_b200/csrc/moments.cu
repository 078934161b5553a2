// moments.cu — streaming first/second moments of Inception features for the search's FID.
//
// Reference: Evaluator.compute_statistics (evaluations/evaluator_v1.py:218-221) computes
// mu = mean(F, 0) and sigma = np.cov(F, rowvar=False) on the host from all gathered samples.
// Here every rank accumulates n, sum_x and sum_xx = F^T F in fp64 on the device; one NCCL
// all-reduce of (sum_x, sum_xx) per candidate replaces the reference's image all_gather
// (search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:425-430), and the host
// finishes mu / sigma in fp64.
//
// fp64 FMA on CUDA cores (features are fp32, d = 2048, n ~ 1e3 per candidate: 8.4 GFLOP).
// Only tiles on or above the diagonal are computed; off-diagonal tiles are mirrored on store.
#include "common.cuh"

namespace adb {

namespace {

constexpr int MT = 64;  // tile edge
constexpr int MK = 16;  // samples per smem step

__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ f, int n, int d,
                                                     double* __restrict__ sum_x,
                                                     double* __restrict__ sum_xx) {
  const int tiles = (d + MT - 1) / MT;
  // linear block index -> (ti <= tj) upper-triangular tile
  int ti = 0, rem = blockIdx.x;
  while (rem >= tiles - ti) {
    rem -= tiles - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const int d0 = ti * MT, e0 = tj * MT;

  __shared__ double As[MK][MT + 2];
  __shared__ double Bs[MK][MT + 2];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double colsum = 0.0;  // threads 0..63 of diagonal tiles accumulate sum_x

  for (int i0 = 0; i0 < n; i0 += MK) {
    for (int l = threadIdx.x; l < MK * MT; l += 256) {
      const int r = l / MT, c = l - r * MT;
      const int row = i0 + r;
      As[r][c] = (row < n && d0 + c < d) ? (double)f[(size_t)row * d + d0 + c] : 0.0;
      Bs[r][c] = (row < n && e0 + c < d) ? (double)f[(size_t)row * d + e0 + c] : 0.0;
    }
    __syncthreads();
    if (ti == tj && threadIdx.x < MT) {
#pragma unroll
      for (int k = 0; k < MK; ++k) colsum += As[k][threadIdx.x];
    }
#pragma unroll
    for (int k = 0; k < MK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int dd = d0 + ty * 4 + i;
    if (dd >= d) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ee = e0 + tx * 4 + j;
      if (ee >= d) continue;
      sum_xx[(size_t)dd * d + ee] += acc[i][j];
      if (ti != tj) sum_xx[(size_t)ee * d + dd] += acc[i][j];  // mirror
    }
  }
  if (ti == tj && threadIdx.x < MT && d0 + threadIdx.x < d) sum_x[d0 + threadIdx.x] += colsum;
}

}  // namespace

int moments_submit(adb_plan* plan, const float* feats, int n, int d, double* sum_x, double* sum_xx,
                   cudaStream_t stream) {
  ADB_REQUIRE(feats && sum_x && sum_xx && n > 0 && d > 0, "moments_accumulate: bad arguments");
  return submit(plan, stream, "moments", 2.0 * n * (double)d * d, 0.0, [=](cudaStream_t s) -> int {
    const int tiles = (d + MT - 1) / MT;
    const int blocks = tiles * (tiles + 1) / 2;
    moments_kernel<<<blocks, 256, 0, s>>>(feats, n, d, sum_x, sum_xx);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
