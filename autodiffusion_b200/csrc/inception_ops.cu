// inception_ops.cu — the data-movement kernels of the on-device Inception-V3 pool_3 feature extractor.
//
// The reference computes FID features with a TensorFlow Inception graph on the host side of a uint8 round trip
// (evaluations/evaluator_v1.py:252-280, 665-679); the Stable-Diffusion search uses pytorch-fid's InceptionV3
// ("Stable Diffusion"/scripts/search_ea.py:95-127, 171-182). Here every convolution of that network is ONE patch
// gather (this file) + ONE tcgen05 GEMM (conv_igemm.cu as a 1-tap product over the patch rows); BatchNorm is folded
// into the packed weights / bias; ReLU is applied by the consumer when it loads (a conv writes its pre-activation).
// The network's odd spatial sizes (149, 147, 73, 71, 35, 17, 8), VALID / strided / 1x7 / 7x1 / 5x5 windows and
// four-way channel concatenations all live in the gather: a patch row is
//     [ (ky, kx) taps ] x [ source 0 channels | source 1 channels | ... ]   (+ zero tail up to k_pad)
// with out-of-image taps zero. The extractor is <0.3 % of a candidate's FLOPs: these kernels are written for
// generality and coalescing (a thread moves 8 channels = 16 bytes), not for the last GB/s.
#include "common.cuh"

namespace adb {

namespace {

constexpr int MAX_SRC = 4;

struct Sources {
  const __nv_bfloat16* ptr[MAX_SRC];
  int c[MAX_SRC];      // channels of each source (multiples of 8)
  int relu[MAX_SRC];   // apply max(x, 0) on load
  int nsrc;
  int ctot;
};

__device__ __forceinline__ uint4 relu8(uint4 v) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  return v;
}

// out [n*ho*wo, k_pad] bf16; one thread per (row, tap, 8-channel vector)
__global__ void __launch_bounds__(256) gather_patches_kernel(Sources s, __nv_bfloat16* __restrict__ out, int n, int h, int w,
                                                            int kh, int kw, int stride, int ph, int pw, int ho, int wo,
                                                            int k_pad) {
  const int vecs_per_tap = s.ctot / 8;
  const int vecs_per_row = k_pad / 8;
  const size_t total = (size_t)n * ho * wo * vecs_per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / vecs_per_row;
    const int kv = (int)(i - row * vecs_per_row);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int tap = kv / vecs_per_tap;
    if (tap < kh * kw) {
      int cv = (kv - tap * vecs_per_tap) * 8;  // channel offset inside the concatenation
      const int ky = tap / kw, kx = tap - ky * kw;
      const int x = (int)(row % wo);
      const size_t t = row / wo;
      const int y = (int)(t % ho);
      const int img = (int)(t / ho);
      const int iy = y * stride - ph + ky, ix = x * stride - pw + kx;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        int si = 0;
        while (si + 1 < s.nsrc && cv >= s.c[si]) {
          cv -= s.c[si];
          ++si;
        }
        v = __ldg(reinterpret_cast<const uint4*>(s.ptr[si] + (((size_t)img * h + iy) * w + ix) * s.c[si] + cv));
        if (s.relu[si]) v = relu8(v);
      }
    }
    *reinterpret_cast<uint4*>(out + row * k_pad + (size_t)kv * 8) = v;
  }
}

// 3x3 pooling over a (virtual) channel concatenation, ReLU on load. mode 0: max, 1: average over the full 3x3 window
// (count_include_pad = True, torchvision), 2: average over the in-image taps only (the FID Inception, pytorch-fid's
// FIDInceptionA/C/E_1). Max pooling ignores out-of-image taps. out [n, ho, wo, ctot] bf16.
__global__ void __launch_bounds__(256) pool3x3_kernel(Sources s, __nv_bfloat16* __restrict__ out, int n, int h, int w,
                                                     int stride, int pad, int ho, int wo, int mode) {
  const int vecs = s.ctot / 8;
  const size_t total = (size_t)n * ho * wo * vecs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / vecs;
    int cv = (int)(i - pix * vecs) * 8;
    const int cv_out = cv;
    const int x = (int)(pix % wo);
    const size_t t = pix / wo;
    const int y = (int)(t % ho);
    const int img = (int)(t / ho);
    int si = 0;
    while (si + 1 < s.nsrc && cv >= s.c[si]) {
      cv -= s.c[si];
      ++si;
    }
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = mode == 0 ? -INFINITY : 0.f;
    int cnt = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = y * stride - pad + ky, ix = x * stride - pad + kx;
        if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
        uint4 v = __ldg(reinterpret_cast<const uint4*>(s.ptr[si] + (((size_t)img * h + iy) * w + ix) * s.c[si] + cv));
        if (s.relu[si]) v = relu8(v);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = bf16_lo(u[k]), hi = bf16_hi(u[k]);
          if (mode == 0) {
            acc[2 * k] = fmaxf(acc[2 * k], lo);
            acc[2 * k + 1] = fmaxf(acc[2 * k + 1], hi);
          } else {
            acc[2 * k] += lo;
            acc[2 * k + 1] += hi;
          }
        }
        ++cnt;
      }
    if (mode != 0) {
      const float inv = 1.0f / (float)(mode == 1 ? 9 : cnt);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= inv;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + pix * s.ctot + cv_out) = o;
  }
}

// global average pool (ReLU on load) over a channel concatenation: out fp32 [n, ctot]; one thread per (image, channel pair)
__global__ void __launch_bounds__(256) global_avgpool_kernel(Sources s, float* __restrict__ out, int n, int hw) {
  const int pairs = s.ctot / 2;
  const size_t total = (size_t)n * pairs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int img = (int)(i / pairs);
    int c = (int)(i - (size_t)img * pairs) * 2;
    const int c_out = c;
    int si = 0;
    while (si + 1 < s.nsrc && c >= s.c[si]) {
      c -= s.c[si];
      ++si;
    }
    float a0 = 0.f, a1 = 0.f;
    const __nv_bfloat16* base = s.ptr[si] + (size_t)img * hw * s.c[si] + c;
    for (int p = 0; p < hw; ++p) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)p * s.c[si]));
      float lo = bf16_lo(u), hi = bf16_hi(u);
      if (s.relu[si]) {
        lo = fmaxf(lo, 0.f);
        hi = fmaxf(hi, 0.f);
      }
      a0 += lo;
      a1 += hi;
    }
    out[(size_t)img * s.ctot + c_out] = a0 / (float)hw;
    out[(size_t)img * s.ctot + c_out + 1] = a1 / (float)hw;
  }
}

// uint8 NHWC [n, h, w, 3] -> bf16 NHWC [n, oh, ow, 8] (channels 3..7 zero: the first conv's gather wants 16-byte
// pixels): bilinear, half-pixel centres, no antialiasing, source index clamped - F.interpolate(mode="bilinear",
// align_corners=False) as pytorch-fid resizes - then x / 127.5 - 1.
__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                                int n, int h, int w, int oh, int ow) {
  const size_t total = (size_t)n * oh * ow;
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % ow);
    const size_t t = i / ow;
    const int y = (int)(t % oh);
    const int img = (int)(t / oh);
    const float fy = fmaxf(((float)y + 0.5f) * sy - 0.5f, 0.f);
    const float fx = fmaxf(((float)x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const uint8_t* b = in + (size_t)img * h * w * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = b[((size_t)y0 * w + x0) * 3 + c], p01 = b[((size_t)y0 * w + x1) * 3 + c];
      const float p10 = b[((size_t)y1 * w + x0) * 3 + c], p11 = b[((size_t)y1 * w + x1) * 3 + c];
      const float top = p00 + (p01 - p00) * lx, bot = p10 + (p11 - p10) * lx;
      v[c] = (top + (bot - top) * ly) / 127.5f - 1.0f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], 0.f);
    o.z = 0u;
    o.w = 0u;
    *reinterpret_cast<uint4*>(out + i * 8) = o;
  }
}

int fill_sources(Sources& s, const void* const* ptrs, const int* chans, const int* relu, int nsrc, const char* what) {
  ADB_REQUIRE(nsrc >= 1 && nsrc <= MAX_SRC, "%s: 1..%d sources", what, MAX_SRC);
  s.nsrc = nsrc;
  s.ctot = 0;
  for (int i = 0; i < MAX_SRC; ++i) {
    s.ptr[i] = nullptr;
    s.c[i] = 0;
    s.relu[i] = 0;
  }
  for (int i = 0; i < nsrc; ++i) {
    ADB_REQUIRE(ptrs[i] != nullptr && chans[i] > 0 && chans[i] % 8 == 0, "%s: source %d needs a pointer and channels %% 8 == 0", what, i);
    s.ptr[i] = reinterpret_cast<const __nv_bfloat16*>(ptrs[i]);
    s.c[i] = chans[i];
    s.relu[i] = relu[i] ? 1 : 0;
    s.ctot += chans[i];
  }
  return ADB_OK;
}

unsigned grid_for(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

int gather_patches_submit(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out,
                          int n, int h, int w, int kh, int kw, int stride, int ph, int pw, int k_pad, cudaStream_t stream) {
  Sources s;
  int r = fill_sources(s, ptrs, chans, relu, nsrc, "gather_patches");
  if (r != ADB_OK) return r;
  ADB_REQUIRE(out && n > 0 && h > 0 && w > 0 && kh > 0 && kw > 0 && stride > 0 && ph >= 0 && pw >= 0, "gather_patches: bad arguments");
  const int ho = (h + 2 * ph - kh) / stride + 1, wo = (w + 2 * pw - kw) / stride + 1;
  ADB_REQUIRE(ho > 0 && wo > 0 && k_pad % 8 == 0 && k_pad >= kh * kw * s.ctot, "gather_patches: k_pad must be a multiple of 8 and >= kh*kw*channels");
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const double bytes = 2.0 * (double)n * ho * wo * k_pad * 2.0;  // patches written + (at most) as many bytes gathered
  return submit(plan, stream, "gather_patches", 0.0, bytes, [=](cudaStream_t st) -> int {
    const size_t total = (size_t)n * ho * wo * (k_pad / 8);
    gather_patches_kernel<<<grid_for(total), 256, 0, st>>>(s, o, n, h, w, kh, kw, stride, ph, pw, ho, wo, k_pad);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pool3x3_submit(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out, int n, int h,
                   int w, int stride, int pad, int mode, cudaStream_t stream) {
  Sources s;
  int r = fill_sources(s, ptrs, chans, relu, nsrc, "pool3x3");
  if (r != ADB_OK) return r;
  ADB_REQUIRE(out && n > 0 && h >= 3 - 2 * pad && w >= 3 - 2 * pad && (stride == 1 || stride == 2) && (pad == 0 || pad == 1) &&
                  mode >= 0 && mode <= 2, "pool3x3: bad arguments");
  const int ho = (h + 2 * pad - 3) / stride + 1, wo = (w + 2 * pad - 3) / stride + 1;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const double bytes = 2.0 * ((double)n * h * w + (double)n * ho * wo) * s.ctot;
  return submit(plan, stream, "pool3x3", 0.0, bytes, [=](cudaStream_t st) -> int {
    const size_t total = (size_t)n * ho * wo * (s.ctot / 8);
    pool3x3_kernel<<<grid_for(total), 256, 0, st>>>(s, o, n, h, w, stride, pad, ho, wo, mode);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int global_avgpool_submit(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, float* out, int n,
                          int hw, cudaStream_t stream) {
  Sources s;
  int r = fill_sources(s, ptrs, chans, relu, nsrc, "global_avgpool");
  if (r != ADB_OK) return r;
  ADB_REQUIRE(out && n > 0 && hw > 0, "global_avgpool: bad arguments");
  return submit(plan, stream, "global_avgpool", 0.0, 2.0 * (double)n * hw * s.ctot, [=](cudaStream_t st) -> int {
    global_avgpool_kernel<<<grid_for((size_t)n * s.ctot / 2), 256, 0, st>>>(s, out, n, hw);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int resize_bilinear_u8_submit(adb_plan* plan, const uint8_t* in, void* out, int n, int h, int w, int oh, int ow,
                              cudaStream_t stream) {
  ADB_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "resize_bilinear_u8: bad arguments");
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  return submit(plan, stream, "resize_bilinear_u8", 0.0, (double)n * (3.0 * h * w + 16.0 * oh * ow), [=](cudaStream_t st) -> int {
    resize_bilinear_u8_kernel<<<grid_for((size_t)n * oh * ow), 256, 0, st>>>(in, o, n, h, w, oh, ow);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
