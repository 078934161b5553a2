// groupnorm.cu — GroupNorm(32 groups) over bf16 NHWC activations, fused with the FiLM
// scale-shift, SiLU and the 2x resample that follow it in the reference ResBlock.
//
// Reference: GroupNorm32.forward (guided_diffusion/nn.py:17-19: fp32 statistics over
// (C/32 channels x H x W) per sample, eps 1e-5, affine), nn.SiLU, `out_norm(h)*(1+scale)+shift`
// (dynamic_unet.py:262-265), h_upd = avg-pool / nearest-upsample (dynamic_unet.py:253-254),
// and the concat input `th.cat([h, hs.pop()], 1)` (dynamic_unet.py:699) read from two sources.
//
// Memory-bound: pass 1 reads the tensor once and reduces (sum, sum of squares) per
// (sample, group) — fp32 in registers over <= a few hundred values, fp64 across threads and
// CTAs; pass 2 re-reads it (L2-resident for the per-sample slabs that fit) and writes the
// normalised bf16 result: 2 B read (+2 B L2 re-read) + 2 B written per element.
#include <stdlib.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_GROUPS = 32;

struct GnParams {
  const __nv_bfloat16* src0;
  const __nv_bfloat16* src1;
  int c0, c1, C;
  int n, H, W;
  const float* gamma;
  const float* beta;
  float eps;
  const float* scale_shift;
  int ss_stride;
  int silu;
  int resample;
  __nv_bfloat16* out;
  double* stats;  // [n][32][2]
  int stats_ready;
  int reverse;
  int splits;
};

__device__ __forceinline__ uint4 load_vec(const GnParams& p, size_t pix, int v) {
  const int ch = v * 8;
  if (ch < p.c0) return __ldg(reinterpret_cast<const uint4*>(p.src0 + pix * p.c0 + ch));
  return __ldg(reinterpret_cast<const uint4*>(p.src1 + pix * p.c1 + (ch - p.c0)));
}

__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x);
  f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
  f[4] = bf16_lo(r.z); f[5] = bf16_hi(r.z);
  f[6] = bf16_lo(r.w); f[7] = bf16_hi(r.w);
}

// ---- pass 1: statistics ----
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const GnParams p) {
  __shared__ double s_acc[GN_GROUPS][2];
  const int n = blockIdx.y;
  const int V = p.C / 8;
  const int P = p.H * p.W;
  const int cpg = p.C / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS * 2) (&s_acc[0][0])[threadIdx.x] = 0.0;
  __syncthreads();

  const int per = (P + p.splits - 1) / p.splits;
  const int p_begin = blockIdx.x * per;
  const int p_end = min(P, p_begin + per);

  // thread -> (vector slot, pixel lane). V may exceed the block: loop over slots.
  for (int v = threadIdx.x % min(V, GN_THREADS); v < V; v += GN_THREADS) {
    const int slots = min(V, GN_THREADS);
    const int lanes = max(1, GN_THREADS / slots);
    const int pl = threadIdx.x / slots;
    if (pl >= lanes) break;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    int pix = p_begin + pl;
    // 4 independent 16-byte loads in flight per thread
    for (; pix + 3 * lanes < p_end; pix += 4 * lanes) {
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = load_vec(p, (size_t)n * P + pix + u * lanes, v);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(r[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] = fmaf(f[i], f[i], q[i]);
        }
      }
    }
    for (; pix < p_end; pix += lanes) {
      float f[8];
      unpack8(load_vec(p, (size_t)n * P + pix, v), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    }
    // fold the 8 channels into their groups (consecutive channels mostly share a group)
    int g_cur = (v * 8) / cpg;
    double ds = 0.0, dq = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (v * 8 + i) / cpg;
      if (g != g_cur) {
        atomicAdd(&s_acc[g_cur][0], ds);
        atomicAdd(&s_acc[g_cur][1], dq);
        ds = dq = 0.0;
        g_cur = g;
      }
      ds += (double)s[i];
      dq += (double)q[i];
    }
    atomicAdd(&s_acc[g_cur][0], ds);
    atomicAdd(&s_acc[g_cur][1], dq);
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS * 2) {
    atomicAdd(p.stats + (size_t)n * GN_GROUPS * 2 + threadIdx.x, (&s_acc[0][0])[threadIdx.x]);
  }
}

// ---- pass 2: normalise (+FiLM) (+SiLU) (+resample) ----
__device__ __forceinline__ void affine_act8(const float* f, const float* a, const float* b, int silu,
                                            float* o) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float y = fmaf(f[i], a[i], b[i]);
    o[i] = silu ? silu_f(y) : y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* o) {
  uint4 r;
  r.x = pack_bf16x2(o[0], o[1]);
  r.y = pack_bf16x2(o[2], o[3]);
  r.z = pack_bf16x2(o[4], o[5]);
  r.w = pack_bf16x2(o[6], o[7]);
  return r;
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const GnParams p) {
  extern __shared__ float s_ab[];  // [2][C]
  float* s_a = s_ab;
  float* s_b = s_ab + p.C;
  // reverse traversal (last image / last pixels first): the input was just written front-to-back by
  // its producer, so its tail is what is still resident in the 126 MB L2
  const int n = p.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int V = p.C / 8;
  const int P = p.H * p.W;
  const int cpg = p.C / GN_GROUPS;
  const double cnt = (double)cpg * (double)P;
  for (int c = threadIdx.x; c < p.C; c += GN_THREADS) {
    const int g = c / cpg;
    const double sum = p.stats[((size_t)n * GN_GROUPS + g) * 2 + 0];
    const double sq = p.stats[((size_t)n * GN_GROUPS + g) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    float a = rstd * p.gamma[c];
    float b = p.beta[c] - (float)mean * a;
    if (p.scale_shift != nullptr) {
      const float sc = 1.0f + p.scale_shift[(size_t)n * p.ss_stride + c];
      const float sh = p.scale_shift[(size_t)n * p.ss_stride + p.C + c];
      a = a * sc;
      b = fmaf(b, sc, sh);
    }
    s_a[c] = a;
    s_b[c] = b;
  }
  __syncthreads();

  const int slots = min(V, GN_THREADS);
  const int lanes = max(1, GN_THREADS / slots);
  const int pl = threadIdx.x / slots;
  if (pl >= lanes) return;

  // iteration space: output pixels for AVGPOOL2, input pixels otherwise
  const int Wo = (p.resample == ADB_RESAMPLE_AVGPOOL2) ? p.W / 2 : p.W;
  const int Ho = (p.resample == ADB_RESAMPLE_AVGPOOL2) ? p.H / 2 : p.H;
  const int PI = Ho * Wo;
  const int per = (PI + p.splits - 1) / p.splits;
  const int p_begin = (p.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * per;
  const int p_end = min(PI, p_begin + per);

  for (int v = threadIdx.x % slots; v < V; v += GN_THREADS) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = s_a[v * 8 + i];
      b[i] = s_b[v * 8 + i];
    }
    int pix = p_begin + pl;
    if (p.resample == ADB_RESAMPLE_NONE) {
      for (; pix + 3 * lanes < p_end; pix += 4 * lanes) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = load_vec(p, (size_t)n * P + pix + u * lanes, v);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8], o[8];
          unpack8(r[u], f);
          affine_act8(f, a, b, p.silu, o);
          *reinterpret_cast<uint4*>(p.out + ((size_t)n * P + pix + u * lanes) * p.C + v * 8) = pack8(o);
        }
      }
    }
    for (; pix < p_end; pix += lanes) {
      float f[8], o[8];
      if (p.resample == ADB_RESAMPLE_NONE) {
        unpack8(load_vec(p, (size_t)n * P + pix, v), f);
        affine_act8(f, a, b, p.silu, o);
        *reinterpret_cast<uint4*>(p.out + ((size_t)n * P + pix) * p.C + v * 8) = pack8(o);
      } else if (p.resample == ADB_RESAMPLE_AVGPOOL2) {
        const int yo = pix / Wo, xo = pix - yo * Wo;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int yi = 2 * yo + (s >> 1), xi = 2 * xo + (s & 1);
          unpack8(load_vec(p, (size_t)n * P + (size_t)yi * p.W + xi, v), f);
          affine_act8(f, a, b, p.silu, o);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += o[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= 0.25f;
        *reinterpret_cast<uint4*>(p.out + ((size_t)n * PI + pix) * p.C + v * 8) = pack8(acc);
      } else {  // nearest 2x: each input pixel lands on a 2x2 output patch
        const int yi = pix / p.W, xi = pix - yi * p.W;
        unpack8(load_vec(p, (size_t)n * P + pix, v), f);
        affine_act8(f, a, b, p.silu, o);
        const uint4 r = pack8(o);
        const int W2 = 2 * p.W;
        __nv_bfloat16* base = p.out + ((size_t)n * 4 * P) * p.C + v * 8;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const size_t opix = (size_t)(2 * yi + (s >> 1)) * W2 + (2 * xi + (s & 1));
          *reinterpret_cast<uint4*>(base + opix * p.C) = r;
        }
      }
    }
  }
}

// ---- plain 2x resample of a bf16 NHWC tensor (x_upd of a skipped up/down ResBlock) ----
__global__ void __launch_bounds__(256) resample2x_kernel(const __nv_bfloat16* __restrict__ src,
                                                        __nv_bfloat16* __restrict__ dst, int n, int H,
                                                        int W, int C, int mode) {
  const int V = C / 8;
  const int Ho = (mode == ADB_RESAMPLE_AVGPOOL2) ? H / 2 : H * 2;
  const int Wo = (mode == ADB_RESAMPLE_AVGPOOL2) ? W / 2 : W * 2;
  const size_t total = (size_t)n * Ho * Wo * V;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    size_t r = i / V;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int img = (int)(r / Ho);
    uint4 o;
    if (mode == ADB_RESAMPLE_AVGPOOL2) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const size_t pix = ((size_t)img * H + 2 * yo + (s >> 1)) * W + 2 * xo + (s & 1);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(src + pix * C + v * 8)), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
      o = pack8(acc);
    } else {
      const size_t pix = ((size_t)img * H + (yo >> 1)) * W + (xo >> 1);
      o = __ldg(reinterpret_cast<const uint4*>(src + pix * C + v * 8));
    }
    *reinterpret_cast<uint4*>(dst + (((size_t)img * Ho + yo) * Wo + xo) * C + v * 8) = o;
  }
}

}  // namespace

int groupnorm_submit(adb_plan* plan, const adb_gn_desc* d, cudaStream_t stream) {
  ADB_REQUIRE(d != nullptr, "groupnorm: null descriptor");
  const int C = d->c0 + d->c1;
  ADB_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "groupnorm: bad geometry");
  ADB_REQUIRE(d->src0 && d->c0 > 0 && d->c0 % 8 == 0, "groupnorm: c0 must be a positive multiple of 8");
  ADB_REQUIRE(d->c1 == 0 || (d->src1 && d->c1 % 8 == 0), "groupnorm: c1 must be a multiple of 8 with a source");
  ADB_REQUIRE(C % GN_GROUPS == 0, "groupnorm: channels (%d) must divide into 32 groups", C);
  ADB_REQUIRE(d->gamma && d->beta && d->out && d->stats, "groupnorm: null gamma/beta/out/stats");
  ADB_REQUIRE(d->resample >= 0 && d->resample <= 2, "groupnorm: bad resample mode");
  if (d->resample == ADB_RESAMPLE_AVGPOOL2)
    ADB_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, "groupnorm: avgpool2 needs even h,w");
  ADB_REQUIRE(2 * C * sizeof(float) <= 48 * 1024, "groupnorm: too many channels (%d)", C);

  GnParams p;
  p.src0 = reinterpret_cast<const __nv_bfloat16*>(d->src0);
  p.src1 = reinterpret_cast<const __nv_bfloat16*>(d->src1);
  p.c0 = d->c0;
  p.c1 = d->c1;
  p.C = C;
  p.n = d->n;
  p.H = d->h;
  p.W = d->w;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.eps = d->eps;
  p.scale_shift = d->scale_shift;
  p.ss_stride = d->ss_stride;
  p.silu = d->silu;
  p.resample = d->resample;
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.stats = d->stats;
  p.stats_ready = d->stats_ready;
  {
    static int rev = -1;
    if (rev < 0) {
      const char* e = getenv("ADB_GN_REVERSE");
      rev = (e && e[0] == '1') ? 1 : 0;
    }
    p.reverse = rev;
  }
  const int P = d->h * d->w;
  int splits = (8 * num_sms() + d->n - 1) / d->n;
  const int max_splits = (P / 64) > 1 ? (P / 64) : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;

  const double in_elems = (double)d->n * P * C;
  const double out_elems = d->resample == ADB_RESAMPLE_AVGPOOL2 ? in_elems / 4 : (d->resample == ADB_RESAMPLE_NEAREST2 ? in_elems * 4 : in_elems);
  return submit(plan, stream, p.stats_ready ? "groupnorm_apply" : "groupnorm", 0.0, 2.0 * (in_elems + out_elems), [p](cudaStream_t s) -> int {
    dim3 grid(p.splits, p.n);
    int launches = 1;
    if (!p.stats_ready) {
      ADB_CUDA(cudaMemsetAsync(p.stats, 0, (size_t)p.n * GN_GROUPS * 2 * sizeof(double), s));
      gn_stats_kernel<<<grid, GN_THREADS, 0, s>>>(p);
      ADB_CUDA(cudaGetLastError());
      launches = 3;
    }
    gn_apply_kernel<<<grid, GN_THREADS, 2 * p.C * sizeof(float), s>>>(p);
    ADB_CUDA(cudaGetLastError());
    return launches;
  });
}

int resample2x_submit(adb_plan* plan, const void* src, void* dst, int n, int h, int w, int c, int mode,
                      cudaStream_t stream) {
  ADB_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "resample2x: bad arguments");
  ADB_REQUIRE(mode == ADB_RESAMPLE_AVGPOOL2 || mode == ADB_RESAMPLE_NEAREST2, "resample2x: bad mode");
  if (mode == ADB_RESAMPLE_AVGPOOL2) ADB_REQUIRE(h % 2 == 0 && w % 2 == 0, "resample2x: avgpool2 needs even h,w");
  return submit(plan, stream, "resample2x", 0.0, 0.0, [=](cudaStream_t s) -> int {
    const size_t out_pix = (mode == ADB_RESAMPLE_AVGPOOL2) ? (size_t)n * (h / 2) * (w / 2) : (size_t)n * h * w * 4;
    const size_t total = out_pix * (c / 8);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    resample2x_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                       reinterpret_cast<__nv_bfloat16*>(dst), n, h, w, c, mode);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
