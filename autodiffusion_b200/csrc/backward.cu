// backward.cu — the memory-bound pieces of the classifier-guidance input gradient
// (cond_fn of search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-390:
//  th.autograd.grad(log_softmax(classifier(x_t, t))[y].sum(), x_t) * classifier_scale):
//
//   * GroupNorm32 (+FiLM) (+SiLU) (+2x average pool) backward            (nn.py:17-19, unet.py:236-258)
//   * AttentionPool2d forward / backward for the single token it returns  (unet.py:22-51)
//   * d log_softmax(logits)[y] / d logits * scale
//
// The convolutions' and projections' data gradients are ordinary implicit GEMMs with transposed
// (and spatially flipped) weights and run on conv_igemm.cu; attention is in attention_bwd.cu.
#include <stdlib.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int GB_THREADS = 256;
constexpr int GN_GROUPS = 32;

__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x);
  f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
  f[4] = bf16_lo(r.z); f[5] = bf16_hi(r.z);
  f[6] = bf16_lo(r.w); f[7] = bf16_hi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float* o) {
  uint4 r;
  r.x = pack_bf16x2(o[0], o[1]);
  r.y = pack_bf16x2(o[2], o[3]);
  r.z = pack_bf16x2(o[4], o[5]);
  r.w = pack_bf16x2(o[6], o[7]);
  return r;
}

struct GnBwdParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dout;
  const __nv_bfloat16* add;
  __nv_bfloat16* dx;
  const double* stats;
  double* bstats;
  const float* gamma;
  const float* beta;
  const float* scale_shift;
  int ss_stride;
  float eps;
  int n, H, W, C;
  int silu, resample, add_mode;
  int splits;
};

// Forward (groupnorm.cu): xh = (x - mean) rstd; z = xh A + B with A = gamma (1+scale), B = beta (1+scale) + shift;
// y = silu(z) (or z), then optionally 2x2 average pooled. Given dy:
//   dz = dy silu'(z);  dxh = dz A;  dx = rstd (dxh - mean_g(dxh) - xh mean_g(dxh xh))
// pass 1 (APPLY = false) reduces S1 = sum dxh and S2 = sum dxh xh per (sample, group);
// pass 2 (APPLY = true) recomputes dxh and writes dx (+ the skip-path gradient `add`).
// VEC = channels per thread (8: 16-byte accesses, 4: 8-byte accesses and half the per-channel constants in
// registers -> more resident blocks)
template <int VEC> struct VecT;
template <> struct VecT<8> { using T = uint4; };
template <> struct VecT<4> { using T = uint2; };
__device__ __forceinline__ void unpackv(const uint4& r, float* f) { unpack8(r, f); }
__device__ __forceinline__ void unpackv(const uint2& r, float* f) {
  f[0] = bf16_lo(r.x); f[1] = bf16_hi(r.x);
  f[2] = bf16_lo(r.y); f[3] = bf16_hi(r.y);
}
__device__ __forceinline__ void packv(const float* o, uint4& r) { r = pack8(o); }
__device__ __forceinline__ void packv(const float* o, uint2& r) {
  r.x = pack_bf16x2(o[0], o[1]);
  r.y = pack_bf16x2(o[2], o[3]);
}

template <bool APPLY, int VEC>
__global__ void __launch_bounds__(GB_THREADS, VEC == 8 ? 2 : 3) gn_bwd_kernel(const GnBwdParams p) {
  using LT = typename VecT<VEC>::T;
  extern __shared__ float s_par[];  // [6][C]: r, m0, A, B, k1, k2
  float* s_r = s_par;
  float* s_m0 = s_par + p.C;
  float* s_A = s_par + 2 * p.C;
  float* s_B = s_par + 3 * p.C;
  float* s_k1 = s_par + 4 * p.C;
  float* s_k2 = s_par + 5 * p.C;
  __shared__ double s_acc[GN_GROUPS][2];
  const int n = blockIdx.y;
  const int V = p.C / VEC;
  const int P = p.H * p.W;
  const int cpg = p.C / GN_GROUPS;
  const double cnt = (double)cpg * (double)P;
  for (int c = threadIdx.x; c < p.C; c += GB_THREADS) {
    const int g = c / cpg;
    const double sum = p.stats[((size_t)n * GN_GROUPS + g) * 2 + 0];
    const double sq = p.stats[((size_t)n * GN_GROUPS + g) * 2 + 1];
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    float A = p.gamma[c], B = p.beta[c];
    if (p.scale_shift != nullptr) {
      const float sc = 1.0f + p.scale_shift[(size_t)n * p.ss_stride + c];
      const float sh = p.scale_shift[(size_t)n * p.ss_stride + p.C + c];
      A *= sc;
      B = fmaf(B, sc, sh);
    }
    s_r[c] = rstd;
    s_m0[c] = -(float)mean * rstd;
    s_A[c] = A;
    s_B[c] = B;
    if (APPLY) {
      s_k1[c] = (float)(p.bstats[((size_t)n * GN_GROUPS + g) * 2 + 0] / cnt);
      s_k2[c] = (float)(p.bstats[((size_t)n * GN_GROUPS + g) * 2 + 1] / cnt);
    }
  }
  if (!APPLY && threadIdx.x < GN_GROUPS * 2) (&s_acc[0][0])[threadIdx.x] = 0.0;
  __syncthreads();

  const int slots = min(V, GB_THREADS);
  const int lanes = max(1, GB_THREADS / slots);
  const int pl = threadIdx.x / slots;
  const int per = (P + p.splits - 1) / p.splits;
  const int p_begin = blockIdx.x * per;
  const int p_end = min(P, p_begin + per);
  const int Wh = p.W / 2;
  const size_t Ph = (size_t)(p.H / 2) * Wh;
  const float dscale = (p.resample == ADB_RESAMPLE_AVGPOOL2) ? 0.25f : 1.0f;

  if (pl < lanes) {
    for (int v = threadIdx.x % slots; v < V; v += GB_THREADS) {
      float r[VEC], m0[VEC], A[VEC], B[VEC], k1[VEC], k2[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        r[i] = s_r[v * VEC + i];
        m0[i] = s_m0[v * VEC + i];
        A[i] = s_A[v * VEC + i];
        B[i] = s_B[v * VEC + i];
        if (APPLY) {
          k1[i] = s_k1[v * VEC + i];
          k2[i] = s_k2[v * VEC + i];
        }
      }
      float s1[VEC], s2[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) s1[i] = s2[i] = 0.f;
      const bool half_d = p.resample == ADB_RESAMPLE_AVGPOOL2;
      const bool half_a = p.add_mode == ADB_RES_AVGPOOL2;
      const bool has_add = APPLY && p.add_mode != ADB_RES_NONE;
      const float ascale = half_a ? 0.25f : 1.0f;
      // U pixels per trip: all of a trip's 16-byte loads (x, dout, add) are issued before any math
      constexpr int U = 4;
      for (int pix0 = p_begin + pl; pix0 < p_end; pix0 += U * lanes) {
        LT xr[U], dr[U], ar[U];
        size_t ipix[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = pix0 + u * lanes;
          const bool ok = pix < p_end;
          const int pc = ok ? pix : p_begin;  // clamp: the tail re-reads a valid pixel and is discarded
          ipix[u] = (size_t)n * P + pc;
          size_t hpix = 0;  // the matching pixel of a half-resolution gradient
          if (half_d || half_a) {
            const int y = pc / p.W, xx = pc - y * p.W;
            hpix = (size_t)n * Ph + (size_t)(y >> 1) * Wh + (xx >> 1);
          }
          xr[u] = __ldg(reinterpret_cast<const LT*>(p.x + ipix[u] * p.C) + v);
          dr[u] = __ldg(reinterpret_cast<const LT*>(p.dout + (half_d ? hpix : ipix[u]) * p.C) + v);
          if (has_add) ar[u] = __ldg(reinterpret_cast<const LT*>(p.add + (half_a ? hpix : ipix[u]) * p.C) + v);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (pix0 + u * lanes >= p_end) break;
          float xf[VEC], df[VEC], o[VEC];
          unpackv(xr[u], xf);
          unpackv(dr[u], df);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float xh = fmaf(xf[i], r[i], m0[i]);
            const float z = fmaf(xh, A[i], B[i]);
            float dz = df[i] * dscale;
            if (p.silu) {
              float th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
              const float sg = fmaf(0.5f, th, 0.5f);  // sigmoid(z)
              dz *= sg * fmaf(z, 1.0f - sg, 1.0f);     // silu'(z) = s (1 + z (1 - s))
            }
            const float dxh = dz * A[i];
            if (APPLY) {
              o[i] = r[i] * (dxh - k1[i] - xh * k2[i]);
            } else {
              s1[i] += dxh;
              s2[i] = fmaf(dxh, xh, s2[i]);
            }
          }
          if (APPLY) {
            if (has_add) {
              float af[VEC];
              unpackv(ar[u], af);
#pragma unroll
              for (int i = 0; i < VEC; ++i) o[i] = fmaf(af[i], ascale, o[i]);
            }
            LT ov;
            packv(o, ov);
            *(reinterpret_cast<LT*>(p.dx + ipix[u] * p.C) + v) = ov;
          }
        }
      }
      if (!APPLY) {
        int g_cur = (v * VEC) / cpg;
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const int g = (v * VEC + i) / cpg;
          if (g != g_cur) {
            atomicAdd(&s_acc[g_cur][0], d1);
            atomicAdd(&s_acc[g_cur][1], d2);
            d1 = d2 = 0.0;
            g_cur = g;
          }
          d1 += (double)s1[i];
          d2 += (double)s2[i];
        }
        atomicAdd(&s_acc[g_cur][0], d1);
        atomicAdd(&s_acc[g_cur][1], d2);
      }
    }
  }
  if (!APPLY) {
    __syncthreads();
    if (threadIdx.x < GN_GROUPS * 2)
      atomicAdd(p.bstats + (size_t)n * GN_GROUPS * 2 + threadIdx.x, (&s_acc[0][0])[threadIdx.x]);
  }
}

// ---- AttentionPool2d (unet.py:22-51), restricted to the token it returns (x[:, :, 0]) ----
// tokens = [mean_p(h) | h_p] + positional_embedding; only token 0's query is ever used.
// prepare: xp[n,p,:] = h[n,p,:] + pos[:,1+p] (bf16) and mean[n,:] = mean_p h[n,p,:] + pos[:,0] (fp32)
__global__ void __launch_bounds__(256) pool_prepare_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ pos,
                                                          __nv_bfloat16* __restrict__ xp, float* __restrict__ mean,
                                                          int P, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int pp = 0; pp < P; ++pp) {
      const float v = __bfloat162float(h[((size_t)n * P + pp) * C + c]);
      acc += v;
      xp[((size_t)n * P + pp) * C + c] = __float2bfloat16(v + pos[(size_t)c * (P + 1) + 1 + pp]);
    }
    mean[(size_t)n * C + c] = acc / (float)P + pos[(size_t)c * (P + 1)];
  }
}

// one CTA per (sample, head): scores of query token 0 against the P+1 keys, softmax, weighted values.
// qkv0 fp32 [n, 3C] = (q | k | v) of the mean token; kv bf16 [n, P, 2C] = (k | v) of the pixel tokens.
__global__ void __launch_bounds__(128) pool_attn_fwd_kernel(const float* __restrict__ qkv0, const __nv_bfloat16* __restrict__ kv,
                                                           float* __restrict__ out0, float* __restrict__ probs,
                                                           int P, int C) {
  __shared__ float s_q[64], s_p[128], s_red[2];
  const int heads = C / 64;
  const int n = blockIdx.x / heads, hd = blockIdx.x % heads;
  const int tid = threadIdx.x;
  if (tid < 64) s_q[tid] = qkv0[(size_t)n * 3 * C + hd * 64 + tid];
  __syncthreads();
  float s = -INFINITY;
  if (tid <= P) {
    float acc = 0.f;
    if (tid == 0) {
      for (int d = 0; d < 64; ++d) acc = fmaf(s_q[d], qkv0[(size_t)n * 3 * C + C + hd * 64 + d], acc);
    } else {
      const __nv_bfloat16* kp = kv + ((size_t)n * P + (tid - 1)) * 2 * C + hd * 64;
      for (int d = 0; d < 64; ++d) acc = fmaf(s_q[d], __bfloat162float(kp[d]), acc);
    }
    s = acc * 0.125f;  // (q 64^-1/4) . (k 64^-1/4)
  }
  s_p[tid] = s;
  __syncthreads();
  if (tid == 0) {
    float m = -INFINITY;
    for (int j = 0; j <= P; ++j) m = fmaxf(m, s_p[j]);
    float l = 0.f;
    for (int j = 0; j <= P; ++j) l += __expf(s_p[j] - m);
    s_red[0] = m;
    s_red[1] = 1.0f / l;
  }
  __syncthreads();
  if (tid <= P) {
    const float pj = __expf(s - s_red[0]) * s_red[1];
    s_p[tid] = pj;
    probs[((size_t)n * heads + hd) * (P + 1) + tid] = pj;
  }
  __syncthreads();
  if (tid < 64) {
    float acc = s_p[0] * qkv0[(size_t)n * 3 * C + 2 * C + hd * 64 + tid];
    for (int j = 1; j <= P; ++j)
      acc = fmaf(s_p[j], __bfloat162float(kv[((size_t)n * P + (j - 1)) * 2 * C + C + hd * 64 + tid]), acc);
    out0[(size_t)n * C + hd * 64 + tid] = acc;
  }
}

// backward of the above: dout0 fp32 [n, C] -> dqkv0 fp32 [n, 3C], dkv bf16 [n, P, 2C]
__global__ void __launch_bounds__(128) pool_attn_bwd_kernel(const float* __restrict__ dout0, const float* __restrict__ probs,
                                                           const float* __restrict__ qkv0, const __nv_bfloat16* __restrict__ kv,
                                                           float* __restrict__ dqkv0, __nv_bfloat16* __restrict__ dkv,
                                                           int P, int C) {
  __shared__ float s_q[64], s_do[64], s_ds[128], s_dp[128], s_sum;
  const int heads = C / 64;
  const int n = blockIdx.x / heads, hd = blockIdx.x % heads;
  const int tid = threadIdx.x;
  if (tid < 64) {
    s_q[tid] = qkv0[(size_t)n * 3 * C + hd * 64 + tid];
    s_do[tid] = dout0[(size_t)n * C + hd * 64 + tid];
  }
  __syncthreads();
  float pj = 0.f, dp = 0.f;
  if (tid <= P) {
    pj = probs[((size_t)n * heads + hd) * (P + 1) + tid];
    if (tid == 0) {
      for (int d = 0; d < 64; ++d) dp = fmaf(s_do[d], qkv0[(size_t)n * 3 * C + 2 * C + hd * 64 + d], dp);
    } else {
      const __nv_bfloat16* vp = kv + ((size_t)n * P + (tid - 1)) * 2 * C + C + hd * 64;
      for (int d = 0; d < 64; ++d) dp = fmaf(s_do[d], __bfloat162float(vp[d]), dp);
    }
  }
  s_dp[tid] = pj * dp;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int j = 0; j <= P; ++j) t += s_dp[j];
    s_sum = t;
  }
  __syncthreads();
  if (tid <= P) {
    const float ds = pj * (dp - s_sum) * 0.125f;
    s_ds[tid] = ds;
    if (tid == 0) {
      for (int d = 0; d < 64; ++d) {
        dqkv0[(size_t)n * 3 * C + C + hd * 64 + d] = ds * s_q[d];        // dk of the mean token
        dqkv0[(size_t)n * 3 * C + 2 * C + hd * 64 + d] = pj * s_do[d];   // dv of the mean token
      }
    } else {
      __nv_bfloat16* dk = dkv + ((size_t)n * P + (tid - 1)) * 2 * C + hd * 64;
      __nv_bfloat16* dv = dk + C;
      for (int d = 0; d < 64; ++d) {
        dk[d] = __float2bfloat16(ds * s_q[d]);
        dv[d] = __float2bfloat16(pj * s_do[d]);
      }
    }
  }
  __syncthreads();
  if (tid < 64) {  // dq0[d] = sum_j ds_j k_j[d]
    float acc = s_ds[0] * qkv0[(size_t)n * 3 * C + C + hd * 64 + tid];
    for (int j = 1; j <= P; ++j)
      acc = fmaf(s_ds[j], __bfloat162float(kv[((size_t)n * P + (j - 1)) * 2 * C + hd * 64 + tid]), acc);
    dqkv0[(size_t)n * 3 * C + hd * 64 + tid] = acc;
  }
}

// dh[n,p,:] = dxp[n,p,:] + dmean[n,:] / P
__global__ void __launch_bounds__(256) pool_merge_kernel(const __nv_bfloat16* __restrict__ dxp, const float* __restrict__ dmean,
                                                        __nv_bfloat16* __restrict__ dh, int n, int P, int C) {
  const size_t total = (size_t)n * P * C;
  const float inv = 1.0f / (float)P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t nn = i / ((size_t)P * C);
    dh[i] = __float2bfloat16(__bfloat162float(dxp[i]) + dmean[nn * C + c] * inv);
  }
}

// dlogits[n, c] = scale * ((c == y[n]) - softmax(logits[n])[c])
__global__ void __launch_bounds__(256) logsoftmax_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ y,
                                                             float* __restrict__ dlogits, int K, float scale) {
  __shared__ float s_red[256];
  const int n = blockIdx.x;
  const float* l = logits + (size_t)n * K;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < K; c += 256) m = fmaxf(m, l[c]);
  s_red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + s]);
    __syncthreads();
  }
  m = s_red[0];
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < K; c += 256) sum += __expf(l[c] - m);
  s_red[threadIdx.x] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  const float inv = 1.0f / s_red[0];
  const int yy = (int)y[n];
  for (int c = threadIdx.x; c < K; c += 256)
    dlogits[(size_t)n * K + c] = scale * ((c == yy ? 1.0f : 0.0f) - __expf(l[c] - m) * inv);
}

}  // namespace

int gn_backward_submit(adb_plan* plan, const adb_gn_bwd_desc* d, cudaStream_t stream) {
  ADB_REQUIRE(d != nullptr, "gn_backward: null descriptor");
  ADB_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->c % GN_GROUPS == 0 && d->c % 8 == 0,
              "gn_backward: bad geometry (c must be a multiple of 32)");
  ADB_REQUIRE(d->x && d->dout && d->dx && d->stats && d->bstats && d->gamma && d->beta, "gn_backward: null pointer");
  ADB_REQUIRE(d->resample == ADB_RESAMPLE_NONE || d->resample == ADB_RESAMPLE_AVGPOOL2, "gn_backward: resample must be none or avgpool2");
  ADB_REQUIRE(d->add_mode == ADB_RES_NONE || ((d->add_mode == ADB_RES_SAME || d->add_mode == ADB_RES_AVGPOOL2) && d->add),
              "gn_backward: add_mode must be none / same / avgpool2 with a tensor");
  if (d->resample == ADB_RESAMPLE_AVGPOOL2 || d->add_mode == ADB_RES_AVGPOOL2)
    ADB_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, "gn_backward: avgpool2 needs even h,w");
  ADB_REQUIRE(6 * d->c * sizeof(float) <= 48 * 1024, "gn_backward: too many channels (%d)", d->c);
  GnBwdParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->x);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(d->dout);
  p.add = reinterpret_cast<const __nv_bfloat16*>(d->add);
  p.dx = reinterpret_cast<__nv_bfloat16*>(d->dx);
  p.stats = d->stats;
  p.bstats = d->bstats;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.scale_shift = d->scale_shift;
  p.ss_stride = d->ss_stride;
  p.eps = d->eps;
  p.n = d->n;
  p.H = d->h;
  p.W = d->w;
  p.C = d->c;
  p.silu = d->silu;
  p.resample = d->resample;
  p.add_mode = d->add_mode;
  const int P = d->h * d->w;
  int splits = (8 * num_sms() + d->n - 1) / d->n;
  const int max_splits = (P / 64) > 1 ? (P / 64) : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  const double elems = (double)d->n * P * d->c;
  const double dscale = d->resample == ADB_RESAMPLE_AVGPOOL2 ? 0.25 : 1.0;
  // ALGORITHMIC traffic (the roofline numerator): x, dout, add read once, dx written once, 2 bytes each. The two passes
  // below execute 1.5x that (x and dout are read by both): a cluster-per-sample single-launch form that re-reads its slice
  // from L2 was measured SLOWER (0.36-0.40 ms vs 0.31 ms at 256 x 64x64x128; profiles/README.md) and is not kept.
  const double bytes = 2.0 * elems * (1.0 + dscale + (d->add_mode == ADB_RES_NONE ? 0.0 : (d->add_mode == ADB_RES_AVGPOOL2 ? 0.25 : 1.0)) + 1.0);
  const int ready = d->bstats_ready ? 1 : 0;
  return submit(plan, stream, "groupnorm_bwd", 0.0, bytes, [p, ready](cudaStream_t s) -> int {
    dim3 grid(p.splits, p.n);
    const size_t smem = 6 * (size_t)p.C * sizeof(float);
    if (ready) {  // the two sums came out of the producing data-gradient conv's epilogue (adb_conv_desc.gnb_*)
      gn_bwd_kernel<true, 4><<<grid, GB_THREADS, smem, s>>>(p);
      ADB_CUDA(cudaGetLastError());
      return 1;
    }
    ADB_CUDA(cudaMemsetAsync(p.bstats, 0, (size_t)p.n * GN_GROUPS * 2 * sizeof(double), s));
    // 4 channels per thread (80 registers, 3 blocks/SM) measured 3.9 TB/s vs 3.5 TB/s for 8 channels per thread
    // (128 registers, 2 blocks/SM) and 2.9 TB/s for the first version (157 registers, 1 block/SM)
    static int vec4 = -1;
    if (vec4 < 0) {
      const char* e = getenv("ADB_GNB_VEC8");
      vec4 = (e && e[0] == '1') ? 0 : 1;
    }
    if (vec4) {
      gn_bwd_kernel<false, 4><<<grid, GB_THREADS, smem, s>>>(p);
      ADB_CUDA(cudaGetLastError());
      gn_bwd_kernel<true, 4><<<grid, GB_THREADS, smem, s>>>(p);
    } else {
      gn_bwd_kernel<false, 8><<<grid, GB_THREADS, smem, s>>>(p);
      ADB_CUDA(cudaGetLastError());
      gn_bwd_kernel<true, 8><<<grid, GB_THREADS, smem, s>>>(p);
    }
    ADB_CUDA(cudaGetLastError());
    return 3;
  });
}

int pool_prepare_submit(adb_plan* plan, const void* h, const float* pos, void* xp, float* mean, int n, int P, int C,
                        cudaStream_t stream) {
  ADB_REQUIRE(h && pos && xp && mean && n > 0 && P > 0 && C > 0, "pool_prepare: bad arguments");
  return submit(plan, stream, "pool_prepare", 0.0, 0.0, [=](cudaStream_t s) -> int {
    pool_prepare_kernel<<<n, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(h), pos,
                                          reinterpret_cast<__nv_bfloat16*>(xp), mean, P, C);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pool_attention_submit(adb_plan* plan, const float* qkv0, const void* kv, float* out0, float* probs, int n, int P, int C,
                          cudaStream_t stream) {
  ADB_REQUIRE(qkv0 && kv && out0 && probs && n > 0 && P > 0 && P + 1 <= 128 && C % 64 == 0,
              "pool_attention: bad arguments (P + 1 <= 128, C %% 64 == 0)");
  return submit(plan, stream, "pool_attention", 0.0, 0.0, [=](cudaStream_t s) -> int {
    pool_attn_fwd_kernel<<<n * (C / 64), 128, 0, s>>>(qkv0, reinterpret_cast<const __nv_bfloat16*>(kv), out0, probs, P, C);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pool_attention_backward_submit(adb_plan* plan, const float* dout0, const float* probs, const float* qkv0, const void* kv,
                                   float* dqkv0, void* dkv, int n, int P, int C, cudaStream_t stream) {
  ADB_REQUIRE(dout0 && probs && qkv0 && kv && dqkv0 && dkv && n > 0 && P > 0 && P + 1 <= 128 && C % 64 == 0,
              "pool_attention_backward: bad arguments");
  return submit(plan, stream, "pool_attention_bwd", 0.0, 0.0, [=](cudaStream_t s) -> int {
    pool_attn_bwd_kernel<<<n * (C / 64), 128, 0, s>>>(dout0, probs, qkv0, reinterpret_cast<const __nv_bfloat16*>(kv), dqkv0,
                                                      reinterpret_cast<__nv_bfloat16*>(dkv), P, C);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pool_merge_submit(adb_plan* plan, const void* dxp, const float* dmean, void* dh, int n, int P, int C, cudaStream_t stream) {
  ADB_REQUIRE(dxp && dmean && dh && n > 0 && P > 0 && C > 0, "pool_merge: bad arguments");
  return submit(plan, stream, "pool_merge", 0.0, 0.0, [=](cudaStream_t s) -> int {
    const size_t total = (size_t)n * P * C;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    pool_merge_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dxp), dmean,
                                                       reinterpret_cast<__nv_bfloat16*>(dh), n, P, C);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int logsoftmax_grad_submit(adb_plan* plan, const float* logits, const int64_t* y, float* dlogits, int n, int k, float scale,
                           cudaStream_t stream) {
  ADB_REQUIRE(logits && y && dlogits && n > 0 && k > 0, "logsoftmax_grad: bad arguments");
  return submit(plan, stream, "logsoftmax_grad", 0.0, 0.0, [=](cudaStream_t s) -> int {
    logsoftmax_grad_kernel<<<n, 256, 0, s>>>(logits, y, dlogits, k, scale);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
