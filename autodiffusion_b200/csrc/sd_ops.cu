// sd_ops.cu — the memory-bound pieces of the Stable-Diffusion transformer blocks and its sampler step.
//
// Reference ("Stable Diffusion"/ldm/...): nn.LayerNorm in BasicTransformerBlock (modules/attention.py:205-216),
// GEGLU (modules/attention.py:37-44), p_sample_ddim with classifier-free guidance (models/diffusion/ddim.py:177-217).
// All HBM-bound: 16-byte vector loads, one pass over the data, statistics in registers.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int LN_MAX_VEC = 8;  // 16-byte vectors per lane: c <= 32 * 8 * 8 = 2048

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// Tried and dropped (48 LayerNorms of an SD forward at batch 64: 2.6 ms with this kernel): two rows in flight per warp
// (3.0 ms), the lane's gamma / beta slice kept in registers across rows (3.9 ms - occupancy).
// One warp per token row; the row stays in registers between the mean, the variance and the normalise pass.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const uint4* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, uint4* __restrict__ out,
                                                        int rows, int nvec, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_c = 1.0f / (float)(nvec * 8);
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const uint4* xr = x + (size_t)row * nvec;
    float f[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const uint4 v = __ldg(xr + vi);
        unpack8(v, f[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[i][k];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dlt = f[i][k] - mean;
          q = fmaf(dlt, dlt, q);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float rstd = rsqrtf(q * inv_c + eps);
    uint4* orow = out + (size_t)row * nvec;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
        uint4 o;
        o.x = pack_bf16x2(fmaf((f[i][0] - mean) * rstd, g0.x, b0.x), fmaf((f[i][1] - mean) * rstd, g0.y, b0.y));
        o.y = pack_bf16x2(fmaf((f[i][2] - mean) * rstd, g0.z, b0.z), fmaf((f[i][3] - mean) * rstd, g0.w, b0.w));
        o.z = pack_bf16x2(fmaf((f[i][4] - mean) * rstd, g1.x, b1.x), fmaf((f[i][5] - mean) * rstd, g1.y, b1.y));
        o.w = pack_bf16x2(fmaf((f[i][6] - mean) * rstd, g1.z, b1.z), fmaf((f[i][7] - mean) * rstd, g1.w, b1.w));
        orow[vi] = o;
      }
    }
  }
}

// erf by Abramowitz-Stegun 7.1.26 (absolute error <= 1.5e-7, far below the bf16 rounding of the product): one
// reciprocal, one exp2 and five FMAs instead of erff's ~25 instructions - with erff the kernel sat at ~60 % of the
// issue slots and 4.1 TB/s; the GELU is F.gelu's exact (erf) form, not the tanh approximation.
__device__ __forceinline__ float gelu_erf(float g) {
  const float z = fabsf(g) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-p * t, e, 1.0f);  // erf(|x| / sqrt 2)
  return 0.5f * g * (1.0f + copysignf(erf_abs, g));
}

// out[r, j] = a[r, j] * gelu(gate[r, j]); a = x[:, :inner], gate = x[:, inner:]
__global__ void __launch_bounds__(256) geglu_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, size_t total_vec,
                                                    int inner_vec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / inner_vec;
    const size_t j = i - r * inner_vec;
    const uint4 av = __ldg(x + r * 2 * inner_vec + j);
    const uint4 gv = __ldg(x + r * 2 * inner_vec + inner_vec + j);
    float a[8], g[8];
    unpack8(av, a);
    unpack8(gv, g);
    uint4 o;
    o.x = pack_bf16x2(a[0] * gelu_erf(g[0]), a[1] * gelu_erf(g[1]));
    o.y = pack_bf16x2(a[2] * gelu_erf(g[2]), a[3] * gelu_erf(g[3]));
    o.z = pack_bf16x2(a[4] * gelu_erf(g[4]), a[5] * gelu_erf(g[5]));
    o.w = pack_bf16x2(a[6] * gelu_erf(g[6]), a[7] * gelu_erf(g[7]));
    out[i] = o;
  }
}

struct CfgCoef {
  float v[4];
};

// Every operation of ddim.py:187-216 keeps its own fp32 rounding (explicit _rn intrinsics: no FMA contraction).
__global__ void __launch_bounds__(256) cfg_ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                            float* __restrict__ x_prev, float* __restrict__ pred_x0,
                                                            size_t total, int cfg, float scale, CfgCoef cf) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float e = eps[i];
    if (cfg) {
      const float eu = e;
      const float ec = eps[total + i];
      e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
    }
    const float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(cf.v[0], e)), cf.v[1]);
    const float dir = __fmul_rn(cf.v[3], e);
    x_prev[i] = __fadd_rn(__fmul_rn(cf.v[2], x0), dir);
    if (pred_x0 != nullptr) pred_x0[i] = x0;
  }
}

// e_t = e_u + scale * (e_c - e_u) (plms.py:202-207 / ddim.py:187-190), kept as its own tensor for the multistep history
__global__ void __launch_bounds__(256) cfg_combine_kernel(const float* __restrict__ eps, float* __restrict__ e_out,
                                                          size_t total, int cfg, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float e = eps[i];
    if (cfg) e = __fadd_rn(e, __fmul_rn(scale, __fsub_rn(eps[total + i], e)));
    e_out[i] = e;
  }
}

// p_sample_plms after the model call (plms.py:241-257): e' from the eps history, then the eta = 0 DDIM update with e'.
// mode 0: e' = e_t; 1: (e_t + o1) / 2 (o1 = e_t_next, pseudo improved Euler); 2: (3 e_t - o1) / 2;
// 3: (23 e_t - 16 o1 + 5 o2) / 12; 4: (55 e_t - 59 o1 + 37 o2 - 9 o3) / 24 - each product, sum and division rounded
// separately, left to right, as the reference's tensor expression evaluates.
__global__ void __launch_bounds__(256) plms_update_kernel(const float* __restrict__ x, const float* __restrict__ e_t,
                                                          const float* __restrict__ o1, const float* __restrict__ o2,
                                                          const float* __restrict__ o3, int mode, CfgCoef cf,
                                                          float* __restrict__ x_prev, float* __restrict__ pred_x0,
                                                          size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float e = e_t[i];
    if (mode == 1) {
      e = __fdiv_rn(__fadd_rn(e, o1[i]), 2.0f);
    } else if (mode == 2) {
      e = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e), o1[i]), 2.0f);
    } else if (mode == 3) {
      e = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e), __fmul_rn(16.0f, o1[i])), __fmul_rn(5.0f, o2[i])), 12.0f);
    } else if (mode == 4) {
      const float a = __fsub_rn(__fmul_rn(55.0f, e), __fmul_rn(59.0f, o1[i]));
      e = __fdiv_rn(__fsub_rn(__fadd_rn(a, __fmul_rn(37.0f, o2[i])), __fmul_rn(9.0f, o3[i])), 24.0f);
    }
    const float x0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(cf.v[0], e)), cf.v[1]);
    x_prev[i] = __fadd_rn(__fmul_rn(cf.v[2], x0), __fmul_rn(cf.v[3], e));
    if (pred_x0 != nullptr) pred_x0[i] = x0;
  }
}

// DPM-Solver++ data prediction (dpm_solver.py:336-343, 386-391): noise = e_u + scale (e_c - e_u),
// x0 = (x - sigma_t * noise) / alpha_t, each op rounded separately.
__global__ void __launch_bounds__(256) dpm_x0_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                     float* __restrict__ x0, size_t total, int cfg, float scale, float sigma,
                                                     float alpha) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float e = eps[i];
    if (cfg) e = __fadd_rn(e, __fmul_rn(scale, __fsub_rn(eps[total + i], e)));
    x0[i] = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(sigma, e)), alpha);
  }
}

// multistep updates (dpm_solver.py:519-533, 770-790, data-prediction branch, 'dpm_solver' type):
// order 1: x_t = c0 x - c1 m0;  order 2: x_t = (c0 x - c1 m0) - c2 (inv_r0 (m0 - m1)), with
// c0 = sigma_t / sigma_s, c1 = alpha_t * expm1(-h) or alpha_t * (exp(-h) - 1), c2 = 0.5 * c1 - computed by the host
// in fp32 as the reference computes them.
__global__ void __launch_bounds__(256) dpm_update_kernel(const float* __restrict__ x, const float* __restrict__ m0,
                                                         const float* __restrict__ m1, float* __restrict__ x_out, size_t total,
                                                         int order, float c0, float c1, float c2, float inv_r0) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float a = __fmul_rn(c0, x[i]);
    const float mm = m0[i];
    float r = __fsub_rn(a, __fmul_rn(c1, mm));
    if (order == 2) r = __fsub_rn(r, __fmul_rn(c2, __fmul_rn(inv_r0, __fsub_rn(mm, m1[i]))));
    x_out[i] = r;
  }
}

__global__ void __launch_bounds__(256) pad_context_kernel(const float* __restrict__ ctx, __nv_bfloat16* __restrict__ out,
                                                          int n, int t, int c, int t_pad) {
  const size_t total = (size_t)n * t_pad * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / c;
    const int col = (int)(i - row * c);
    const int img = (int)(row / t_pad);
    const int tok = (int)(row - (size_t)img * t_pad);
    out[i] = __float2bfloat16(tok < t ? ctx[((size_t)img * t + tok) * c + col] : 0.0f);
  }
}

unsigned grid_for(size_t work_items, int per_block) {
  size_t blocks = (work_items + per_block - 1) / per_block;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return blocks ? (unsigned)blocks : 1u;
}

}  // namespace

int layernorm_submit(adb_plan* plan, const void* x, const float* gamma, const float* beta, void* out, int rows, int c,
                     float eps, cudaStream_t stream) {
  ADB_REQUIRE(x && gamma && beta && out && rows > 0, "layernorm: bad arguments");
  ADB_REQUIRE(c > 0 && c % 8 == 0 && c <= 32 * 8 * LN_MAX_VEC, "layernorm: c = %d unsupported (multiple of 8, <= 2048)", c);
  const int nvec = c / 8;
  const double bytes = 4.0 * (double)rows * c;  // 2 B read + 2 B written per element
  return submit(plan, stream, "layernorm", 0.0, bytes, [=](cudaStream_t s) -> int {
    const unsigned grid = grid_for((size_t)rows, 8);
    const uint4* xi = reinterpret_cast<const uint4*>(x);
    uint4* oo = reinterpret_cast<uint4*>(out);
    if (nvec <= 64) layernorm_kernel<2><<<grid, 256, 0, s>>>(xi, gamma, beta, oo, rows, nvec, eps);
    else if (nvec <= 96) layernorm_kernel<3><<<grid, 256, 0, s>>>(xi, gamma, beta, oo, rows, nvec, eps);
    else if (nvec <= 160) layernorm_kernel<5><<<grid, 256, 0, s>>>(xi, gamma, beta, oo, rows, nvec, eps);
    else layernorm_kernel<LN_MAX_VEC><<<grid, 256, 0, s>>>(xi, gamma, beta, oo, rows, nvec, eps);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int geglu_submit(adb_plan* plan, const void* x, void* out, int rows, int inner, cudaStream_t stream) {
  ADB_REQUIRE(x && out && rows > 0 && inner > 0 && inner % 8 == 0, "geglu: bad arguments");
  const double bytes = 2.0 * 3.0 * (double)rows * inner;  // two bf16 reads + one bf16 write per output element
  return submit(plan, stream, "geglu", 0.0, bytes, [=](cudaStream_t s) -> int {
    const size_t total_vec = (size_t)rows * (inner / 8);
    geglu_kernel<<<grid_for(total_vec, 256), 256, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out),
                                                         total_vec, inner / 8);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int cfg_ddim_step_submit(adb_plan* plan, const float* x, const float* eps, float* x_prev, float* pred_x0, int n, int chw,
                         int cfg, float scale, const float coef[4], cudaStream_t stream) {
  ADB_REQUIRE(x && eps && x_prev && n > 0 && chw > 0 && coef, "cfg_ddim_step: bad arguments");
  CfgCoef cf;
  for (int i = 0; i < 4; ++i) cf.v[i] = coef[i];
  const size_t total = (size_t)n * chw;
  const double bytes = 4.0 * (double)total * (cfg ? 4.0 : 3.0);
  return submit(plan, stream, "cfg_ddim_step", 0.0, bytes, [=](cudaStream_t s) -> int {
    cfg_ddim_step_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, eps, x_prev, pred_x0, total, cfg, scale, cf);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int cfg_combine_submit(adb_plan* plan, const float* eps, float* e_out, size_t total, int cfg, float scale,
                       cudaStream_t stream) {
  ADB_REQUIRE(eps && e_out && total > 0, "cfg_combine: bad arguments");
  return submit(plan, stream, "cfg_combine", 0.0, 4.0 * (double)total * (cfg ? 3.0 : 2.0), [=](cudaStream_t s) -> int {
    cfg_combine_kernel<<<grid_for(total, 256), 256, 0, s>>>(eps, e_out, total, cfg, scale);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int plms_update_submit(adb_plan* plan, const float* x, const float* e_t, const float* o1, const float* o2, const float* o3,
                       int mode, const float coef[4], float* x_prev, float* pred_x0, size_t total, cudaStream_t stream) {
  ADB_REQUIRE(x && e_t && x_prev && coef && total > 0 && mode >= 0 && mode <= 4, "plms_update: bad arguments");
  ADB_REQUIRE((mode < 1 || o1) && (mode < 3 || o2) && (mode < 4 || o3), "plms_update: mode %d needs more eps history", mode);
  CfgCoef cf;
  for (int i = 0; i < 4; ++i) cf.v[i] = coef[i];
  return submit(plan, stream, "plms_update", 0.0, 4.0 * (double)total * (3.0 + (mode > 1 ? mode - 1 : mode)), [=](cudaStream_t s) -> int {
    plms_update_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, e_t, o1, o2, o3, mode, cf, x_prev, pred_x0, total);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int dpm_x0_submit(adb_plan* plan, const float* x, const float* eps, float* x0, size_t total, int cfg, float scale,
                  float sigma, float alpha, cudaStream_t stream) {
  ADB_REQUIRE(x && eps && x0 && total > 0, "dpm_x0: bad arguments");
  return submit(plan, stream, "dpm_x0", 0.0, 4.0 * (double)total * (cfg ? 4.0 : 3.0), [=](cudaStream_t s) -> int {
    dpm_x0_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, eps, x0, total, cfg, scale, sigma, alpha);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int dpm_update_submit(adb_plan* plan, const float* x, const float* m0, const float* m1, float* x_out, size_t total, int order,
                      float c0, float c1, float c2, float inv_r0, cudaStream_t stream) {
  ADB_REQUIRE(x && m0 && x_out && total > 0 && (order == 1 || (order == 2 && m1)), "dpm_update: bad arguments");
  return submit(plan, stream, "dpm_update", 0.0, 4.0 * (double)total * (order + 2.0), [=](cudaStream_t s) -> int {
    dpm_update_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, m0, m1, x_out, total, order, c0, c1, c2, inv_r0);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pad_context_submit(adb_plan* plan, const float* ctx, void* out, int n, int t, int c, int t_pad, cudaStream_t stream) {
  ADB_REQUIRE(ctx && out && n > 0 && t > 0 && c > 0 && t_pad >= t, "pad_context: bad arguments");
  return submit(plan, stream, "pad_context", 0.0, 0.0, [=](cudaStream_t s) -> int {
    pad_context_kernel<<<grid_for((size_t)n * t_pad * c, 256), 256, 0, s>>>(ctx, reinterpret_cast<__nv_bfloat16*>(out), n, t,
                                                                           c, t_pad);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
