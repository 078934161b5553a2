// elementwise.cu — the small kernels either side of the tensor-core path:
// input stem conv, timestep/label embedding MLP, fused guidance + DDIM update, uint8 pack.
#include "common.cuh"

namespace adb {

namespace {

// ------------------------------------------------------------------------------------
// stem: fp32 NCHW [n,cin,h,w] -> 3x3 conv (pad 1) -> bf16 NHWC [n,h,w,cout]
// reference: input_blocks.0.0 = conv_nd(dims, in_channels, ch, 3, padding=1)
// (guided_diffusion/dynamic_unet.py:501-503), input cast at :693.
// K = 9*cin <= 36: CUDA cores, weights in shared memory.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ wgt,
                                                       const float* __restrict__ bias,
                                                       __nv_bfloat16* __restrict__ out, int n, int cin,
                                                       int H, int W, int cout) {
  extern __shared__ float s_w[];  // [9*cin][cout] then bias[cout]
  const int K = 9 * cin;
  float* s_b = s_w + K * cout;
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) {
    // PyTorch layout [cout][cin][3][3] -> [(tap*cin + ci)][cout]
    const int co = i % cout;
    const int k = i / cout;
    const int tap = k / cin, ci = k - tap * cin;
    s_w[i] = wgt[((size_t)co * cin + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();

  const int V = cout / 8;
  const int slots = min(V, (int)blockDim.x);
  const int lanes = blockDim.x / slots;
  const int v = threadIdx.x % slots;
  const int pl = threadIdx.x / slots;
  if (pl >= lanes) return;
  const size_t P = (size_t)H * W;
  const size_t total = (size_t)n * P;
  for (size_t pix = (size_t)blockIdx.x * lanes + pl; pix < total; pix += (size_t)gridDim.x * lanes) {
    const int img = (int)(pix / P);
    const int rem = (int)(pix - (size_t)img * P);
    const int y = rem / W, xx = rem - y * W;
    for (int vv = v; vv < V; vv += slots) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = s_b[vv * 8 + i];
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xs = xx + tap % 3 - 1;
        if (yy < 0 || yy >= H || xs < 0 || xs >= W) continue;
        for (int ci = 0; ci < cin; ++ci) {
          const float a = __ldg(x + ((size_t)img * cin + ci) * P + (size_t)yy * W + xs);
          const float* wr = s_w + (tap * cin + ci) * cout + vv * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, wr[i], acc[i]);
        }
      }
      uint4 o;
      o.x = pack_bf16x2(acc[0], acc[1]);
      o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]);
      o.w = pack_bf16x2(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(out + pix * cout + vv * 8) = o;
    }
  }
}

// Fast path (cout % 64 == 0, W % 32 == 0): one warp = 32 consecutive pixels of a row x 64 output
// channels. The 9*cin inputs of a pixel are loaded once (coalesced along x) and kept in registers;
// weights are read from shared memory as warp-uniform float4 broadcasts; 64 fp32 accumulators.
__global__ void __launch_bounds__(256) stem_conv64_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ wgt,
                                                         const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int n, int cin,
                                                         int H, int W, int cout) {
  extern __shared__ float s_w[];  // [9*cin][cout] then bias[cout]
  const int K = 9 * cin;
  float* s_b = s_w + K * cout;
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) {
    const int co = i % cout;
    const int k = i / cout;
    const int tap = k / cin, ci = k - tap * cin;
    s_w[i] = wgt[((size_t)co * cin + ci) * 9 + tap];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int ncg = cout / 64;
  const size_t P = (size_t)H * W;
  const size_t items = (size_t)n * P / 32 * ncg;
  for (size_t item = (size_t)blockIdx.x * 8 + warp; item < items; item += (size_t)gridDim.x * 8) {
    const int cg = (int)(item % ncg);
    const size_t pix = (item / ncg) * 32 + lane;
    const int img = (int)(pix / P);
    const int rem = (int)(pix - (size_t)img * P);
    const int y = rem / W, xx = rem - y * W;
    float a[36];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xs = xx + tap % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xs >= 0 && xs < W;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
        a[tap * 4 + ci] = (ok && ci < cin) ? __ldg(x + ((size_t)img * cin + ci) * P + (size_t)yy * W + xs) : 0.f;
    }
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = s_b[cg * 64 + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci < cin) {
          const float av = a[tap * 4 + ci];
          const float4* wr = reinterpret_cast<const float4*>(s_w + (tap * cin + ci) * cout + cg * 64);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 w4 = wr[j];
            acc[4 * j + 0] = fmaf(av, w4.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(av, w4.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(av, w4.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(av, w4.w, acc[4 * j + 3]);
          }
        }
      }
    }
    uint4* op = reinterpret_cast<uint4*>(out + pix * cout + cg * 64);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 o;
      o.x = pack_bf16x2(acc[8 * g + 0], acc[8 * g + 1]);
      o.y = pack_bf16x2(acc[8 * g + 2], acc[8 * g + 3]);
      o.z = pack_bf16x2(acc[8 * g + 4], acc[8 * g + 5]);
      o.w = pack_bf16x2(acc[8 * g + 6], acc[8 * g + 7]);
      op[g] = o;
    }
  }
}

// ------------------------------------------------------------------------------------
// timestep_embedding (guided_diffusion/nn.py:103-121), fp32 as written there:
// args = t * freqs; out = [cos(args) | sin(args)]
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void timestep_embedding_kernel(const T* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ out, int b, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b * half) return;
  const int row = i / half, k = i - row * half;
  // freqs[k] = exp(-ln(max_period) * k / half) is a per-model constant table built once on the
  // host: |t * f| reaches ~1e3, so a 1-ulp difference between exp implementations would move
  // sin/cos by ~1e-4; taking the table as an input keeps this kernel a pure function of it.
  const float arg = __fmul_rn((float)t[row], freqs[k]);
  out[(size_t)row * dim + k] = cosf(arg);
  out[(size_t)row * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[(size_t)row * dim + dim - 1] = 0.f;
}

// ------------------------------------------------------------------------------------
// fp32 linear: out[b,j] = bias[j] + sum_k act(x[b,k]) W[j,k] (+ table[idx[b], j])
// (time_embed.{0,2}, label_emb add, and all ResBlock emb_layers as one product;
//  dynamic_unet.py:490-498,208-214,259,687-691). 64x64x16 smem tiles, 4x4 per thread.
// ------------------------------------------------------------------------------------
constexpr int LT_M = 64, LT_N = 64, LT_K = 16;

__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x,
                                                    const float* __restrict__ w,
                                                    const float* __restrict__ bias,
                                                    float* __restrict__ out, int B, int K, int N,
                                                    int silu_in, const float* __restrict__ table,
                                                    const int64_t* __restrict__ idx) {
  __shared__ float As[LT_K][LT_M + 4];
  __shared__ float Bs[LT_K][LT_N + 4];
  const int m0 = blockIdx.y * LT_M;
  const int n0 = blockIdx.x * LT_N;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 256 threads, each loads 4 consecutive k of one row (64 rows x 16 k)
  const int lrow = threadIdx.x / 4;
  const int lk = (threadIdx.x % 4) * 4;
  for (int k0 = 0; k0 < K; k0 += LT_K) {
    {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      const int m = m0 + lrow;
      if (m < B) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k0 + lk + i;
          if (k < K) {
            float val = x[(size_t)m * K + k];
            a[i] = silu_in ? (val / (1.0f + expf(-val))) : val;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[lk + i][lrow] = a[i];
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      const int nn = n0 + lrow;
      if (nn < N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = k0 + lk + i;
          if (k < K) bv[i] = w[(size_t)nn * K + k];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[lk + i][lrow] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LT_K; ++k) {
      float a[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bq[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= N) continue;
      float r = acc[i][j] + (bias ? bias[nn] : 0.f);
      if (table) r += table[(size_t)idx[m] * N + nn];
      out[(size_t)m * N + nn] = r;
    }
  }
}

// ------------------------------------------------------------------------------------
// fused guidance + DDIM update, eta = 0. Reference chain, op for op in fp32
// (gaussian_diffusion.py): p_mean_variance:309-311 (+clamp :296-297), condition_score
// :381-389, ddim_sample :565-583. Explicit _rn intrinsics keep every intermediate rounding
// of the reference's separate elementwise ops (no FMA contraction) so that, given the same
// eps / grad, x_{t-1} is bit-identical to the reference's.
//   coef[0]=sqrt_recip_alphas_cumprod[i]  coef[1]=sqrt_recipm1_alphas_cumprod[i]
//   coef[2]=sqrt(1-alpha_bar)  coef[3]=sqrt(alpha_bar_prev)  coef[4]=sqrt(1-alpha_bar_prev)
// ------------------------------------------------------------------------------------
struct DdimCoef {
  float v[5];
};

__global__ void __launch_bounds__(256) ddim_step_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ model_out,
                                                       int eps_channels, const float* __restrict__ grad,
                                                       float* __restrict__ x_prev,
                                                       float* __restrict__ pred_xstart, int n, int c,
                                                       int hw, DdimCoef cf, int clip) {
  const size_t per = (size_t)c * hw;
  const size_t total = (size_t)n * per;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / per;
    const size_t r = i - img * per;
    const float xv = x[i];
    const float eps = model_out[img * (size_t)eps_channels * hw + r];
    const float ax = __fmul_rn(cf.v[0], xv);
    float x0 = __fsub_rn(ax, __fmul_rn(cf.v[1], eps));
    if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    if (grad != nullptr) {
      float e = __fdiv_rn(__fsub_rn(ax, x0), cf.v[1]);
      e = __fsub_rn(e, __fmul_rn(cf.v[2], grad[i]));
      x0 = __fsub_rn(ax, __fmul_rn(cf.v[1], e));
    }
    const float e3 = __fdiv_rn(__fsub_rn(ax, x0), cf.v[1]);
    x_prev[i] = __fadd_rn(__fmul_rn(x0, cf.v[3]), __fmul_rn(cf.v[4], e3));
    if (pred_xstart != nullptr) pred_xstart[i] = x0;
  }
}

// ((s+1)*127.5).clamp(0,255).to(uint8), NCHW -> NHWC (…progressive.py:421-423)
__global__ void __launch_bounds__(256) pack_uint8_kernel(const float* __restrict__ s,
                                                        uint8_t* __restrict__ out, int n, int c,
                                                        int hw) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw;
    const size_t p = i - img * hw;
    for (int ch = 0; ch < c; ++ch) {
      float v = __fmul_rn(__fadd_rn(s[(img * c + ch) * hw + p], 1.0f), 127.5f);
      v = fminf(fmaxf(v, 0.0f), 255.0f);
      out[i * c + ch] = (uint8_t)v;  // truncation, as .to(th.uint8)
    }
  }
}

}  // namespace

int stem_conv_submit(adb_plan* plan, const float* x, const float* weight, const float* bias, void* out,
                     int n, int cin, int h, int w, int cout, cudaStream_t stream) {
  ADB_REQUIRE(x && weight && out, "stem_conv: null pointer");
  ADB_REQUIRE(n > 0 && cin > 0 && cin <= 4 && h > 0 && w > 0, "stem_conv: bad geometry (cin<=4)");
  ADB_REQUIRE(cout > 0 && cout % 8 == 0, "stem_conv: cout %% 8 != 0");
  const size_t smem = ((size_t)9 * cin * cout + cout) * sizeof(float);
  ADB_REQUIRE(smem <= 48 * 1024, "stem_conv: weights do not fit shared memory");
  return submit(plan, stream, "stem_conv", 2.0 * 9 * cin * (double)cout * n * h * w, 0.0, [=](cudaStream_t s) -> int {
    const size_t total = (size_t)n * h * w;
    if (cout % 64 == 0 && w % 32 == 0) {
      const size_t items = total / 32 * (cout / 64);
      size_t blocks = (items + 7) / 8;
      const size_t cap = (size_t)num_sms() * 16;
      if (blocks > cap) blocks = cap;
      stem_conv64_kernel<<<(unsigned)blocks, 256, smem, s>>>(x, weight, bias, reinterpret_cast<__nv_bfloat16*>(out),
                                                             n, cin, h, w, cout);
    } else {
      const int V = cout / 8;
      const int slots = V < 256 ? V : 256;
      const int lanes = 256 / slots;
      size_t blocks = (total + lanes - 1) / lanes;
      const size_t cap = (size_t)num_sms() * 8;
      if (blocks > cap) blocks = cap;
      stem_conv_kernel<<<(unsigned)blocks, 256, smem, s>>>(x, weight, bias, reinterpret_cast<__nv_bfloat16*>(out), n,
                                                           cin, h, w, cout);
    }
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int timestep_embedding_submit(adb_plan* plan, const int64_t* t, const float* freqs, float* out, int b,
                              int dim, cudaStream_t stream) {
  ADB_REQUIRE(t && freqs && out && b > 0 && dim >= 2, "timestep_embedding: bad arguments");
  return submit(plan, stream, "timestep_embedding", 0.0, 0.0, [=](cudaStream_t s) -> int {
    const int total = b * (dim / 2);
    timestep_embedding_kernel<int64_t><<<(total + 127) / 128, 128, 0, s>>>(t, freqs, out, b, dim);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

// fractional model timesteps (DPM-Solver feeds (t - 1/N) * 1000, dpm_solver.py:278-286)
int timestep_embedding_f32_submit(adb_plan* plan, const float* t, const float* freqs, float* out, int b, int dim,
                                  cudaStream_t stream) {
  ADB_REQUIRE(t && freqs && out && b > 0 && dim >= 2, "timestep_embedding_f32: bad arguments");
  return submit(plan, stream, "timestep_embedding", 0.0, 0.0, [=](cudaStream_t s) -> int {
    const int total = b * (dim / 2);
    timestep_embedding_kernel<float><<<(total + 127) / 128, 128, 0, s>>>(t, freqs, out, b, dim);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int linear_submit(adb_plan* plan, const float* x, const float* w, const float* bias, float* out, int b,
                  int k, int nout, int silu_in, const float* table, const int64_t* idx,
                  cudaStream_t stream) {
  ADB_REQUIRE(x && w && out && b > 0 && k > 0 && nout > 0, "linear: bad arguments");
  ADB_REQUIRE((table == nullptr) == (idx == nullptr), "linear: table and idx go together");
  return submit(plan, stream, "linear", 2.0 * b * (double)k * nout, 0.0, [=](cudaStream_t s) -> int {
    dim3 grid((nout + LT_N - 1) / LT_N, (b + LT_M - 1) / LT_M);
    linear_kernel<<<grid, 256, 0, s>>>(x, w, bias, out, b, k, nout, silu_in, table, idx);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

// Stem im2col: fp32 NCHW [n,3,h,w] -> bf16 [n,h,w,64], channel k = tap*3 + ci of the 3x3 zero-padded
// neighbourhood for k < 27, the bf16 rounding RESIDUAL of the same value at 32 + k, zeros elsewhere. One 64-deep
// tensor-core k-step over [hi | lo] x [W | W] then is the whole input convolution (input_blocks.0.0,
// dynamic_unet.py:501-503) with ~16 mantissa bits of x kept (the reference feeds x.type(fp16), :693).
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         int n, int H, int W) {
  const size_t P = (size_t)H * W;
  const size_t total = (size_t)n * P;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
    const size_t img = pix / P;
    const int rem = (int)(pix - img * P);
    const int y = rem / W, xx = rem - y * W;
    uint32_t hi[16], lo[16];  // 32 bf16 each, packed in pairs
    float v[28];
    v[27] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xs = xx + tap % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xs >= 0 && xs < W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        v[tap * 3 + ci] = ok ? __ldg(x + (img * 3 + ci) * P + (size_t)yy * W + xs) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      const __nv_bfloat16 h0 = __float2bfloat16(v[2 * i]), h1 = __float2bfloat16(v[2 * i + 1]);
      __nv_bfloat162 hp;
      hp.x = h0;
      hp.y = h1;
      hi[i] = *reinterpret_cast<uint32_t*>(&hp);
      lo[i] = pack_bf16x2(v[2 * i] - __bfloat162float(h0), v[2 * i + 1] - __bfloat162float(h1));
    }
    hi[14] = hi[15] = lo[14] = lo[15] = 0u;
    uint4* o = reinterpret_cast<uint4*>(out + pix * 64);
#pragma unroll
    for (int g = 0; g < 4; ++g) o[g] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
#pragma unroll
    for (int g = 0; g < 4; ++g) o[4 + g] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
  }
}

int stem_im2col_submit(adb_plan* plan, const float* x, void* out, int n, int h, int w, cudaStream_t stream) {
  ADB_REQUIRE(x && out && n > 0 && h > 0 && w > 0, "stem_im2col: bad arguments");
  return submit(plan, stream, "stem_im2col", 0.0, 0.0, [=](cudaStream_t s) -> int {
    const size_t total = (size_t)n * h * w;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    stem_im2col_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n, h, w);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

// act(x) (SiLU or identity) split into two bf16 addends: hi = bf16(v), lo = bf16(v - hi). hi + lo carries
// 16 mantissa bits, so a bf16 tensor-core product over [hi | lo | hi] x [W_hi | W_hi | W_lo] reproduces the
// fp32 Linear to ~2^-16 relative (the reference keeps its embedding MLP in fp32, fp16_util.py:15-22).
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ lo, size_t total, int silu_in) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (silu_in) v = v / (1.0f + expf(-v));
    const __nv_bfloat16 h = __float2bfloat16(v);
    hi[i] = h;
    lo[i] = __float2bfloat16(v - __bfloat162float(h));
  }
}

int split_bf16_submit(adb_plan* plan, const float* x, void* hi, void* lo, size_t total, int silu_in, cudaStream_t stream) {
  ADB_REQUIRE(x && hi && lo && total > 0, "split_bf16: bad arguments");
  return submit(plan, stream, "split_bf16", 0.0, 0.0, [=](cudaStream_t s) -> int {
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(hi),
                                                       reinterpret_cast<__nv_bfloat16*>(lo), total, silu_in);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int ddim_step_submit(adb_plan* plan, const float* x, const float* model_out, int eps_channels,
                     const float* grad, float* x_prev, float* pred_xstart, int n, int c, int hw,
                     const float* coef, int clip_denoised, cudaStream_t stream) {
  ADB_REQUIRE(x && model_out && x_prev && coef, "ddim_step: null pointer");
  ADB_REQUIRE(n > 0 && c > 0 && hw > 0 && eps_channels >= c, "ddim_step: bad geometry");
  DdimCoef cf;
  for (int i = 0; i < 5; ++i) cf.v[i] = coef[i];
  return submit(plan, stream, "ddim_step", 0.0, 4.0 * (grad ? 4.0 : 3.0) * (double)n * c * hw, [=](cudaStream_t s) -> int {
    const size_t total = (size_t)n * c * hw;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    ddim_step_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, model_out, eps_channels, grad, x_prev,
                                                      pred_xstart, n, c, hw, cf, clip_denoised);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

int pack_uint8_submit(adb_plan* plan, const float* sample, uint8_t* out, int n, int c, int hw,
                      cudaStream_t stream) {
  ADB_REQUIRE(sample && out && n > 0 && c > 0 && hw > 0, "pack_uint8: bad arguments");
  // compulsory traffic: the fp32 sample read once, one byte per element written
  return submit(plan, stream, "pack_uint8", 0.0, 5.0 * (double)n * c * hw, [=](cudaStream_t s) -> int {
    const size_t total = (size_t)n * hw;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    pack_uint8_kernel<<<(unsigned)blocks, 256, 0, s>>>(sample, out, n, c, hw);
    ADB_CUDA(cudaGetLastError());
    return 1;
  });
}

}  // namespace adb
