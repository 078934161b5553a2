// attention_sd.cu — fused softmax attention for the Stable-Diffusion transformer blocks: separate Q and
// K/V matrices (self- and cross-attention), head dims 40 / 80 / 160 stored in 64-column chunks, key masking.
//
// Replaces CrossAttention.forward between its projections (reference: "Stable Diffusion"/ldm/modules/attention.py:
// 170-194): out = softmax(q k^T * dim_head^-0.5) v per (batch, head), heads laid out as (h d) along channels.
//
// Layout contract (chosen by the weight packer, autodiffusion_b200/sd_unet.py): every head occupies DP = 64*NC
// columns of the projection outputs; columns [d, DP) are exactly zero (zero weight rows, the projections have no
// bias), so they add nothing to q.k and produce zeros in the output, whose consumer (to_out) has zero weight
// columns there. The kernel therefore needs no per-d variants: QK^T runs ceil(d/16) 16-deep k-steps, PV writes
// 64-column chunks (the last one n_last <= 64 wide).
//
// Same skeleton as attention2.cu (P kept in TMEM over S, 64-key tiles, double-buffered S, lazy rescaling); see
// that file for the barrier protocol. What differs: NC chunks of Q/K/V per tile, the Q and K/V tensor maps are
// different matrices with different rows per batch element, keys >= tk_valid are masked to -inf (77 context
// tokens inside a 128-row padded context), scale is a parameter.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int AT_THREADS = 192;
constexpr int BM = 128;
constexpr int KT = 64;
constexpr int CH = 64;                    // columns per head-dim chunk = one 128-byte swizzle row
constexpr int Q_CHUNK_BYTES = BM * CH * 2;   // 16 KiB
constexpr int KV_CHUNK_BYTES = KT * CH * 2;  // 8 KiB
constexpr int O_COL = 128;

__device__ __forceinline__ float ex2_approx_sd(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnSdParams {
  CUtensorMap tmQ;   // box {64, 128} over the Q matrix [b*tq, q_width]
  CUtensorMap tmKV;  // box {64, 64} over the K/V matrix [b*tk_rows, kv_width]
  __nv_bfloat16* out;  // [b*tq, heads*DP]
  int tq;        // queries per batch element
  int tk_rows;   // K/V rows per batch element in the K/V matrix (>= tk_valid)
  int tk_valid;  // keys that take part in the softmax
  int heads;
  int q_col0, k_col0, v_col0;  // column of head 0 in the Q / K / V matrices (head h is DP*h further)
  int ksteps;    // 16-deep k-steps of q.k: ceil(d/16)
  int n_last;    // width of the last PV chunk (multiple of 16)
  int d_head;    // real head dim: output columns [d_head, DP) are written as zeros
  float sc;      // d^-0.5 * log2(e)
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));  // one FMNMX3 on sm_100
  return d;
}

// 2^x on the FMA pipe (Cody-Waite split + degree-3 polynomial, relative error 1.1e-4 - below the 2^-9 rounding of the
// bf16 P it feeds): every 4th probability takes this path so the 16-lane MUFU unit is not the only exp engine.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);  // [-0.5, 0.5]
  float q = fmaf(f, 0.05459282f, 0.24221784f);
  q = fmaf(q, f, 0.6933686f);
  q = fmaf(q, f, 1.0f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

template <int NC>
struct SdCfg {
  static constexpr int STAGES = (NC == 1) ? 4 : (NC == 2 ? 2 : 3);
  static constexpr int Q_BYTES = NC * Q_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = 2 * NC * KV_CHUNK_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (O_COL + NC * CH <= 256) ? 256 : 512;
  static constexpr int MIN_CTAS = (NC <= 2) ? 2 : 1;  // NC = 2: 32 KB Q + 2 x 32 KB stages -> two CTAs per SM
};

// ONES: column d_head of V is 1.0 in every key row (the packer puts it there through the V projection's bias), so
// the PV product accumulates the softmax denominator in O[:, d_head] - the 64 row-sum FADDs per thread and tile
// disappear from the exp-bound softmax loop, and the denominator is consistent with the bf16-rounded P.
template <int NC, bool ONES, bool POLY>
__global__ void __launch_bounds__(AT_THREADS, SdCfg<NC>::MIN_CTAS) attention_sd_kernel(const __grid_constant__ AttnSdParams p) {
  using C = SdCfg<NC>;
  constexpr int STAGES = C::STAGES;
  constexpr int B_Q = 0;
  constexpr int B_KV_FULL = 1;
  constexpr int B_KV_EMPTY = B_KV_FULL + STAGES;
  constexpr int B_S_FULL = B_KV_EMPTY + STAGES;
  constexpr int B_P_FULL = B_S_FULL + 2;
  constexpr int B_PV_DONE = B_P_FULL + 2;
  constexpr int NUM_BARS = B_PV_DONE + 2;
  constexpr int DP = NC * CH;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NUM_BARS];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int st) { return smem_base + C::Q_BYTES + st * C::STAGE_BYTES; };
  auto v_smem = [&](int st) { return k_smem(st) + NC * KV_CHUNK_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * BM;
  const int qrow_base = b * p.tq;
  const int krow_base = b * p.tk_rows;
  const int qc = p.q_col0 + h * DP;
  const int kc = p.k_col0 + h * DP;
  const int vc = p.v_col0 + h * DP;
  const int nkt = (p.tk_valid + KT - 1) / KT;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    for (int i = 0; i < NUM_BARS; ++i) mbar_init(bar(i), (i == B_P_FULL || i == B_P_FULL + 1) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_slot_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(B_Q), C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < NC; ++c) tma_load_2d(q_smem + c * Q_CHUNK_BYTES, &p.tmQ, bar(B_Q), qc + c * CH, qrow_base + q0);
      for (int j = 0; j < nkt; ++j) {
        const int st = j % STAGES;
        const uint32_t use = (uint32_t)(j / STAGES);
        mbar_wait(bar(B_KV_EMPTY + st), (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(B_KV_FULL + st), C::STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(k_smem(st) + c * KV_CHUNK_BYTES, &p.tmKV, bar(B_KV_FULL + st), kc + c * CH, krow_base + j * KT);
          tma_load_2d(v_smem(st) + c * KV_CHUNK_BYTES, &p.tmKV, bar(B_KV_FULL + st), vc + c * CH, krow_base + j * KT);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, KT, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, CH, 0, 1);  // B = V, MN-major (as stored)
      const uint32_t idesc_o_last = umma_idesc_bf16(BM, p.n_last, 0, 1);
      auto issue_s = [&](int j) {
        const int st = j % STAGES;
        mbar_wait(bar(B_KV_FULL + st), (uint32_t)(j / STAGES) & 1u);
        tc_fence_after();
        for (int kk = 0; kk < p.ksteps; ++kk) {
          const int c = kk >> 2, w = kk & 3;
          const uint64_t a_desc = umma_desc_kmajor_sw128(q_smem + c * Q_CHUNK_BYTES) + 2u * w;
          const uint64_t b_desc = umma_desc_kmajor_sw128(k_smem(st) + c * KV_CHUNK_BYTES) + 2u * w;
          umma_bf16_ss(tmem_base + (j & 1) * KT, a_desc, b_desc, idesc_s, kk != 0);
        }
        umma_commit(bar(B_S_FULL + (j & 1)));
      };
      mbar_wait(bar(B_Q), 0);
      issue_s(0);
      for (int j = 0; j < nkt; ++j) {
        if (j + 1 < nkt) issue_s(j + 1);
        const int st = j % STAGES;
        mbar_wait(bar(B_P_FULL + (j & 1)), (uint32_t)(j >> 1) & 1u);  // P_j in TMEM (and O rescaled)
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint32_t idesc = (c == NC - 1) ? idesc_o_last : idesc_o;
#pragma unroll
          for (int kk = 0; kk < KT / 16; ++kk) {
            const uint64_t b_desc = umma_desc_mnmajor_sw128(v_smem(st) + c * KV_CHUNK_BYTES + kk * 2048, 1024);
            umma_bf16_ts(tmem_base + O_COL + c * CH, tmem_base + (j & 1) * KT + 8 * kk, b_desc, idesc, (j | kk) != 0);
          }
        }
        umma_commit(bar(B_KV_EMPTY + st));
        umma_commit(bar(B_PV_DONE + (j & 1)));
      }
    }
  } else {
    // ===================== softmax (warps 0..3): one query row per thread =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = p.sc;
    const int ocols = (NC - 1) * CH + p.n_last;  // O columns the PV product writes
    float m_run = -INFINITY;
    float l_run = 0.f;
    for (int j = 0; j < nkt; ++j) {
      const uint32_t s_addr = lane_addr + (j & 1) * KT;
      mbar_wait(bar(B_S_FULL + (j & 1)), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t sr[KT];
      tmem_ld_32x32b_x32(s_addr, sr);
      tmem_ld_32x32b_x32(s_addr + 32, sr + 32);
      tmem_wait_ld();
      const int kvalid = p.tk_valid - j * KT;  // keys of this tile inside the sequence (warp-uniform)
      if (kvalid < KT) {
#pragma unroll
        for (int i = 0; i < KT; ++i)
          if (i >= kvalid) sr[i] = 0xff800000u;  // -inf
      }
      // 8 independent chains of 3-input maxima: 64 values in 28 + 4 FMNMX3
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = fmax3(__uint_as_float(sr[i]), __uint_as_float(sr[8 + i]), __uint_as_float(sr[16 + i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[24 + i]), __uint_as_float(sr[32 + i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[40 + i]), __uint_as_float(sr[48 + i]));
#pragma unroll
      for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], mxs[4 + i], __uint_as_float(sr[56 + i]));
#pragma unroll
      for (int i = 0; i < 4; ++i) mxs[i] = fmaxf(mxs[i], __uint_as_float(sr[60 + i]));
      const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      // sc > 0: max and scaling commute. Lazy rescaling as in attention2.cu (threshold 8 in log2 units).
      const float m_tile = mx * sc;
      const bool jump = m_tile > m_run + 8.0f;
      float alpha = 1.0f;
      float m_new = m_run;
      if (__any_sync(0xffffffffu, jump)) {
        m_new = fmaxf(m_run, m_tile);
        alpha = ex2_approx_sd(m_run - m_new);
        if (j > 0) {
          mbar_wait(bar(B_PV_DONE + ((j - 1) & 1)), (uint32_t)((j - 1) >> 1) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DP; c += 32) {
            if (c >= ocols) break;
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + O_COL + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(lane_addr + O_COL + c, v);
          }
        }
      }
      if (ONES) {
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) {
          float x0, x1;
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), f32x2_pack(sc, sc),
                                 f32x2_pack(-m_new, -m_new)), x0, x1);
          const float p0 = ex2_approx_sd(x0);
          const float p1 = (POLY && (i & 1)) ? ex2_poly(x1) : ex2_approx_sd(x1);
          sr[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x32(s_addr, sr);
      } else {
        // packed pairs (FFMA2 / FADD2), same roundings and summation order as the scalar form
        const uint64_t sc2 = f32x2_pack(sc, sc), nm2 = f32x2_pack(-m_new, -m_new);
        uint64_t ps2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ps2[i] = f32x2_pack(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) {
          float x0, x1;
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sc2, nm2), x0, x1);
          const float p0 = ex2_approx_sd(x0);
          const float p1 = ex2_approx_sd(x1);
          ps2[i & 3] = f32x2_add(ps2[i & 3], f32x2_pack(p0, p1));
          sr[i] = pack_bf16x2(p0, p1);
        }
        float ps[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) f32x2_unpack(ps2[i], ps[2 * i], ps[2 * i + 1]);
        const float ps0 = (ps[0] + ps[1]) + (ps[2] + ps[3]);
        const float ps1 = (ps[4] + ps[5]) + (ps[6] + ps[7]);
        tmem_st_32x32b_x32(s_addr, sr);
        l_run = l_run * alpha + (ps0 + ps1);
      }
      m_run = m_new;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_P_FULL + (j & 1)));
    }
    mbar_wait(bar(B_PV_DONE + ((nkt - 1) & 1)), (uint32_t)((nkt - 1) >> 1) & 1u);
    tc_fence_after();
    if (ONES) {  // the denominator accumulated by the tensor core in O[:, d_head]
      uint32_t lv[16];
      tmem_ld_32x32b_x16(lane_addr + O_COL + (p.d_head & ~15), lv);
      tmem_wait_ld();
      l_run = __uint_as_float(lv[0]);
#pragma unroll
      for (int i = 1; i < 16; ++i)
        if (i == (p.d_head & 15)) l_run = __uint_as_float(lv[i]);
    }
    const float inv = 1.0f / l_run;
    const bool ok = (q0 + row) < p.tq;
    __nv_bfloat16* orow = p.out + ((size_t)(qrow_base + q0 + row)) * ((size_t)p.heads * DP) + h * DP;
#pragma unroll 1
    for (int c = 0; c < DP; c += 32) {
      uint32_t v[32];
      if (c < ocols) {
        tmem_ld_32x32b_x32(lane_addr + O_COL + c, v);
        tmem_wait_ld();
      }
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c + i >= p.d_head) v[i] = 0u;  // the head's padding columns (never written, or the ones column)
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Experiment (off by default, ADB_ATTN_SD_V2=1; parity-tested): TWO softmax warp groups per CTA, each owning every other key tile with its own S buffer and
// its own output accumulator (split-KV inside the CTA). Measured on the kernel above at T = 4096, d = 40: XU pipe
// 48 % busy, issue slots 55 %, top stalls fixed-latency waits and TMEM round trips - with 2 softmax warps per
// scheduler the exp pipe starves on latency, not on throughput. Here
//   * warps 0-3 take tiles 0, 2, 4, ... (S0 / P0, accumulator O0), warps 6-9 tiles 1, 3, 5, ... (S1 / P1, O1): the
//     groups never exchange anything per tile - each keeps its own running maximum - and the two partial results
//     are merged once at the end: out = (f0 O0 + f1 O1) / (f0 l0 + f1 l1), f_g = 2^(m_g - max(m0, m1));
//   * S is read from TMEM twice (maximum pass, exponential pass) in 32-column halves, so a softmax thread holds 32
//     scores instead of 64: 320 threads x 2 CTAs fit the register file, 4 softmax warps per scheduler.
// TMEM columns: S0 [0,64) | S1 [64,128) | O0 [128, 128+DP) | O1 [128+DP, 128+2DP). Barrier protocol as above (the
// per-buffer P_FULL / PV_DONE barriers were already indexed by tile parity = group).
template <int NC>
struct Sd2Cfg {
  static constexpr int STAGES = (NC == 1) ? 4 : 3;
  static constexpr int Q_BYTES = NC * Q_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = 2 * NC * KV_CHUNK_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (O_COL + 2 * NC * CH <= 256) ? 256 : 512;
  static constexpr int MIN_CTAS = (NC == 1) ? 2 : 1;
};
constexpr int AT2_THREADS = 320;

template <int NC, bool ONES>
__global__ void __launch_bounds__(AT2_THREADS, Sd2Cfg<NC>::MIN_CTAS) attention_sd2_kernel(const __grid_constant__ AttnSdParams p) {
  using C = Sd2Cfg<NC>;
  constexpr int STAGES = C::STAGES;
  constexpr int B_Q = 0;
  constexpr int B_KV_FULL = 1;
  constexpr int B_KV_EMPTY = B_KV_FULL + STAGES;
  constexpr int B_S_FULL = B_KV_EMPTY + STAGES;
  constexpr int B_P_FULL = B_S_FULL + 2;
  constexpr int B_PV_DONE = B_P_FULL + 2;
  constexpr int B_MERGE = B_PV_DONE + 2;
  constexpr int NUM_BARS = B_MERGE + 1;
  constexpr int DP = NC * CH;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NUM_BARS];
  __shared__ uint32_t tmem_slot_s;
  __shared__ float merge_m[BM], merge_l[BM];  // group 1's running maximum / denominator per row
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int st) { return smem_base + C::Q_BYTES + st * C::STAGE_BYTES; };
  auto v_smem = [&](int st) { return k_smem(st) + NC * KV_CHUNK_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * BM;
  const int qrow_base = b * p.tq;
  const int krow_base = b * p.tk_rows;
  const int qc = p.q_col0 + h * DP;
  const int kc = p.k_col0 + h * DP;
  const int vc = p.v_col0 + h * DP;
  const int nkt = (p.tk_valid + KT - 1) / KT;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    for (int i = 0; i < NUM_BARS; ++i)
      mbar_init(bar(i), (i == B_P_FULL || i == B_P_FULL + 1 || i == B_MERGE) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_slot_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(B_Q), C::Q_BYTES);
#pragma unroll
      for (int c = 0; c < NC; ++c) tma_load_2d(q_smem + c * Q_CHUNK_BYTES, &p.tmQ, bar(B_Q), qc + c * CH, qrow_base + q0);
      for (int j = 0; j < nkt; ++j) {
        const int st = j % STAGES;
        const uint32_t use = (uint32_t)(j / STAGES);
        mbar_wait(bar(B_KV_EMPTY + st), (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(B_KV_FULL + st), C::STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(k_smem(st) + c * KV_CHUNK_BYTES, &p.tmKV, bar(B_KV_FULL + st), kc + c * CH, krow_base + j * KT);
          tma_load_2d(v_smem(st) + c * KV_CHUNK_BYTES, &p.tmKV, bar(B_KV_FULL + st), vc + c * CH, krow_base + j * KT);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, KT, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, CH, 0, 1);
      const uint32_t idesc_o_last = umma_idesc_bf16(BM, p.n_last, 0, 1);
      auto issue_s = [&](int j) {
        const int st = j % STAGES;
        mbar_wait(bar(B_KV_FULL + st), (uint32_t)(j / STAGES) & 1u);
        tc_fence_after();
        for (int kk = 0; kk < p.ksteps; ++kk) {
          const int c = kk >> 2, w = kk & 3;
          const uint64_t a_desc = umma_desc_kmajor_sw128(q_smem + c * Q_CHUNK_BYTES) + 2u * w;
          const uint64_t b_desc = umma_desc_kmajor_sw128(k_smem(st) + c * KV_CHUNK_BYTES) + 2u * w;
          umma_bf16_ss(tmem_base + (j & 1) * KT, a_desc, b_desc, idesc_s, kk != 0);
        }
        umma_commit(bar(B_S_FULL + (j & 1)));
      };
      mbar_wait(bar(B_Q), 0);
      issue_s(0);
      for (int j = 0; j < nkt; ++j) {
        if (j + 1 < nkt) issue_s(j + 1);  // S buffer (j+1)&1 last held P_{j-1}, consumed by PV_{j-1} issued one iteration ago
        const int st = j % STAGES;
        mbar_wait(bar(B_P_FULL + (j & 1)), (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        const uint32_t o_base = tmem_base + O_COL + (j & 1) * DP;  // the group's own accumulator
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint32_t idesc = (c == NC - 1) ? idesc_o_last : idesc_o;
#pragma unroll
          for (int kk = 0; kk < KT / 16; ++kk) {
            const uint64_t b_desc = umma_desc_mnmajor_sw128(v_smem(st) + c * KV_CHUNK_BYTES + kk * 2048, 1024);
            umma_bf16_ts(o_base + c * CH, tmem_base + (j & 1) * KT + 8 * kk, b_desc, idesc, ((j >> 1) | kk) != 0);
          }
        }
        umma_commit(bar(B_KV_EMPTY + st));
        umma_commit(bar(B_PV_DONE + (j & 1)));
      }
    }
  } else {
    // ===================== softmax groups: warps 0-3 (even tiles), warps 6-9 (odd tiles) =====================
    const int grp = warp >= 6 ? 1 : 0;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + grp * KT;
    const uint32_t o_addr = lane_addr + O_COL + grp * DP;
    const float sc = p.sc;
    const int ocols = (NC - 1) * CH + p.n_last;
    float m_run = -INFINITY;
    float l_run = 0.f;
    int j_last = -1;
    for (int j = grp; j < nkt; j += 2) {
      j_last = j;
      mbar_wait(bar(B_S_FULL + grp), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const int kvalid = p.tk_valid - j * KT;  // warp-uniform
      // ---- pass 1: row maximum, 32 columns at a time ----
      float mx;
      {
        uint32_t sr[32];
        float mxs[4];
        tmem_ld_32x32b_x32(s_addr, sr);
        tmem_wait_ld();
        if (kvalid < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= kvalid) sr[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(__uint_as_float(sr[i]), __uint_as_float(sr[4 + i]), __uint_as_float(sr[8 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[12 + i]), __uint_as_float(sr[16 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[20 + i]), __uint_as_float(sr[24 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmaxf(mxs[i], __uint_as_float(sr[28 + i]));
        tmem_ld_32x32b_x32(s_addr + 32, sr);
        tmem_wait_ld();
        if (kvalid < KT) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 + i >= kvalid) sr[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[i]), __uint_as_float(sr[4 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[8 + i]), __uint_as_float(sr[12 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[16 + i]), __uint_as_float(sr[20 + i]));
#pragma unroll
        for (int i = 0; i < 4; ++i) mxs[i] = fmax3(mxs[i], __uint_as_float(sr[24 + i]), __uint_as_float(sr[28 + i]));
        mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      }
      const float m_tile = mx * sc;
      const bool jump = m_tile > m_run + 8.0f;
      float alpha = 1.0f;
      float m_new = m_run;
      if (__any_sync(0xffffffffu, jump)) {
        m_new = fmaxf(m_run, m_tile);
        alpha = ex2_approx_sd(m_run - m_new);
        if (j >= 2) {
          // this group's previous product PV_{j-2}: the latest commit on its barrier (PV_j needs the P written below)
          mbar_wait(bar(B_PV_DONE + grp), (uint32_t)((j - 2) >> 1) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DP; c += 32) {
            if (c >= ocols) break;
            uint32_t v[32];
            tmem_ld_32x32b_x32(o_addr + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(o_addr + c, v);
          }
        }
      }
      // ---- pass 2: probabilities, 32 keys at a time; bf16 P for keys [32 hf, +32) goes to columns [16 hf, +16) of the
      // S buffer - below column 32, so the second half of S is still intact when it is read ----
      float ps = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(s_addr + 32 * hf, sr);
        tmem_wait_ld();
        if (kvalid < KT) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 * hf + i >= kvalid) sr[i] = 0xff800000u;
        }
        uint32_t pk[16];
        float pa = 0.f, pb = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = ex2_approx_sd(fmaf(__uint_as_float(sr[2 * i]), sc, -m_new));
          const float p1 = ex2_approx_sd(fmaf(__uint_as_float(sr[2 * i + 1]), sc, -m_new));
          if (!ONES) {
            pa += p0;
            pb += p1;
          }
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(s_addr + 16 * hf, pk);
        ps += pa + pb;
      }
      if (!ONES) l_run = l_run * alpha + ps;
      m_run = m_new;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_P_FULL + grp));
    }
    if (j_last >= 0) {
      mbar_wait(bar(B_PV_DONE + grp), (uint32_t)(j_last >> 1) & 1u);  // this group's accumulator is final
      tc_fence_after();
    }
    auto read_l = [&](uint32_t oa) -> float {  // ONES: the denominator the tensor core accumulated in O[:, d_head]
      uint32_t lv[16];
      tmem_ld_32x32b_x16(oa + (p.d_head & ~15), lv);
      tmem_wait_ld();
      float l = __uint_as_float(lv[0]);
#pragma unroll
      for (int i = 1; i < 16; ++i)
        if (i == (p.d_head & 15)) l = __uint_as_float(lv[i]);
      return l;
    };
    if (grp == 1) {
      // publish (m1, l1) and leave; group 0 merges. With a single key tile there is nothing here: m1 = -inf.
      if (ONES && j_last >= 0) l_run = read_l(o_addr);
      merge_m[row] = m_run;
      merge_l[row] = l_run;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_MERGE));
    } else {
      if (ONES) l_run = read_l(o_addr);
      mbar_wait(bar(B_MERGE), 0);
      const float m1 = merge_m[row];
      const float l1 = merge_l[row];
      const bool has1 = nkt >= 2;  // uniform: O1 was written at all
      const float m = fmaxf(m_run, m1);
      const float f0 = ex2_approx_sd(m_run - m);
      const float f1 = has1 ? ex2_approx_sd(m1 - m) : 0.f;
      const float inv = 1.0f / (f0 * l_run + (has1 ? f1 * l1 : 0.f));
      const float w0 = f0 * inv, w1 = f1 * inv;
      const bool ok = (q0 + row) < p.tq;
      __nv_bfloat16* orow = p.out + ((size_t)(qrow_base + q0 + row)) * ((size_t)p.heads * DP) + h * DP;
#pragma unroll 1
      for (int c = 0; c < DP; c += 16) {  // 16 columns at a time: the merge must fit the 96-register budget
        uint32_t v[16], u[16];
        if (c < ocols) {
          tmem_ld_32x32b_x16(o_addr + c, v);
          if (has1) tmem_ld_32x32b_x16(o_addr + DP + c, u);
          tmem_wait_ld();
        }
        uint32_t ow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a0 = 0.f, a1 = 0.f;
          if (c + 2 * i < p.d_head) {
            a0 = __uint_as_float(v[2 * i]) * w0;
            if (has1) a0 = fmaf(__uint_as_float(u[2 * i]), w1, a0);
          }
          if (c + 2 * i + 1 < p.d_head) {
            a1 = __uint_as_float(v[2 * i + 1]) * w0;
            if (has1) a1 = fmaf(__uint_as_float(u[2 * i + 1]), w1, a1);
          }
          ow[i] = pack_bf16x2(a0, a1);
        }
        if (ok) {
          *reinterpret_cast<uint4*>(orow + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int NC, bool ONES>
int launch_attn_sd2(const AttnSdParams& ap, int b, cudaStream_t stream) {
  using C = Sd2Cfg<NC>;
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attention_sd2_kernel<NC, ONES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((ap.tq + BM - 1) / BM, b * ap.heads);
  attention_sd2_kernel<NC, ONES><<<grid, AT2_THREADS, C::SMEM_BYTES, stream>>>(ap);
  ADB_CUDA(cudaGetLastError());
  return 1;
}

template <int NC, bool ONES, bool POLY>
int launch_attn_sd(const AttnSdParams& ap, int b, cudaStream_t stream) {
  using C = SdCfg<NC>;
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attention_sd_kernel<NC, ONES, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((ap.tq + BM - 1) / BM, b * ap.heads);
  attention_sd_kernel<NC, ONES, POLY><<<grid, AT_THREADS, C::SMEM_BYTES, stream>>>(ap);
  ADB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace

int attention_sd_submit(adb_plan* plan, const adb_attn_sd_desc* d, cudaStream_t stream) {
  ADB_REQUIRE(d && d->q && d->kv && d->out && d->b > 0 && d->heads > 0, "attention_sd: bad arguments");
  ADB_REQUIRE(d->d_head > 0 && d->d_head <= 192 && d->d_pad % 64 == 0 && d->d_pad >= d->d_head && d->d_pad <= 192,
              "attention_sd: head dim %d (padded %d) unsupported", d->d_head, d->d_pad);
  ADB_REQUIRE(d->tq > 0 && d->tk_valid > 0 && d->tk_rows >= d->tk_valid, "attention_sd: bad sequence lengths");
  ADB_REQUIRE(d->q_width % 8 == 0 && d->kv_width % 8 == 0, "attention_sd: matrix widths must be multiples of 8");
  const int nc = d->d_pad / 64;
  AttnSdParams ap;
  memset(&ap, 0, sizeof(ap));
  {
    const uint64_t dims[2] = {(uint64_t)d->q_width, (uint64_t)d->b * d->tq};
    const uint64_t strides[1] = {(uint64_t)d->q_width * 2};
    const uint32_t box[2] = {64, 128};
    int r = make_tmap_bf16(&ap.tmQ, d->q, 2, dims, strides, box);
    if (r != ADB_OK) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->kv_width, (uint64_t)d->b * d->tk_rows};
    const uint64_t strides[1] = {(uint64_t)d->kv_width * 2};
    const uint32_t box[2] = {64, 64};
    int r = make_tmap_bf16(&ap.tmKV, d->kv, 2, dims, strides, box);
    if (r != ADB_OK) return r;
  }
  ap.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  ap.tq = d->tq;
  ap.tk_rows = d->tk_rows;
  ap.tk_valid = d->tk_valid;
  ap.heads = d->heads;
  ap.q_col0 = d->q_col0;
  ap.k_col0 = d->k_col0;
  ap.v_col0 = d->v_col0;
  ap.ksteps = (d->d_head + 15) / 16;
  static int full_n = -1;
  if (full_n < 0) {
    const char* e = getenv("ADB_ATTN_SD_FULLN");
    full_n = (e && e[0] == '1') ? 1 : 0;
  }
  const int ones = d->v_ones ? 1 : 0;
  ADB_REQUIRE(!ones || d->d_pad > d->d_head, "attention_sd: v_ones needs a padding column (d_pad > d_head)");
  const int rem = d->d_head + ones - (nc - 1) * 64;    // columns of the last chunk the PV product must cover
  ap.n_last = full_n ? 64 : ((rem + 15) / 16) * 16;
  ap.d_head = d->d_head;
  ap.sc = (float)(1.4426950408889634 / sqrt((double)d->d_head));
  const double flops = 4.0 * (double)d->b * d->heads * (double)d->tq * (double)d->tk_valid * d->d_head;
  const int b = d->b;
  // ADB_ATTN_POLY=1: a quarter of the exponentials of long key sequences go to the FMA pipe. Off by default: measured
  // at T = 4096, d = 40 the kernel is latency-bound (XU pipe 48 % busy, issue slots 55 %), and the extra FMA-pipe
  // instructions made it 8 % slower (3.73 vs 3.37 ms at batch 64).
  static int poly_on = -1;
  if (poly_on < 0) {
    const char* e = getenv("ADB_ATTN_POLY");
    poly_on = (e && e[0] == '1') ? 1 : 0;
  }
  const int poly = (ones && poly_on && d->tk_valid >= 1024) ? 1 : 0;
  // ADB_ATTN_SD_V2=1 selects the two-group kernel. Measured (batch 64, T = 4096, d = 40): 3.52 ms vs 3.37-3.59 ms for the
  // single-group kernel - both sit at 64 % of the XU (MUFU.EX2) pipe's peak with the SMs fully active and L2 at 20 %;
  // doubling the softmax warps per scheduler did not move it, so the single-group kernel stays the default.
  static int gen2 = -1;
  if (gen2 < 0) {
    const char* e = getenv("ADB_ATTN_SD_V2");
    gen2 = (e && e[0] == '1') ? 1 : 0;
  }
  if (gen2 && !poly) {
    return submit(plan, stream, "attention_sd", flops, 0.0, [ap, b, nc, ones](cudaStream_t s) -> int {
      switch (nc * 2 + ones) {
        case 2: return launch_attn_sd2<1, false>(ap, b, s);
        case 3: return launch_attn_sd2<1, true>(ap, b, s);
        case 4: return launch_attn_sd2<2, false>(ap, b, s);
        case 5: return launch_attn_sd2<2, true>(ap, b, s);
        case 6: return launch_attn_sd2<3, false>(ap, b, s);
        default: return launch_attn_sd2<3, true>(ap, b, s);
      }
    });
  }
  return submit(plan, stream, "attention_sd", flops, 0.0, [ap, b, nc, ones, poly](cudaStream_t s) -> int {
    switch (nc * 4 + ones * 2 + poly) {
      case 4: return launch_attn_sd<1, false, false>(ap, b, s);
      case 6: return launch_attn_sd<1, true, false>(ap, b, s);
      case 7: return launch_attn_sd<1, true, true>(ap, b, s);
      case 8: return launch_attn_sd<2, false, false>(ap, b, s);
      case 10: return launch_attn_sd<2, true, false>(ap, b, s);
      case 11: return launch_attn_sd<2, true, true>(ap, b, s);
      case 12: return launch_attn_sd<3, false, false>(ap, b, s);
      default: return launch_attn_sd<3, true, false>(ap, b, s);
    }
  });
}

}  // namespace adb
