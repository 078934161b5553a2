// attention_bwd_fused.cu — single-pass input-gradient of the fused softmax attention (head dim 64) on tcgen05 + TMEM.
//
// Same contract as attention_bwd.cu (the search's cond_fn differentiates through every attention block of the noisy
// classifier, search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-390; forward: guided_diffusion/unet.py
// QKVAttentionLegacy / QKVAttention, restated in attention2.cu). attention_bwd.cu rebuilds P = exp2(S - lse) and
// dS = P o (dP - D) twice - once per 128-query tile for dQ, once per 128-key tile for dK / dV - i.e. 7 tile products and
// 2 exponentials per score where 5 and 1 are needed, and (measured, profiles/r02_ncu_full_attention_gn_gemm_batch64_summary.csv)
// both kernels are bound by what every score costs the softmax warps: the TMEM read of S and dP (~64 B/clk/SM) plus the
// MUFU exponential, which share the SM's MIO path and add up. This kernel pays that once:
//
//   CTA = (128-key tile, batch*head), ONE CTA per SM, loop over 64-query steps i:
//     S^T_i = K Q_i^T, dP^T_i = V dO_i^T        (SS MMAs, 128 x 64 fp32 each, double-buffered in TMEM)
//     P^T = exp2(S^T sc - lse), dS^T = P^T o (dP^T - D) / 8      two softmax warpgroups, alternating steps; bf16 results
//                                                written back over S^T / dP^T (TMEM) and dS^T also to shared memory
//     dV += P^T dO_i, dK += dS^T Q_i            (TS MMAs: A from TMEM, B = the row-major dO_i / Q_i tiles, MN-major)
//     every two steps: dQ_pair = dS_pair K      (SS MMA, M = the pair's 128 queries: A = the staged dS^T, MN-major,
//                                                B = the resident K tile, MN-major) -> TMEM -> shared -> TMA reduction
//                                                (cp.reduce.async.bulk.tensor .add.f32) into the fp32 [b*t, C]
//                                                workspace: one partial per key tile, T/128 adds per element
//   TMEM (512 columns): S^T[2] | dP^T[2] | dV | dK | dQ = 7 x 64.
// A small second kernel rounds the fp32 dQ workspace into the bf16 dqkv matrix.
// tcgen05.mma ops of one thread execute in order, which is what makes the in-place P / dS aliasing and the reuse of a
// score buffer two steps later safe without extra barriers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int FB_THREADS = 320;  // warps 0-3: softmax group 0, 4-7: softmax group 1, 8: TMA, 9: MMA + TMEM
constexpr int BM = 128;          // keys per CTA
constexpr int HD = 64;
constexpr int TN = 64;           // queries per step
constexpr int BIG_BYTES = BM * HD * 2;    // 16 KiB: K / V
constexpr int SMALL_BYTES = TN * HD * 2;  // 8 KiB: Q_i / dO_i
constexpr int STAGES = 4;
constexpr int DS_BYTES = 2 * BM * TN * 2;  // 32 KiB: dS^T of a step pair, [2 query chunks][128 keys][64 queries]
constexpr int TMEM_COLS = 512;
constexpr int C_S = 0, C_DP = 64, C_BUF = 128;  // buffer g: S^T at g*128, dP^T at g*128 + 64
constexpr int C_DV = 256, C_DK = 320, C_DQ = 384;
constexpr int DQ_STAGE_BYTES = BM * HD * 4;  // 32 KiB per softmax group: a pair's fp32 dQ tile on its way to the TMA reduction
constexpr int SMEM_BYTES = 2 * BIG_BYTES + STAGES * 2 * SMALL_BYTES + DS_BYTES + 2 * DQ_STAGE_BYTES + 1024;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct FusedParams {
  CUtensorMap tmQKV128;  // qkv [b*t, 3C], box {64, 128}
  CUtensorMap tmQKV64;   // qkv, box {64, 64}
  CUtensorMap tmDO64;    // dO [b*t, C], box {64, 64}
  CUtensorMap tmDQ;      // dq_acc fp32 [b*t, C], box {32, 128}, 128-byte swizzle
  const float* lse;      // [b*heads, T]
  const float* dsum;     // [b*heads, T]
  __nv_bfloat16* dqkv;   // [b*t, 3C]: the k and v columns are written here
  float* dq_acc;         // [b*t, C] fp32, zeroed by the caller: dQ partials are added here
  int T, heads, C, legacy;
};

// barriers
constexpr int B_BIG = 0;
constexpr int B_FULL = 1;                  // [STAGES]
constexpr int B_EMPTY = B_FULL + STAGES;   // [STAGES]
constexpr int B_SP = B_EMPTY + STAGES;     // [2] score accumulators of buffer g ready
constexpr int B_PD = B_SP + 2;             // [2] P^T / dS^T of buffer g written (4 warps)
constexpr int B_DQ_FULL = B_PD + 2;        // dQ of a pair complete in TMEM (also: its staged dS^T has been consumed)
constexpr int B_DQ_EMPTY = B_DQ_FULL + 1;  // dQ read out of TMEM (4 warps)
constexpr int B_DONE = B_DQ_EMPTY + 1;
constexpr int NUM_BARS = B_DONE + 1;

__global__ void __launch_bounds__(FB_THREADS, 1) attn_bwd_fused_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NUM_BARS];
  __shared__ uint32_t tmem_slot_s;
  __shared__ __align__(16) float lse_s[STAGES][TN];
  __shared__ __align__(16) float dsum_s[STAGES][TN];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t k_smem = smem_base;
  const uint32_t v_smem = smem_base + BIG_BYTES;
  auto q_smem = [&](int st) { return smem_base + 2 * BIG_BYTES + st * 2 * SMALL_BYTES; };
  auto do_smem = [&](int st) { return q_smem(st) + SMALL_BYTES; };
  const uint32_t ds_smem = smem_base + 2 * BIG_BYTES + STAGES * 2 * SMALL_BYTES;
  uint8_t* ds_ptr = smem_raw + (ds_smem - smem_u32(smem_raw));
  const uint32_t dqst_smem = ds_smem + DS_BYTES;  // [2 groups][2 column halves][128 rows][32 fp32], 128-byte swizzle
  uint8_t* dqst_ptr = smem_raw + (dqst_smem - smem_u32(smem_raw));
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int k0 = blockIdx.x * BM;
  const int row_base = b * p.T;
  const int qc = p.legacy ? h * 3 * HD : h * HD;
  const int kc = p.legacy ? qc + HD : p.C + h * HD;
  const int vc = p.legacy ? qc + 2 * HD : 2 * p.C + h * HD;
  const int nt = p.T / TN;  // even: T is a multiple of 128

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQKV128);
    tma_prefetch_desc(&p.tmQKV64);
    tma_prefetch_desc(&p.tmDO64);
    tma_prefetch_desc(&p.tmDQ);
    for (int i = 0; i < NUM_BARS; ++i)
      mbar_init(bar(i), (i == B_PD || i == B_PD + 1 || i == B_DQ_EMPTY) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(smem_u32(&tmem_slot_s), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(B_BIG), 2 * BIG_BYTES);
      tma_load_2d(k_smem, &p.tmQKV128, bar(B_BIG), kc, row_base + k0);
      tma_load_2d(v_smem, &p.tmQKV128, bar(B_BIG), vc, row_base + k0);
      for (int i = 0; i < nt; ++i) {
        const int st = i % STAGES;
        mbar_wait(bar(B_EMPTY + st), ((uint32_t)(i / STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(B_FULL + st), 2 * SMALL_BYTES + 2 * TN * 4);
        tma_load_2d(q_smem(st), &p.tmQKV64, bar(B_FULL + st), qc, row_base + i * TN);
        tma_load_2d(do_smem(st), &p.tmDO64, bar(B_FULL + st), h * HD, row_base + i * TN);
        bulk_load_1d(smem_u32(&lse_s[st][0]), p.lse + (size_t)bh * p.T + i * TN, TN * 4, bar(B_FULL + st));
        bulk_load_1d(smem_u32(&dsum_s[st][0]), p.dsum + (size_t)bh * p.T + i * TN, TN * 4, bar(B_FULL + st));
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, TN, 0, 0);   // S^T / dP^T: A = K / V (K-major), B = Q_i / dO_i (K-major)
      constexpr uint32_t idesc_g = umma_idesc_bf16(BM, HD, 0, 1);   // dV / dK: A from TMEM, B = dO_i / Q_i rows: MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(BM, HD, 1, 1);   // dQ: A = staged dS^T (MN-major), B = K rows (MN-major)
      const uint64_t k_desc = umma_desc_kmajor_sw128(k_smem);
      const uint64_t v_desc = umma_desc_kmajor_sw128(v_smem);
      auto issue_scores = [&](int i) {
        const int st = i % STAGES, g = i & 1;
        mbar_wait(bar(B_FULL + st), (uint32_t)(i / STAGES) & 1u);
        tc_fence_after();
        const uint64_t q_desc = umma_desc_kmajor_sw128(q_smem(st));
        const uint64_t do_desc = umma_desc_kmajor_sw128(do_smem(st));
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16_ss(tmem_base + g * C_BUF + C_S, k_desc + 2u * kk, q_desc + 2u * kk, idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16_ss(tmem_base + g * C_BUF + C_DP, v_desc + 2u * kk, do_desc + 2u * kk, idesc_s, kk != 0);
        umma_commit(bar(B_SP + g));
      };
      mbar_wait(bar(B_BIG), 0);
      issue_scores(0);
      issue_scores(1);
      for (int i = 0; i < nt; ++i) {
        const int st = i % STAGES, g = i & 1;
        mbar_wait(bar(B_PD + g), (uint32_t)(i >> 1) & 1u);  // bf16 P^T / dS^T of step i in TMEM, dS^T also staged in smem
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < TN / 16; ++kk) {
          const uint64_t b_desc = umma_desc_mnmajor_sw128(do_smem(st) + kk * 2048, 1024);
          umma_bf16_ts(tmem_base + C_DV, tmem_base + g * C_BUF + C_S + 8 * kk, b_desc, idesc_g, (i | kk) != 0);  // dV += P^T dO_i
        }
#pragma unroll
        for (int kk = 0; kk < TN / 16; ++kk) {
          const uint64_t b_desc = umma_desc_mnmajor_sw128(q_smem(st) + kk * 2048, 1024);
          umma_bf16_ts(tmem_base + C_DK, tmem_base + g * C_BUF + C_DP + 8 * kk, b_desc, idesc_g, (i | kk) != 0);  // dK += dS^T Q_i
        }
        umma_commit(bar(B_EMPTY + st));
        if (i + 2 < nt) issue_scores(i + 2);  // overwrites buffer g: ordered after the two products above (in-order pipe)
        if (g == 1) {
          // the pair (i-1, i) is complete: dQ[128 queries, 64] = dS[128 q, 128 keys] K[128 keys, 64]
          const int pr = i >> 1;
          if (pr > 0) {
            mbar_wait(bar(B_DQ_EMPTY), (uint32_t)(pr - 1) & 1u);  // the previous pair's dQ has left TMEM
            tc_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BM / 16; ++kk) {
            const uint64_t a_desc = umma_desc_mnmajor_sw128(ds_smem + kk * 2048, BM * 128);  // 64-query chunks 16 KiB apart
            const uint64_t b_desc = umma_desc_mnmajor_sw128(k_smem + kk * 2048, 1024);
            umma_bf16_ss(tmem_base + C_DQ, a_desc, b_desc, idesc_q, kk != 0);
          }
          umma_commit(bar(B_DQ_FULL));
        }
      }
      umma_commit(bar(B_DONE));
    }
  } else {
    // ===================== softmax group g = warp / 4: steps i = g, g + 2, ... =====================
    const int g = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // key row of this thread (TMEM lane); also the query row when reading dQ
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t buf_addr = lane_addr + g * C_BUF;
    const float sc = 0.125f * 1.4426950408889634f;
    uint8_t* ds_row = ds_ptr + (size_t)g * (BM * 128) + (size_t)row * 128;  // this thread's 128-byte row of chunk g
    const int sw = row & 7;

    uint8_t* my_stage = dqst_ptr + (size_t)g * DQ_STAGE_BYTES;
    const uint32_t my_stage_u32 = dqst_smem + (uint32_t)g * DQ_STAGE_BYTES;
    const bool issuer = (quarter == 0 && lane == 0);  // the group's thread that owns its bulk async-group
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };

    auto read_dq = [&](int pr) {
      // dQ partial of pair pr: rows = its 128 queries, fp32. TMEM -> registers -> swizzled shared tile -> ONE TMA
      // reduction per 32-column half (cp.reduce.async.bulk.tensor .add): the adds reach L2 as whole lines. (Per-thread
      // red.global.add.v4 scatters 32 rows per warp instruction and was measured slower than the two-kernel form.)
      mbar_wait(bar(B_DQ_FULL), (uint32_t)pr & 1u);
      tc_fence_after();
      if (issuer) tma_store_wait_read0();  // this group's previous reduction has finished reading the staging tile
      group_sync();
#pragma unroll 1
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + C_DQ + c, v);
        tmem_wait_ld();
        if (c + 32 >= HD) {  // everything this warp needs is in registers: the accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_DQ_EMPTY));
        }
        uint8_t* srow = my_stage + (size_t)(c / 32) * (BM * 128) + (size_t)row * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(srow + ((u ^ sw) << 4)) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
      }
      fence_proxy_async_smem();
      group_sync();
      if (issuer) {
        const int r0 = row_base + pr * 2 * TN;
        tma_reduce_add_2d(&p.tmDQ, my_stage_u32, h * HD, r0);
        tma_reduce_add_2d(&p.tmDQ, my_stage_u32 + BM * 128, h * HD + 32, r0);
        tma_store_commit();
      }
    };

    for (int i = g; i < nt; i += 2) {
      const int st = i % STAGES;
      mbar_wait(bar(B_FULL + st), (uint32_t)(i / STAGES) & 1u);  // lse / D of step i visible to this thread
      mbar_wait(bar(B_SP + g), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      uint32_t pt[TN / 2], ds[TN / 2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t sv[32], dp[32];
        tmem_ld_32x32b_x32(buf_addr + C_S + hf * 32, sv);
        tmem_ld_32x32b_x32(buf_addr + C_DP + hf * 32, dp);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float2 l2 = *reinterpret_cast<const float2*>(&lse_s[st][hf * 32 + 2 * c]);
          const float2 d2 = *reinterpret_cast<const float2*>(&dsum_s[st][hf * 32 + 2 * c]);
          // packed pairs (FFMA2 / FADD2 / FMUL2), same roundings as the scalar form (the 1/8 is exact)
          float x0, x1, d0, d1;
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), f32x2_pack(sc, sc),
                                 f32x2_pack(-l2.x, -l2.y)), x0, x1);
          const float p0 = ex2_approx(x0);
          const float p1 = ex2_approx(x1);
          pt[hf * 16 + c] = pack_bf16x2(p0, p1);
          const uint64_t e2 = f32x2_add(f32x2_pack(__uint_as_float(dp[2 * c]), __uint_as_float(dp[2 * c + 1])),
                                        f32x2_pack(-d2.x, -d2.y));
          f32x2_unpack(f32x2_mul(f32x2_mul(f32x2_pack(p0, p1), f32x2_pack(0.125f, 0.125f)), e2), d0, d1);
          ds[hf * 16 + c] = pack_bf16x2(d0, d1);
        }
      }
      tmem_st_32x32b_x32(buf_addr + C_S, pt);
      tmem_st_32x32b_x32(buf_addr + C_DP, ds);
      // stage dS^T for the pair's dQ product: row = key, 64 queries = 128 bytes, 128-byte swizzle (16-byte unit u of row r
      // lives at unit u ^ (r & 7)). The previous pair's product must have consumed the buffer first.
      if (i >= 2) mbar_wait(bar(B_DQ_FULL), (uint32_t)((i >> 1) - 1) & 1u);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(ds_row + ((u ^ sw) << 4)) = make_uint4(ds[4 * u], ds[4 * u + 1], ds[4 * u + 2], ds[4 * u + 3]);
      fence_proxy_async_smem();
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PD + g));
      // read the dQ of an earlier pair out of TMEM (pairs alternate between the two groups); it was issued a full step
      // ago, so this does not wait
      if (i >= 2) {
        const int pr = (i >> 1) - 1;
        if ((pr & 1) == g) read_dq(pr);
      }
    }
    {
      const int last = nt / 2 - 1;  // pairs this group has not read yet: at most the last two
      const int lo = max(0, last - 1);
      for (int pr = lo; pr <= last; ++pr) {
        const int i_last = nt - 2 + g;  // this group's last step
        const bool done_in_loop = (i_last >= 2) && (pr <= (i_last >> 1) - 1);
        if ((pr & 1) == g && !done_in_loop) read_dq(pr);
      }
    }
    // epilogue: group 0 writes dV, group 1 writes dK
    mbar_wait(bar(B_DONE), 0);
    tc_fence_after();
    const bool ok = (k0 + row) < p.T;
    __nv_bfloat16* orow = p.dqkv + ((size_t)(row_base + k0 + row)) * (3 * p.C) + (g == 0 ? vc : kc);
    const uint32_t acc = lane_addr + (g == 0 ? C_DV : C_DK);
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(acc + c, v);
      tmem_wait_ld();
      if (ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          *reinterpret_cast<uint4*>(orow + c + q * 8) = o;
        }
      }
    }
  }

  if (warp < 8 && (warp & 3) == 0 && lane == 0) tma_store_wait_all0();  // the groups' reductions have left shared memory
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dq_acc fp32 [rows, C] -> bf16 into the q columns of dqkv [rows, 3C] (head h: C columns starting at qcol(h))
__global__ void __launch_bounds__(256) dq_round_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv,
                                                      size_t rows, int heads, int legacy) {
  const int C = heads * HD;
  const size_t total = rows * (size_t)(C / 8);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (C / 8);
    const int c = (int)(i - r * (C / 8)) * 8;
    const int hh = c / HD, d = c - hh * HD;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(acc + r * C + c));
    const float4 bq = __ldcs(reinterpret_cast<const float4*>(acc + r * C + c + 4));
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(bq.x, bq.y);
    o.w = pack_bf16x2(bq.z, bq.w);
    const int col = (legacy ? hh * 3 * HD : hh * HD) + d;
    *reinterpret_cast<uint4*>(dqkv + r * (size_t)(3 * C) + col) = o;
  }
}

}  // namespace

// launches: memset(dq_acc), fused kernel, rounding kernel. Returns the number of launches (kernels + memset) or < 0.
int attention_backward_fused_launch(const void* qkv, const void* dout, const float* lse, const float* dsum, void* dqkv,
                                    float* dq_acc, int b, int t, int heads, int legacy_order, cudaStream_t s,
                                    const CUtensorMap* tm_qkv128, const CUtensorMap* tm_qkv64, const CUtensorMap* tm_do64) {
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  FusedParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.tmQKV128 = *tm_qkv128;
  fp.tmQKV64 = *tm_qkv64;
  fp.tmDO64 = *tm_do64;
  fp.lse = lse;
  fp.dsum = dsum;
  fp.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  fp.dq_acc = dq_acc;
  fp.T = t;
  fp.heads = heads;
  fp.C = heads * HD;
  fp.legacy = legacy_order ? 1 : 0;
  const size_t rows = (size_t)b * t;
  {
    const uint64_t dims[2] = {(uint64_t)fp.C, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)fp.C * 4};
    const uint32_t box[2] = {32, 128};
    int r = make_tmap_f32(&fp.tmDQ, dq_acc, 2, dims, strides, box, 128);
    if (r != ADB_OK) return r;
  }
  ADB_CUDA(cudaMemsetAsync(dq_acc, 0, rows * fp.C * sizeof(float), s));
  dim3 grid(t / BM, b * heads);
  attn_bwd_fused_kernel<<<grid, FB_THREADS, SMEM_BYTES, s>>>(fp);
  ADB_CUDA(cudaGetLastError());
  const size_t vecs = rows * (fp.C / 8);
  size_t blocks = (vecs + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dq_round_kernel<<<(unsigned)blocks, 256, 0, s>>>(dq_acc, fp.dqkv, rows, heads, fp.legacy);
  ADB_CUDA(cudaGetLastError());
  return 3;
}

}  // namespace adb
