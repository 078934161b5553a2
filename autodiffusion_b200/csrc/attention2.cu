// attention2.cu — fused softmax attention, head dim 64, second generation.
//
// Same contract as attention.cu (reference: QKVAttention / QKVAttentionLegacy,
// guided_diffusion/dynamic_unet.py:390-409 / 357-374). What changed, and why:
// the first kernel needed 512 TMEM columns and 112 KB of shared memory, i.e. ONE CTA per SM whose
// 128 softmax threads (exp2 on the 16/clk/SM MUFU pipe) left the tensor pipe idle ~85 % of the time
// (measured 250-280 TFLOP/s). Here
//   * P never goes through shared memory: the softmax warps write bf16 P back into the TMEM columns S
//     occupied (tcgen05.st, P aliases S) and the PV product takes A from TMEM (tcgen05.mma TS form);
//   * keys are processed in tiles of 64 with S double-buffered: S0/P0 [0,64), S1/P1 [64,128),
//     O [128,192) -> a 256-column TMEM allocation, smem 81 KB
//   => two CTAs are resident per SM and 8 softmax warps keep all four MUFU pipes busy; inside a CTA
//      S_{j+1} is computed while the softmax warps work on tile j.
// tcgen05.mma ops issued by one thread execute in order, which is what makes the S/P aliasing safe.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int AT_THREADS = 192;
constexpr int BM = 128;
constexpr int HD = 64;
constexpr int Q_BYTES = BM * HD * 2;  // 16 KiB
constexpr int S_COL = 0;              // S (fp32) and, after softmax, P (bf16x2) share these columns
constexpr int O_COL = 128;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  CUtensorMap tmQ;   // box {64, 128}
  CUtensorMap tmKV;  // box {64, KT}
  __nv_bfloat16* out;
  float* lse;  // optional [b*heads, T]: log2-domain log-sum-exp of the scaled scores (for the backward pass)
  int T, heads, C;
  int legacy;
  int n_bh;  // batch * heads (the persistent kernel's work list)
};

constexpr int KT = 64;                   // keys per tile
constexpr int KV_STAGES = 4;             // K/V ring depth
constexpr int KV_BYTES = KT * HD * 2;    // 8 KiB per operand per stage
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * 2 * KV_BYTES + 1024;

// barrier indices
constexpr int B_Q = 0;
constexpr int B_KV_FULL = 1;                      // [KV_STAGES]
constexpr int B_KV_EMPTY = B_KV_FULL + KV_STAGES; // [KV_STAGES]
constexpr int B_S_FULL = B_KV_EMPTY + KV_STAGES;  // [2]
constexpr int B_P_FULL = B_S_FULL + 2;            // [2]
constexpr int B_PV_DONE = B_P_FULL + 2;            // [2]: PV_j commits to barrier j&1, phase j>>1
constexpr int NUM_BARS = B_PV_DONE + 2;

// TMEM columns: S0 [0,64) | S1 [64,128) | O [128,192); P_j (bf16x2, 32 columns) aliases the head of
// S_(j&1). Double-buffered S lets S_{j+1} = Q K_{j+1}^T run on the tensor pipe while the softmax warps
// are still on tile j, so they never wait for the MMA round trip (measured 23 % of stall samples with
// a single S buffer).
__global__ void __launch_bounds__(AT_THREADS, 2) attention2_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NUM_BARS];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int st) { return smem_base + Q_BYTES + st * 2 * KV_BYTES; };
  auto v_smem = [&](int st) { return k_smem(st) + KV_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * BM;
  const int row_base = b * p.T;
  const int qc = p.legacy ? h * 3 * HD : h * HD;
  const int kc = p.legacy ? qc + HD : p.C + h * HD;
  const int vc = p.legacy ? qc + 2 * HD : 2 * p.C + h * HD;
  const int nkt = p.T / KT;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    for (int i = 0; i < NUM_BARS; ++i) mbar_init(bar(i), (i == B_P_FULL || i == B_P_FULL + 1) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_slot_s), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(B_Q), Q_BYTES);
      tma_load_2d(q_smem, &p.tmQ, bar(B_Q), qc, row_base + q0);
      for (int j = 0; j < nkt; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t use = (uint32_t)(j / KV_STAGES);
        mbar_wait(bar(B_KV_EMPTY + st), (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(B_KV_FULL + st), 2 * KV_BYTES);
        tma_load_2d(k_smem(st), &p.tmKV, bar(B_KV_FULL + st), kc, row_base + j * KT);
        tma_load_2d(v_smem(st), &p.tmKV, bar(B_KV_FULL + st), vc, row_base + j * KT);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, KT, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, HD, 0, 1);  // B = V, MN-major (as stored)
      auto issue_s = [&](int j) {
        const int st = j % KV_STAGES;
        mbar_wait(bar(B_KV_FULL + st), (uint32_t)(j / KV_STAGES) & 1u);  // K_j, V_j landed
        tc_fence_after();
        // S buffer (j&1) last held P_{j-2}: PV_{j-2} was issued earlier by this thread (in-order pipe)
        const uint64_t a_desc = umma_desc_kmajor_sw128(q_smem);
        const uint64_t b_desc = umma_desc_kmajor_sw128(k_smem(st));
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16_ss(tmem_base + (j & 1) * KT, a_desc + 2u * kk, b_desc + 2u * kk, idesc_s, kk != 0);
        umma_commit(bar(B_S_FULL + (j & 1)));
      };
      mbar_wait(bar(B_Q), 0);
      issue_s(0);
      for (int j = 0; j < nkt; ++j) {
        if (j + 1 < nkt) issue_s(j + 1);
        const int st = j % KV_STAGES;
        mbar_wait(bar(B_P_FULL + (j & 1)), (uint32_t)(j >> 1) & 1u);  // P_j in TMEM (and O rescaled)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          // A = P_j from TMEM: 16 keys = 8 packed columns per step; B = V_j rows [16 kk, +16)
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_smem(st) + kk * 2048, 1024);
          umma_bf16_ts(tmem_base + O_COL, tmem_base + (j & 1) * KT + 8 * kk, b_desc, idesc_o, (j | kk) != 0);
        }
        umma_commit(bar(B_KV_EMPTY + st));  // K/V stage free
        umma_commit(bar(B_PV_DONE + (j & 1)));  // O stable up to tile j
      }
    }
  } else {
    // ===================== softmax (warps 0..3): one query row per thread =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // (64^-1/4)^2 * log2(e)
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
      // 8 independent max chains (a single 64-deep dependent chain costs ~250 cycles of latency that
      // two resident warps per scheduler cannot hide)
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(sr[i]);
#pragma unroll
      for (int i = 8; i < KT; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(sr[i]));
      const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                             fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      // Lazy rescaling: the reference maximum m_run only moves when the tile maximum exceeds it by
      // more than 8 (in log2 units), so p = exp2(s*sc - m_run) <= 256 stays comfortably inside
      // fp32/bf16 range while O and l need rescaling only on the rare big jumps (the final O / l
      // normalisation is exact either way).
      const float m_tile = mx * sc;
      const bool jump = m_tile > m_run + 8.0f;  // always true on the first tile (m_run = -inf)
      float alpha = 1.0f;
      float m_new = m_run;
      if (__any_sync(0xffffffffu, jump)) {
        m_new = fmaxf(m_run, m_tile);
        alpha = ex2_approx(m_run - m_new);  // 0 on the first tile
        if (j > 0) {
          // O may only be touched once PV_{j-1} has completed (S_j was issued before it). S_j being
          // complete implies PV_{j-2} and everything before it is (one in-order pipe), so barrier
          // (j-1)&1 is at most one phase behind the one waited for: the parity wait is unambiguous.
          mbar_wait(bar(B_PV_DONE + ((j - 1) & 1)), (uint32_t)((j - 1) >> 1) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + O_COL + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(lane_addr + O_COL + c, v);
          }
        }
      }
      // p = exp2(s*sc - m_new) (one FFMA + one MUFU per element), row sum; bf16 P packed in place
      // The scale-subtract and the eight row-sum chains run on packed pairs (FFMA2 / FADD2: one issue slot per two
      // scores); same roundings and the same summation order as the scalar form.
      const uint64_t sc2 = f32x2_pack(sc, sc), nm2 = f32x2_pack(-m_new, -m_new);
      uint64_t ps2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ps2[i] = f32x2_pack(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KT / 2; ++i) {
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sc2, nm2), x0, x1);
        const float p0 = ex2_approx(x0);
        const float p1 = ex2_approx(x1);
        ps2[i & 3] = f32x2_add(ps2[i & 3], f32x2_pack(p0, p1));
        sr[i] = pack_bf16x2(p0, p1);
      }
      float ps[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) f32x2_unpack(ps2[i], ps[2 * i], ps[2 * i + 1]);
      const float ps0 = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      const float ps1 = (ps[4] + ps[5]) + (ps[6] + ps[7]);
      // P_j overwrites the head of S_j (all of S_j is in registers by now)
      tmem_st_32x32b_x32(s_addr, sr);
      l_run = l_run * alpha + (ps0 + ps1);
      m_run = m_new;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_P_FULL + (j & 1)));
    }
    // epilogue: O / l -> bf16. PV_{nkt-3} is known complete here (S_{nkt-1} was issued after it); it is
    // the previous phase of the barrier PV_{nkt-1} commits to, so this parity wait cannot alias: it
    // neither passes early (barrier one phase behind) nor hangs (all PVs already complete).
    mbar_wait(bar(B_PV_DONE + ((nkt - 1) & 1)), (uint32_t)((nkt - 1) >> 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const bool ok = (q0 + row) < p.T;
    // P_ij = exp2(s_ij * sc - lse): what attention_bwd.cu recomputes the probabilities from
    if (p.lse != nullptr && ok) p.lse[(size_t)bh * p.T + q0 + row] = m_run + __log2f(l_run);
    __nv_bfloat16* orow = p.out + ((size_t)(row_base + q0 + row)) * p.C + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(lane_addr + O_COL + c, v);
      tmem_wait_ld();
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
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent form, used for T <= 256. A CTA of attention2_kernel lives for T / 64 key tiles only (4 at T = 256, 1 at
// T = 64), and its fixed cost - launch, TMEM allocation, barrier set-up, the first Q / K / V round trip to HBM, the first
// S product, the drain - is worth several key tiles. Here one CTA
// walks many (batch*head, query tile) items: TMEM and barriers are set up once; the TMA warp runs ahead into the next
// item (Q double-buffered, the K/V ring simply continues); the MMA warp issues S of the next item's first key tile while
// the softmax warps are still on the last tile of this one; O is double-buffered (TMEM: S0 | S1 | O0 | O1 = 256 columns)
// so that the next item's PV products do not wait for this item's output to be read. Every barrier phase is derived from
// a running counter (g = key-tile steps done by this CTA, it = items done), never from per-item indices.
constexpr int PB_Q_FULL = 0;                          // [2]
constexpr int PB_Q_EMPTY = 2;                         // [2]
constexpr int PB_KV_FULL = 4;                         // [KV_STAGES]
constexpr int PB_KV_EMPTY = PB_KV_FULL + KV_STAGES;   // [KV_STAGES]
constexpr int PB_S_FULL = PB_KV_EMPTY + KV_STAGES;    // [2]
constexpr int PB_P_FULL = PB_S_FULL + 2;              // [2], 4 arrivals (one per softmax warp)
constexpr int PB_PV_DONE = PB_P_FULL + 2;             // [2]
constexpr int PB_O_EMPTY = PB_PV_DONE + 2;            // [2], 4 arrivals
constexpr int PNUM_BARS = PB_O_EMPTY + 2;
constexpr int PSMEM_BYTES = 2 * Q_BYTES + KV_STAGES * 2 * KV_BYTES + 1024;

__global__ void __launch_bounds__(AT_THREADS, 2) attention2p_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[PNUM_BARS];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto q_smem = [&](int qb) { return smem_base + qb * Q_BYTES; };
  auto k_smem = [&](int st) { return smem_base + 2 * Q_BYTES + st * 2 * KV_BYTES; };
  auto v_smem = [&](int st) { return k_smem(st) + KV_BYTES; };
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkt = p.T / KT;
  const int nq = (p.T + BM - 1) / BM;
  const int total = nq * p.n_bh;
  const int first = (int)blockIdx.x, step = (int)gridDim.x;
  const int n_items = first < total ? (total - first + step - 1) / step : 0;
  // item w -> (batch*head, query tile); the query tile runs fastest so that neighbouring CTAs share K / V in L2
  auto decode = [&](int w, int& bh, int& q0, int& row_base, int& qc, int& kc, int& vc, int& h) {
    bh = w / nq;
    q0 = (w - bh * nq) * BM;
    const int b = bh / p.heads;
    h = bh - b * p.heads;
    row_base = b * p.T;
    qc = p.legacy ? h * 3 * HD : h * HD;
    kc = p.legacy ? qc + HD : p.C + h * HD;
    vc = p.legacy ? qc + 2 * HD : 2 * p.C + h * HD;
  };

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    for (int i = 0; i < PNUM_BARS; ++i) {
      const bool four = (i >= PB_P_FULL && i < PB_P_FULL + 2) || (i >= PB_O_EMPTY && i < PB_O_EMPTY + 2);
      mbar_init(bar(i), four ? 4 : 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_slot_s), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < n_items; ++it) {
        int bh, q0, row_base, qc, kc, vc, h;
        decode(first + it * step, bh, q0, row_base, qc, kc, vc, h);
        const int qb = it & 1;
        mbar_wait(bar(PB_Q_EMPTY + qb), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(PB_Q_FULL + qb), Q_BYTES);
        tma_load_2d(q_smem(qb), &p.tmQ, bar(PB_Q_FULL + qb), qc, row_base + q0);
        for (int j = 0; j < nkt; ++j) {
          const int g = it * nkt + j;
          const int st = g % KV_STAGES;
          mbar_wait(bar(PB_KV_EMPTY + st), ((uint32_t)(g / KV_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(bar(PB_KV_FULL + st), 2 * KV_BYTES);
          tma_load_2d(k_smem(st), &p.tmKV, bar(PB_KV_FULL + st), kc, row_base + j * KT);
          tma_load_2d(v_smem(st), &p.tmKV, bar(PB_KV_FULL + st), vc, row_base + j * KT);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_items > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, KT, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, HD, 0, 1);  // B = V, MN-major (as stored)
      const int total_steps = n_items * nkt;
      auto issue_s = [&](int g) {
        const int it = g / nkt;
        const int j = g - it * nkt;
        const int st = g % KV_STAGES;
        if (j == 0) mbar_wait(bar(PB_Q_FULL + (it & 1)), (uint32_t)(it >> 1) & 1u);
        mbar_wait(bar(PB_KV_FULL + st), (uint32_t)(g / KV_STAGES) & 1u);
        tc_fence_after();
        // S buffer (g & 1) last held P_{g-2}: PV_{g-2} was issued earlier by this thread (in-order pipe)
        const uint64_t a_desc = umma_desc_kmajor_sw128(q_smem(it & 1));
        const uint64_t b_desc = umma_desc_kmajor_sw128(k_smem(st));
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16_ss(tmem_base + (g & 1) * KT, a_desc + 2u * kk, b_desc + 2u * kk, idesc_s, kk != 0);
        umma_commit(bar(PB_S_FULL + (g & 1)));
        if (j == nkt - 1) umma_commit(bar(PB_Q_EMPTY + (it & 1)));  // every S product of this item has been issued
      };
      issue_s(0);
      for (int g = 0; g < total_steps; ++g) {
        if (g + 1 < total_steps) issue_s(g + 1);
        const int it = g / nkt;
        const int j = g - it * nkt;
        const int st = g % KV_STAGES;
        const uint32_t o_tmem = tmem_base + O_COL + (it & 1) * HD;
        if (j == 0) {  // the output of item it-2 has been read out of this O buffer
          mbar_wait(bar(PB_O_EMPTY + (it & 1)), ((uint32_t)(it >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(bar(PB_P_FULL + (g & 1)), (uint32_t)(g >> 1) & 1u);  // P_g in TMEM (and O rescaled)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_smem(st) + kk * 2048, 1024);
          umma_bf16_ts(o_tmem, tmem_base + (g & 1) * KT + 8 * kk, b_desc, idesc_o, (j | kk) != 0);
        }
        umma_commit(bar(PB_KV_EMPTY + st));       // K/V stage free
        umma_commit(bar(PB_PV_DONE + (g & 1)));   // O stable up to this tile
      }
    }
  } else {
    // ===================== softmax + output (warps 0..3): one query row per thread =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // (64^-1/4)^2 * log2(e)
    int g = 0;
    for (int it = 0; it < n_items; ++it) {
      int bh, q0, row_base, qc, kc, vc, h;
      decode(first + it * step, bh, q0, row_base, qc, kc, vc, h);
      const uint32_t o_addr = lane_addr + O_COL + (it & 1) * HD;
      float m_run = -INFINITY;
      float l_run = 0.f;
      for (int j = 0; j < nkt; ++j, ++g) {
        const uint32_t s_addr = lane_addr + (g & 1) * KT;
        mbar_wait(bar(PB_S_FULL + (g & 1)), (uint32_t)(g >> 1) & 1u);
        tc_fence_after();
        uint32_t sr[KT];
        tmem_ld_32x32b_x32(s_addr, sr);
        tmem_ld_32x32b_x32(s_addr + 32, sr + 32);
        tmem_wait_ld();
        float mxs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(sr[i]);
#pragma unroll
        for (int i = 8; i < KT; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(sr[i]));
        const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                               fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
        const float m_tile = mx * sc;
        const bool jump = m_tile > m_run + 8.0f;  // lazy rescaling, as in attention2_kernel
        float alpha = 1.0f;
        float m_new = m_run;
        if (__any_sync(0xffffffffu, jump)) {
          m_new = fmaxf(m_run, m_tile);
          alpha = ex2_approx(m_run - m_new);  // 0 on the first tile
          if (j > 0) {
            // PV_{g-1} complete (the next completion of that barrier needs P_{g+1} from these warps: no aliasing)
            mbar_wait(bar(PB_PV_DONE + ((g - 1) & 1)), (uint32_t)((g - 1) >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < HD; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(o_addr + c, v);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + c, v);
            }
          }
        }
        const uint64_t sc2 = f32x2_pack(sc, sc), nm2 = f32x2_pack(-m_new, -m_new);
        uint64_t ps2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ps2[i] = f32x2_pack(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < KT / 2; ++i) {
          float x0, x1;
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), sc2, nm2), x0, x1);
          const float p0 = ex2_approx(x0);
          const float p1 = ex2_approx(x1);
          ps2[i & 3] = f32x2_add(ps2[i & 3], f32x2_pack(p0, p1));
          sr[i] = pack_bf16x2(p0, p1);
        }
        float ps[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) f32x2_unpack(ps2[i], ps[2 * i], ps[2 * i + 1]);
        const float ps0 = (ps[0] + ps[1]) + (ps[2] + ps[3]);
        const float ps1 = (ps[4] + ps[5]) + (ps[6] + ps[7]);
        tmem_st_32x32b_x32(s_addr, sr);  // P_g overwrites the head of S_g
        l_run = l_run * alpha + (ps0 + ps1);
        m_run = m_new;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(PB_P_FULL + (g & 1)));
      }
      // output of this item: O / l -> bf16 (PV of its last tile is tile g-1 of this CTA)
      mbar_wait(bar(PB_PV_DONE + ((g - 1) & 1)), (uint32_t)((g - 1) >> 1) & 1u);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      const bool ok = (q0 + row) < p.T;
      if (p.lse != nullptr && ok) p.lse[(size_t)bh * p.T + q0 + row] = m_run + __log2f(l_run);
      __nv_bfloat16* orow = p.out + ((size_t)(row_base + q0 + row)) * p.C + h * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(o_addr + c, v);
        tmem_wait_ld();
        if (ok) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[gq * 8 + 0]) * inv, __uint_as_float(v[gq * 8 + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(v[gq * 8 + 2]) * inv, __uint_as_float(v[gq * 8 + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(v[gq * 8 + 4]) * inv, __uint_as_float(v[gq * 8 + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(v[gq * 8 + 6]) * inv, __uint_as_float(v[gq * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + gq * 8) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PB_O_EMPTY + (it & 1)));  // this O buffer may be overwritten (item it + 2)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int launch_attn2p(const AttnParams& ap, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attention2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM_BYTES));
    attr_set = true;
  }
  const int total = ((ap.T + BM - 1) / BM) * ap.n_bh;
  const int grid = total < 2 * num_sms() ? total : 2 * num_sms();
  attention2p_kernel<<<grid, AT_THREADS, PSMEM_BYTES, stream>>>(ap);
  ADB_CUDA(cudaGetLastError());
  return 1;
}

int launch_attn2(const AttnParams& ap, int b, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((ap.T + BM - 1) / BM, b * ap.heads);
  attention2_kernel<<<grid, AT_THREADS, SMEM_BYTES, stream>>>(ap);
  ADB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace

int attention_v1_submit(adb_plan* plan, const void* qkv, void* out, int b, int t, int heads, int legacy_order,
                        cudaStream_t stream);

int attention_submit(adb_plan* plan, const void* qkv, void* out, float* lse, int b, int t, int heads, int legacy_order,
                     cudaStream_t stream) {
  static int use_v1 = -1;
  if (use_v1 < 0) {
    const char* e = getenv("ADB_ATTENTION_V1");
    use_v1 = (e && e[0] == '1') ? 1 : 0;
  }
  if (use_v1 && lse == nullptr) return attention_v1_submit(plan, qkv, out, b, t, heads, legacy_order, stream);
  ADB_REQUIRE(qkv && out && b > 0 && heads > 0, "attention: bad arguments");
  ADB_REQUIRE(t == 64 || (t >= 128 && t % 128 == 0), "attention: sequence length %d unsupported (64 or a multiple of 128)", t);
  const int C = heads * HD;
  const int bn = KT;
  AttnParams ap;
  memset(&ap, 0, sizeof(ap));
  const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)b * t};
  const uint64_t strides[1] = {(uint64_t)3 * C * 2};
  const uint32_t boxq[2] = {64, 128};
  const uint32_t boxkv[2] = {64, (uint32_t)bn};
  int r = make_tmap_bf16(&ap.tmQ, qkv, 2, dims, strides, boxq);
  if (r != ADB_OK) return r;
  r = make_tmap_bf16(&ap.tmKV, qkv, 2, dims, strides, boxkv);
  if (r != ADB_OK) return r;
  ap.out = reinterpret_cast<__nv_bfloat16*>(out);
  ap.lse = lse;
  ap.T = t;
  ap.heads = heads;
  ap.C = C;
  ap.legacy = legacy_order ? 1 : 0;
  ap.n_bh = b * heads;
  // The persistent kernel for short sequences (measured at batch 256: T = 256 331 -> 419 TFLOP/s, T = 64 80 -> 111; at
  // T = 1024 the per-CTA kernel is 8 % faster, 665 vs 611). ADB_ATTN_PERSIST=0 / 1 forces one form (A/B testing).
  static int persist = -2;
  if (persist == -2) {
    const char* e = getenv("ADB_ATTN_PERSIST");
    persist = (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
  }
  const int use_p = persist >= 0 ? persist : (t <= 256 ? 1 : 0);
  const double flops = 4.0 * (double)b * heads * (double)t * (double)t * HD;  // QK^T and PV
  return submit(plan, stream, "attention", flops, 0.0, [ap, b, use_p](cudaStream_t s) -> int {
    return use_p ? launch_attn2p(ap, s) : launch_attn2(ap, b, s);
  });
}

}  // namespace adb
