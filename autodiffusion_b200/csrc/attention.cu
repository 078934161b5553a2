// attention.cu — fused softmax attention for head dim 64 on tcgen05 + TMEM.
//
// Replaces QKVAttention.forward / QKVAttentionLegacy.forward
// (guided_diffusion/dynamic_unet.py:390-409 / 357-374):
//     w = softmax_fp32((q*s)^T (k*s)),  a = v w^T,  s = 64^(-1/4)
// without ever materialising the T x T score matrix in HBM (the reference writes
// [B*heads, T, T] twice). Both channel layouts are just different column offsets into the
// k=1-conv output matrix qkv[B*T, 3*C] (pixels as rows):
//   new order   : q = h*64,      k = C + h*64,   v = 2C + h*64
//   legacy order: q = h*192,     k = h*192 + 64, v = h*192 + 128
//
// One CTA = one 128-query tile of one (batch, head); keys are streamed in tiles of BN.
//   warp 4    : TMA producer (Q once; K,V double-buffered), 128-byte swizzled boxes.
//   warp 5    : one thread issues tcgen05.mma (top warp ids win SMSP arbitration):
//                 S_j = Q K_j^T   (128 x BN x 64;  A, B K-major)        -> TMEM S[j&1]
//                 O  += P_j V_j   (128 x 64 x BN;  A = P K-major smem,
//                                  B = V MN-major smem, i.e. V is used as stored) -> TMEM O
//   warps 0-3 : online softmax, one query row per thread (TMEM lane = row): running max and
//               sum in fp32, exp2 with the scale folded in, P written to smem as bf16 in the
//               swizzled K-major layout the MMA expects, O rescaled in TMEM when the max moves.
// S is double-buffered in TMEM so S_{j+1} is computed while softmax works on S_j.
#include <string.h>

#include "common.cuh"

namespace adb {

namespace {

constexpr int AT_THREADS = 192;
constexpr int BM = 128;
constexpr int HD = 64;
constexpr int Q_BYTES = BM * HD * 2;  // 16 KiB
constexpr int O_COL = 256;            // TMEM column of the O accumulator

struct AttnParams {
  CUtensorMap tmQ;   // box {64, 128}
  CUtensorMap tmKV;  // box {64, BN}
  __nv_bfloat16* out;
  int T, heads, C;
  int legacy;
};

template <int BN>
struct ACfg {
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int P_BYTES = BM * BN * 2;
  // padded above half an SM's shared memory: the 512-column TMEM allocation allows one CTA per
  // SM anyway, and a second resident CTA would only spin inside tcgen05.alloc
  static constexpr int RAW_BYTES = Q_BYTES + 4 * KV_BYTES + P_BYTES + 1024;
  static constexpr int SMEM_BYTES = RAW_BYTES > 120 * 1024 ? RAW_BYTES : 120 * 1024;
};

template <int BN>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  using C = ACfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // barriers: 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5,6 s_full | 7,8 s_empty | 9 p_full | 10 pv_done
  __shared__ __align__(8) uint64_t bars[11];
  __shared__ uint32_t tmem_slot_s;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int st) { return smem_base + Q_BYTES + st * 2 * C::KV_BYTES; };
  auto v_smem = [&](int st) { return k_smem(st) + C::KV_BYTES; };
  const uint32_t p_smem = smem_base + Q_BYTES + 4 * C::KV_BYTES;
  uint8_t* p_gen = smem_gen + Q_BYTES + 4 * C::KV_BYTES;
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
  const int nkt = p.T / BN;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmKV);
    mbar_init(bar(0), 1);
    for (int i = 1; i <= 6; ++i) mbar_init(bar(i), 1);
    mbar_init(bar(7), 4);
    mbar_init(bar(8), 4);
    mbar_init(bar(9), 4);
    mbar_init(bar(10), 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_slot_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(0), Q_BYTES);
      tma_load_2d(q_smem, &p.tmQ, bar(0), qc, row_base + q0);
      for (int j = 0; j < nkt; ++j) {
        const int st = j & 1;
        mbar_wait(bar(3 + st), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(1 + st), 2 * C::KV_BYTES);
        tma_load_2d(k_smem(st), &p.tmKV, bar(1 + st), kc, row_base + j * BN);
        tma_load_2d(v_smem(st), &p.tmKV, bar(1 + st), vc, row_base + j * BN);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BM, HD, 0, 1);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(bar(1 + st), (uint32_t)(j >> 1) & 1u);          // K_j, V_j landed
        mbar_wait(bar(7 + st), ((uint32_t)(j >> 1) & 1u) ^ 1u);   // S[st] drained by softmax
        tc_fence_after();
        const uint64_t a_desc = umma_desc_kmajor_sw128(q_smem);
        const uint64_t b_desc = umma_desc_kmajor_sw128(k_smem(st));
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16_ss(tmem_base + st * BN, a_desc + 2u * kk, b_desc + 2u * kk, idesc_s, kk != 0);
        umma_commit(bar(5 + st));
      };
      mbar_wait(bar(0), 0);
      issue_s(0);
      for (int j = 0; j < nkt; ++j) {
        if (j + 1 < nkt) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(bar(9), (uint32_t)j & 1u);  // P_j written (and O rescaled)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(p_smem + (kk >> 2) * (BM * 128)) + 2u * (kk & 3);
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_smem(st) + kk * 2048, 1024);
          umma_bf16_ss(tmem_base + O_COL, a_desc, b_desc, idesc_o, (j | kk) != 0);
        }
        umma_commit(bar(3 + st));  // K/V stage free
        umma_commit(bar(10));      // P buffer free, O stable
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // (64^-1/4)^2 * log2(e)
    float m_run = -INFINITY;
    float l_run = 0.f;
    for (int j = 0; j < nkt; ++j) {
      const int st = j & 1;
      mbar_wait(bar(5 + st), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + st * BN + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx * sc);
      const float alpha = exp2f(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      // P buffer and O are stable once PV_{j-1} has completed
      if (j > 0) {
        mbar_wait(bar(10), (uint32_t)(j - 1) & 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + O_COL + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(lane_addr + O_COL + c, v);
          }
          tmem_wait_st();
        }
      }
      // pass 2: p = exp2(s*sc - m_new), row sum, bf16 P into swizzled K-major smem
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + st * BN + c, v);
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          f[i] = exp2f(fmaf(__uint_as_float(v[i]), sc, -m_new));
          psum += f[i];
        }
        uint8_t* sub = p_gen + (c >> 6) * (BM * 128) + row * 128;
        const int chunk0 = (c & 63) >> 3;  // 16-byte chunk index of key c within the 64-key subtile
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
          o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
          o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
          o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
          *reinterpret_cast<uint4*>(sub + (((chunk0 + g) ^ (row & 7)) << 4)) = o;
        }
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;
      fence_proxy_async_smem();  // P visible to the MMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(7 + st));  // S[st] may be overwritten
        mbar_arrive(bar(9));       // P_j ready
      }
    }
    // epilogue: O / l -> bf16
    mbar_wait(bar(10), (uint32_t)(nkt - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const bool ok = (q0 + row) < p.T;
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
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN>
int launch_attn(const AttnParams& ap, int b, cudaStream_t stream) {
  using C = ACfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(attention_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((ap.T + BM - 1) / BM, b * ap.heads);
  attention_kernel<BN><<<grid, AT_THREADS, C::SMEM_BYTES, stream>>>(ap);
  ADB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace

int attention_v1_submit(adb_plan* plan, const void* qkv, void* out, int b, int t, int heads, int legacy_order,
                     cudaStream_t stream) {
  ADB_REQUIRE(qkv && out && b > 0 && heads > 0, "attention: bad arguments");
  ADB_REQUIRE(t == 64 || (t >= 128 && t % 128 == 0), "attention: sequence length %d unsupported (64 or a multiple of 128)", t);
  const int C = heads * HD;
  const int bn = (t == 64) ? 64 : 128;
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
  ap.T = t;
  ap.heads = heads;
  ap.C = C;
  ap.legacy = legacy_order ? 1 : 0;
  const double flops = 4.0 * (double)b * heads * (double)t * (double)t * HD;  // QK^T and PV
  return submit(plan, stream, "attention", flops, 0.0, [ap, b, bn](cudaStream_t s) -> int {
    return bn == 64 ? launch_attn<64>(ap, b, s) : launch_attn<128>(ap, b, s);
  });
}

}  // namespace adb
