// conv_igemm.cu — implicit-GEMM 3x3 / 1x1 convolution and k=1 GEMM on tcgen05 + TMEM.
//
// Replaces every nn.Conv2d 3x3/1x1 and nn.Conv1d k=1 of the UNet torso
// (reference: guided_diffusion/nn.py:22-32 via dynamic_unet.py:194,220,231,303,311,653).
//
// GEMM view: rows = output pixels m = (n*H + y)*W + x of a bf16 NHWC tensor, cols = Cout,
// K = sum over segments of taps*Cin. One CTA computes a 128 x BLOCK_N tile:
//   warp 8    : TMA producer. The A tile of tap (dy,dx), channel chunk c0 is ONE 4-D tiled
//               TMA box {64 ch, bw, bh, bn} at (c0, x0+dx, y0+dy, n0); out-of-bounds rows
//               and columns are zero-filled by the TMA unit = the conv's zero padding.
//               The B tile is a 2-D box {64, BLOCK_N} of the K-major weight matrix.
//   warp 9    : allocates TMEM, issues tcgen05.mma (128 x BLOCK_N x 16, bf16 -> fp32).
//   warps 0-7 : epilogue (two warps per TMEM lane quarter, half of the columns each). tcgen05.ld the accumulator (lane = pixel row), + bias,
//               + residual (same / 2x2-avg-pooled / nearest-upsampled source), store.
// Persistent grid (<= #SM CTAs), static round-robin over tiles with n fastest so CTAs that
// run concurrently share the A tile in L2; smem ring of STAGES stages; two TMEM accumulator
// stages so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace adb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int NUM_THREADS = 320;  // 8 epilogue warps, TMA warp, MMA warp
// The SMSP arbiter favours the highest warp id among eligible warps: the two single-thread issue
// warps get the top ids so that epilogue instruction streams cannot starve TMA / tcgen05.mma issue.
constexpr int TMA_WARP = 8;
constexpr int MMA_WARP = 9;
constexpr int STAT_BINS = 40;      // >= groups one epilogue warp can touch per tile (96 cols / cpg + 2)

struct ConvKParams {
  CUtensorMap tmA[3];
  CUtensorMap tmW;  // box {64, BLOCK_N / NCTA}: in 2-CTA mode each CTA of the pair loads half of the N tile
  CUtensorMap tmOut;  // bf16 [M, cout] output, box {32 ch, 32 rows}, 64-byte swizzle (epilogue TMA stores)
  CUtensorMap tmRes;  // same geometry over the same-resolution residual (epilogue TMA loads)
  int tma_epi;        // 1: epilogue moves output (and RES_SAME residual) by TMA through shared memory
  int seg_taps[3];
  int seg_cin[3];
  int seg_chunks[3];
  int seg_kbase[3];
  int seg_stride[3];  // 2: 3x3 stride-2 pad-1 segment; tmA[s] is then the 5-D view [n, h, 2, w, 2*cin] of the input
  int nseg;
  int M, cout, H, W;
  int m_tiles, n_tiles;
  int k_iters;
  const float* bias;
  int bias_stride;  // floats between consecutive images' bias rows (0: one shared row)
  const __nv_bfloat16* residual;
  int res_mode;
  void* out;
  int out_mode;
  double* stats;   // optional [n][32][2] GroupNorm sums of the output, or nullptr
  int cpg;         // channels per group = cout / 32
  // second, optional target: the GroupNorm of a consumer that reads this tensor as channels
  // [choff2, choff2 + cout) of a concat (dynamic_unet.py:699) whose groups are cpg2 channels wide
  double* stats2;
  int cpg2;
  int choff2;
  // x / cpg == (x * magic) >> 32 for the column indices that occur (x * cpg < 2^32)
  unsigned long long cpg_magic, cpg2_magic;
  // GroupNorm-backward sums mode (adb_conv_desc.gnb_*): this conv is a data gradient dY; tmRes then maps the consumer
  // GroupNorm's forward input x, `stats` is the target of sum dxh / sum dxh*xh, and nothing is added to the output.
  int b_one_box;  // 256-wide pair tile: the CTA's 128 weight rows are ONE TMA box (default; ADB_CONV_B128BOX=0: two 64-row boxes)
  int gnb;
  double gnb_inv_cnt;  // 1 / (channels per group * pixels per image)
  const double* gnb_stats;
  const float* gnb_gamma;
  const float* gnb_beta;
  const float* gnb_ss;
  int gnb_ss_stride;
  float gnb_eps;
  int gnb_silu;
};

__device__ __forceinline__ int fast_div(int x, unsigned long long magic) {
  return (int)(((unsigned long long)(unsigned)x * magic) >> 32);
}

template <int BLOCK_N, int NCTA>
struct Cfg {
  static constexpr int B_STAGE_BYTES = (BLOCK_N / NCTA) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per CTA
  // epilogue staging: per epilogue warp two 2 KB residual buffers + one 2 KB output buffer
  static constexpr int EPI_WARP_BYTES = 3 * 2048;
  static constexpr int EPI_BYTES = 8 * EPI_WARP_BYTES;
  static constexpr int RING_BUDGET = 227 * 1024 - EPI_BYTES - 2048;
  static constexpr int STAGES = RING_BUDGET / STAGE_BYTES > 8 ? 8 : RING_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/;
};

// GNB: the epilogue also reduces the consumer GroupNorm's backward sums (ConvKParams::gnb); a separate instantiation so
// that the plain kernel's register allocation is untouched.
template <int BLOCK_N, int NCTA, bool GNB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  using C = Cfg<BLOCK_N, NCTA>;
  // NCTA == 2: the two CTAs of a cluster (one TPC) compute a 256 x BLOCK_N tile with
  // tcgen05.mma.cta_group::2. Each CTA stages its own 128 A rows and HALF of the B tile, so the
  // bytes an SM must ingest per MMA drop from 40 KB to 28 KB per 64-deep K step (the measured limiter
  // of the 1-CTA kernel: ~48 B/clk/SM of TMA traffic at ~51 % tensor-pipe activity). CTA 0 issues the
  // MMAs; full barriers live in CTA 0 (both producers arrive there), empty / accumulator-ready
  // barriers are multicast to both CTAs by tcgen05.commit.
  const uint32_t cta_rank = (NCTA == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_slot_s;
  __shared__ float stat_bins[8][2 * STAT_BINS][2];  // [warp][target 0 | target 1]
  __shared__ float stat_cols[8][64];  // per epilogue warp: column sums / sums of squares of one chunk
  __shared__ __align__(8) uint64_t res_bars[8][2];  // per epilogue warp: residual TMA landed (2 buffers)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = smem_u32(&tmem_slot_s);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == TMA_WARP && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmW);
    if (p.tma_epi) {
      tma_prefetch_desc(&p.tmOut);
      if (p.res_mode == ADB_RES_SAME) tma_prefetch_desc(&p.tmRes);
    }
    for (int w = 0; w < 8; ++w) {
      mbar_init(smem_u32(&res_bars[w][0]), 1);
      mbar_init(smem_u32(&res_bars[w][1]), 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), NCTA);  // one producer arrive per CTA of the pair (on the leader's barrier)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8 * NCTA);  // one arrive per epilogue warp (of both CTAs, on the leader's)
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    if (NCTA == 2) {
      tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot_s);

  // work items: (group of NCTA consecutive m-tiles) x n-tile, n fastest; one item per cluster at a time
  const int m_groups = (p.m_tiles + NCTA - 1) / NCTA;
  const int num_tiles = m_groups * p.n_tiles;
  const int tile0 = blockIdx.x / NCTA;
  const int tile_step = gridDim.x / NCTA;
  const int P = p.H * p.W;

  if (warp == TMA_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_tile = (tile / p.n_tiles) * NCTA + (int)cta_rank;
        const int n_tile = tile % p.n_tiles;
        const int m0 = m_tile * BLOCK_M;
        const int img = m0 / P;
        const int rem = m0 - img * P;
        const int y0 = rem / p.W;
        const int x0 = rem - y0 * p.W;
        // one instance per segment kind, so that the unit-stride loop (every layer of the ADM models) carries none of
        // the stride-2 bookkeeping: with a shared loop the extra per-tap / per-chunk instructions of this single issuing
        // thread cost the 192-wide tiles 10-12 % (measured, r01 v16 -> v18)
        auto produce = [&](auto s2_tag, int s) {
          constexpr bool S2 = decltype(s2_tag)::value;
          const int taps = p.seg_taps[s];
          const int chunks = p.seg_chunks[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            const int kb = p.seg_kbase[s] + tap * p.seg_cin[s];
            // stride 2 (Downsample.op): input row 2y+dy is row-pair y + (dy < 0 ? -1 : 0), parity (dy != 0) of the
            // [n, h, 2, w, 2*cin] view; same along x, where the parity selects the upper cin channels of a pixel pair
            const int sy = dy < 0 ? -1 : 0, py = dy != 0 ? 1 : 0;
            const int sx = dx < 0 ? -1 : 0, pxc = dx != 0 ? p.seg_cin[s] : 0;
            for (int ch = 0; ch < chunks; ++ch) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
              const uint32_t b_dst = a_dst + A_STAGE_BYTES;
              if (NCTA == 2) {
                const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
                if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                else mbar_arrive_cluster(lead_full);
                if (S2) tma_load_5d_2sm(a_dst, &p.tmA[s], lead_full, pxc + ch * BLOCK_K, x0 + sx, py, y0 + sy, img);
                else tma_load_4d_2sm(a_dst, &p.tmA[s], lead_full, ch * BLOCK_K, x0 + dx, y0 + dy, img);
                if (BLOCK_N == 256 && !p.b_one_box) {  // two 64-row boxes
                  tma_load_2d_2sm(b_dst, &p.tmW, lead_full, kb + ch * BLOCK_K, n_tile * BLOCK_N + (int)cta_rank * 128);
                  tma_load_2d_2sm(b_dst + 8192, &p.tmW, lead_full, kb + ch * BLOCK_K, n_tile * BLOCK_N + (int)cta_rank * 128 + 64);
                } else {
                  tma_load_2d_2sm(b_dst, &p.tmW, lead_full, kb + ch * BLOCK_K,
                                  n_tile * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
                }
              } else {
                mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                if (S2) tma_load_5d(a_dst, &p.tmA[s], full_bar(stage), pxc + ch * BLOCK_K, x0 + sx, py, y0 + sy, img);
                else tma_load_4d(a_dst, &p.tmA[s], full_bar(stage), ch * BLOCK_K, x0 + dx, y0 + dy, img);
                tma_load_2d(b_dst, &p.tmW, full_bar(stage), kb + ch * BLOCK_K, n_tile * BLOCK_N);
              }
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        };
        for (int s = 0; s < p.nseg; ++s) {
          if (p.seg_stride[s] == 2) produce(std::true_type{}, s);
          else produce(std::false_type{}, s);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0 && is_leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M * NCTA, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int k = 0; k < p.k_iters; ++k) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
          const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in 16-byte units
            if (NCTA == 2) umma_bf16_ss_2sm(d_tmem, a_desc + 2u * kk, b_desc + 2u * kk, idesc, (k | kk) != 0);
            else umma_bf16_ss(d_tmem, a_desc + 2u * kk, b_desc + 2u * kk, idesc, (k | kk) != 0);
          }
          if (NCTA == 2) umma_commit_2sm_mc(empty_bar(stage), 0x3);
          else umma_commit(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (NCTA == 2) umma_commit_2sm_mc(tfull_bar(acc), 0x3);
        else umma_commit(tfull_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 0..7) =====================
    // Two warps per TMEM lane quarter (warp % 4), each draining half of the tile's columns.
    const int ew = warp;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int CHUNK_COLS = (BLOCK_N >= 32) ? 32 : BLOCK_N;
    constexpr int CHUNKS = BLOCK_N / CHUNK_COLS;
    constexpr int CH_HALF = (CHUNKS + 1) / 2;
    const int cbeg = half * CH_HALF * CHUNK_COLS;
    const int cend = min(BLOCK_N, cbeg + CH_HALF * CHUNK_COLS);
    const int row = quarter * 32 + lane;
    float* my_bins = &stat_bins[ew][0][0];
    float* my_cols = &stat_cols[ew][0];
    // per-warp staging buffers behind the operand ring: [residual 0 | residual 1 | output], 2 KB each
    const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES + ew * C::EPI_WARP_BYTES;
    uint8_t* epi_gen = smem_raw + (epi_base - smem_u32(smem_raw));
    const uint32_t epi_out = epi_base + 4096;
    uint8_t* epi_out_gen = epi_gen + 4096;
    const uint32_t res_bar0 = smem_u32(&res_bars[ew][0]);
    uint32_t res_phase = 0u;  // bit b = parity to wait for on residual buffer b
    int res_buf = 0;
    float* my_bins2 = my_bins + 2 * STAT_BINS;
    if (p.stats != nullptr) {
      for (int i = lane; i < STAT_BINS * 4; i += 32) my_bins[i] = 0.f;
      __syncwarp();
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_tile = (tile / p.n_tiles) * NCTA + (int)cta_rank;
      const int n_tile = tile % p.n_tiles;
      const int m = m_tile * BLOCK_M + row;
      const int n0 = n_tile * BLOCK_N;
      // Fast path (warp-uniform): a full tile half - every row and column valid - staged through shared memory
      // by TMA, residual none or same-resolution. No per-element predicates, group binning by a segmented
      // warp scan, and the residual sub-tiles of the first two chunks are requested BEFORE the accumulator
      // wait so that their HBM latency hides behind the main loop.
      const bool fast = p.tma_epi && CHUNK_COLS == 32 && (m_tile + 1) * BLOCK_M <= p.M && cbeg < cend &&
                        n0 + cend <= p.cout && (p.res_mode == ADB_RES_NONE || p.res_mode == ADB_RES_SAME);
      const int row0f = m_tile * BLOCK_M + quarter * 32;
      const int nch = (cend - cbeg) / CHUNK_COLS;
      if (fast && p.res_mode == ADB_RES_SAME && lane == 0) {
        fence_proxy_async_smem();
        for (int ci = 0; ci < 2 && ci < nch; ++ci) {
          const uint32_t b = (uint32_t)(res_buf ^ ci);
          mbar_arrive_expect_tx(res_bar0 + 8u * b, 2048);
          tma_load_2d(epi_base + 2048u * b, &p.tmRes, res_bar0 + 8u * b, n0 + cbeg + ci * CHUNK_COLS, row0f);
        }
      }
      // gnb: per-column coefficients of the consumer GroupNorm (lane j <-> channel col + j of image imgw), fetched one chunk
      // ahead of their use so that the L2 latency never sits in front of the column walk
      struct { double sum, sq; float ga, be, sc, sh; } gn = {0.0, 0.0, 0.f, 0.f, 0.f, 0.f};
      const int imgw = GNB ? row0f / P : 0;  // the image this warp's 32 rows belong to (h*w % 32 == 0)
      auto gnb_fetch = [&](int col) {
        const int ch = col + lane;
        const double* sp = p.gnb_stats + ((size_t)imgw * 32 + fast_div(ch, p.cpg_magic)) * 2;
        gn.sum = __ldg(sp);
        gn.sq = __ldg(sp + 1);
        gn.ga = __ldg(p.gnb_gamma + ch);
        gn.be = __ldg(p.gnb_beta + ch);
        if (p.gnb_ss != nullptr) {
          gn.sc = __ldg(p.gnb_ss + (size_t)imgw * p.gnb_ss_stride + ch);
          gn.sh = __ldg(p.gnb_ss + (size_t)imgw * p.gnb_ss_stride + p.cout + ch);
        }
      };
      if (GNB && fast) gnb_fetch(n0 + cbeg);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const bool row_ok = m < p.M;
      // geometry of this output pixel (for resampled residuals / NCHW stores)
      const int img = m / P;
      const int rem = m - img * P;
      const int y = rem / p.W;
      const int x = rem - y * p.W;
      // this pixel's bias row: shared (stride 0) or per image (SD ResBlock: conv bias + timestep-embedding projection)
      const float* bias_row = p.bias + (row_ok ? (size_t)img * p.bias_stride : 0);
      const int g_lo = (p.stats != nullptr) ? fast_div(n0 + cbeg, p.cpg_magic) : 0;
      const int g_lo2 = (p.stats2 != nullptr) ? fast_div(p.choff2 + n0 + cbeg, p.cpg2_magic) : 0;
      // single-source residuals (same / nearest-up) are fetched one chunk AHEAD so their HBM
      // latency overlaps the previous chunk's math; the 4-source average-pool variant is not.
      const bool res_tma = p.tma_epi && p.res_mode == ADB_RES_SAME;  // warp-uniform
      const bool res_pf = !res_tma && row_ok && (p.res_mode == ADB_RES_SAME || p.res_mode == ADB_RES_NEAREST2);
      const int row0 = m_tile * BLOCK_M + quarter * 32;  // first output row of this warp's 32x32 sub-tiles
      const __nv_bfloat16* res_row = nullptr;
      if (res_pf) {
        const size_t pix = (p.res_mode == ADB_RES_SAME)
                               ? (size_t)m
                               : ((size_t)img * (p.H / 2) + (y >> 1)) * (p.W / 2) + (x >> 1);
        res_row = p.residual + pix * p.cout;
      }
      uint4 rnext[4];
      auto fetch_res = [&](int col) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rnext[g] = make_uint4(0u, 0u, 0u, 0u);
          if (res_pf && col + g * 8 < p.cout) rnext[g] = __ldg(reinterpret_cast<const uint4*>(res_row + col) + g);
        }
      };
      if (!fast) fetch_res(n0 + cbeg);
      if (!fast && res_tma && lane == 0 && cbeg < cend) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(res_bar0 + 8u * res_buf, 2048);
        tma_load_2d(epi_base + 2048u * res_buf, &p.tmRes, res_bar0 + 8u * res_buf, n0 + cbeg, row0);
      }
      if (fast) {
        const bool res_tma_f = p.res_mode == ADB_RES_SAME;
#pragma unroll 1
        for (int ci = 0; ci < nch; ++ci) {
          const int c = cbeg + ci * CHUNK_COLS;
          const int col0 = n0 + c;
          // gnb: this chunk's column coefficients were requested one chunk ago (the next chunk's are requested after the
          // staging fence below - a membar would wait for them)
          const double g_sum = gn.sum, g_sq = gn.sq;
          const float g_ga = gn.ga, g_be = gn.be, g_sc = gn.sc, g_sh = gn.sh;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + c, v);
          float4 bv[8];
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_row + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = __ldg(b4 + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_wait_ld();
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bv[j].x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bv[j].y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bv[j].z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bv[j].w;
          }
          if (res_tma_f) {
            mbar_wait(res_bar0 + 8u * res_buf, (res_phase >> res_buf) & 1u);
            res_phase ^= 1u << res_buf;
            const uint8_t* rrow = epi_gen + 2048 * res_buf + lane * 64;
            uint4 rc[4];
            if (!GNB)
#pragma unroll
            for (int g = 0; g < 4; ++g) rc[g] = *reinterpret_cast<const uint4*>(rrow + ((g ^ ((lane >> 1) & 3)) << 4));
            __syncwarp();
            if (!GNB && ci + 2 < nch && lane == 0) {  // this buffer is free again: request the chunk after next
              fence_proxy_async_smem();
              mbar_arrive_expect_tx(res_bar0 + 8u * res_buf, 2048);
              tma_load_2d(epi_base + 2048u * res_buf, &p.tmRes, res_bar0 + 8u * res_buf, col0 + 2 * CHUNK_COLS, row0f);
            }
            res_buf ^= 1;
            if (!GNB)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              f[g * 8 + 0] += bf16_lo(rc[g].x);
              f[g * 8 + 1] += bf16_hi(rc[g].x);
              f[g * 8 + 2] += bf16_lo(rc[g].y);
              f[g * 8 + 3] += bf16_hi(rc[g].y);
              f[g * 8 + 4] += bf16_lo(rc[g].z);
              f[g * 8 + 5] += bf16_hi(rc[g].z);
              f[g * 8 + 6] += bf16_lo(rc[g].w);
              f[g * 8 + 7] += bf16_hi(rc[g].w);
            }
          }
          uint32_t ow[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) ow[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          if (lane == 0) tma_store_wait_read0();  // previous store has finished reading the staging buffer
          __syncwarp();
          uint8_t* orow = epi_out_gen + lane * 64;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(orow + ((g ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(ow[4 * g], ow[4 * g + 1], ow[4 * g + 2], ow[4 * g + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmOut, epi_out, col0, row0f);
            tma_store_commit();
          }
          if (GNB && ci + 1 < nch) gnb_fetch(col0 + CHUNK_COLS);
          if (p.stats != nullptr) {
            // column sums of the STORED values: lane j walks column j down the staged tile's 32 rows
            float cs = 0.f, cq = 0.f;
            const uint32_t cch = (uint32_t)lane >> 3, cin8 = ((uint32_t)lane & 7u) * 2u;
            if (GNB) {
              // GroupNorm-backward sums: the stored values are dY; the matching 32 x 32 tile of the GroupNorm's input x sits
              // in the residual buffer waited for above (res_buf was flipped since). Lane j owns channel col0 + j.
              // With hz = z / 2 = x * a + b: silu'(z) = (1 + th)(1 + hz (1 - th)) / 2, th = tanh(hz); the sums are taken
              // against raw x and turned into sums against xh = x * rstd + m0 once per column.
              const double mean = g_sum * p.gnb_inv_cnt;
              const float varf = fmaxf((float)fma(g_sq, p.gnb_inv_cnt, -mean * mean), 0.f) + p.gnb_eps;
              float rstd = rsqrtf(varf);
              rstd *= fmaf(-0.5f * varf * rstd, rstd, 1.5f);  // one Newton step: within an ulp of 1 / sqrt
              const float m0 = -(float)mean * rstd;
              const float sc = 1.0f + g_sc;
              const float A = g_ga * sc;
              const float Bc = fmaf(g_be, sc, g_sh);
              const uint8_t* xt = epi_gen + 2048 * (res_buf ^ 1);
              float sx = 0.f;
              if (p.gnb_silu) {
                const float hA = 0.5f * A, a = hA * rstd, b = fmaf(hA, m0, 0.5f * Bc);
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                  const uint32_t off = r * 64 + (((cch ^ ((uint32_t)(r >> 1) & 3u)) << 4) + cin8);
                  const float dy = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(epi_out_gen + off)) << 16);
                  const float xv = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(xt + off)) << 16);
                  const float hz = fmaf(xv, a, b);
                  float th;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hz));
                  const float dxh = dy * (fmaf(hA, th, hA) * fmaf(hz, 1.0f - th, 1.0f));
                  cs += dxh;
                  sx = fmaf(dxh, xv, sx);
                }
              } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                  const uint32_t off = r * 64 + (((cch ^ ((uint32_t)(r >> 1) & 3u)) << 4) + cin8);
                  const float dy = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(epi_out_gen + off)) << 16);
                  const float xv = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(xt + off)) << 16);
                  cs += dy;
                  sx = fmaf(dy, xv, sx);
                }
                cs *= A;
                sx *= A;
              }
              cq = fmaf(rstd, sx, m0 * cs);
              __syncwarp();
              if (ci + 2 < nch && lane == 0) {  // the x buffer is free now: request the chunk after next
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(res_bar0 + 8u * (res_buf ^ 1), 2048);
                tma_load_2d(epi_base + 2048u * (res_buf ^ 1), &p.tmRes, res_bar0 + 8u * (res_buf ^ 1), col0 + 2 * CHUNK_COLS, row0f);
              }
            } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const uint16_t hv = *reinterpret_cast<const uint16_t*>(
                  epi_out_gen + r * 64 + (((cch ^ ((uint32_t)(r >> 1) & 3u)) << 4) + cin8));
              const float val = __uint_as_float((uint32_t)hv << 16);
              cs += val;
              cq = fmaf(val, val, cq);
            }
            }
            // groups are runs of consecutive lanes: segmented inclusive scan, the last lane of a run owns its bin
            {
              const int gi = fast_div(col0 + lane, p.cpg_magic);
              float s1 = cs, q1 = cq;
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                const float s2 = __shfl_up_sync(0xffffffffu, s1, off);
                const float q2 = __shfl_up_sync(0xffffffffu, q1, off);
                const int g2 = __shfl_up_sync(0xffffffffu, gi, off);
                if (lane >= off && g2 == gi) {
                  s1 += s2;
                  q1 += q2;
                }
              }
              const int gn = __shfl_down_sync(0xffffffffu, gi, 1);
              if (lane == 31 || gn != gi) {
                my_bins[2 * (gi - g_lo)] += s1;
                my_bins[2 * (gi - g_lo) + 1] += q1;
              }
            }
            if (p.stats2 != nullptr) {
              const int gi = fast_div(p.choff2 + col0 + lane, p.cpg2_magic);
              float s1 = cs, q1 = cq;
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                const float s2 = __shfl_up_sync(0xffffffffu, s1, off);
                const float q2 = __shfl_up_sync(0xffffffffu, q1, off);
                const int g2 = __shfl_up_sync(0xffffffffu, gi, off);
                if (lane >= off && g2 == gi) {
                  s1 += s2;
                  q1 += q2;
                }
              }
              const int gn = __shfl_down_sync(0xffffffffu, gi, 1);
              if (lane == 31 || gn != gi) {
                my_bins2[2 * (gi - g_lo2)] += s1;
                my_bins2[2 * (gi - g_lo2) + 1] += q1;
              }
            }
            __syncwarp();
          }
        }
      } else {
#pragma unroll 1
      for (int c = cbeg; c < cend; c += CHUNK_COLS) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + c;
        if (CHUNK_COLS == 32) {
          tmem_ld_32x32b_x32(taddr, v);
        } else {
          tmem_ld_32x32b_x16(taddr, v);
        }
        const int col0 = n0 + c;
        const int ncols = min(32, p.cout - col0);  // <= 0 for padded weight rows of the last N tile
        // global loads for this chunk (bias) and the next (residual) go out before the TMEM wait
        uint4 rcur[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
        if (c + CHUNK_COLS < cend) {
          fetch_res(col0 + CHUNK_COLS);
          if (res_tma && lane == 0) {
            // the other buffer was last read (generic proxy) one chunk ago, before a __syncwarp
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(res_bar0 + 8u * (res_buf ^ 1), 2048);
            tma_load_2d(epi_base + 2048u * (res_buf ^ 1), &p.tmRes, res_bar0 + 8u * (res_buf ^ 1), col0 + CHUNK_COLS, row0);
          }
        }
        float4 bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && ncols == 32) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_row + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(b4 + j);
        }
        tmem_wait_ld();
        if (CHUNK_COLS != 32) {
#pragma unroll
          for (int i = 16; i < 32; ++i) v[i] = 0;
        }
        if (res_tma) {
          // consume this chunk's residual buffer even when the chunk is padding, to keep phases in step
          mbar_wait(res_bar0 + 8u * res_buf, (res_phase >> res_buf) & 1u);
          res_phase ^= 1u << res_buf;
        }
        const int cur_buf = res_buf;
        res_buf ^= (res_tma ? 1 : 0);
        if (ncols <= 0) continue;  // warp-uniform
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bv[j].x;
          f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bv[j].y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bv[j].z;
          f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bv[j].w;
        }
        if (p.bias != nullptr && ncols != 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) f[i] += __ldg(bias_row + col0 + i);
        }
        if (p.out_mode == ADB_OUT_BF16_NHWC) {
          // residual: 8-channel vectors lie fully inside cout (cout % 8 == 0)
          if (res_tma) {
            // own row of the landed 32x32 residual sub-tile (64-byte swizzle: chunk ^= (row >> 1) & 3)
            const uint8_t* rrow = epi_gen + 2048 * cur_buf + lane * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g) rcur[g] = *reinterpret_cast<const uint4*>(rrow + ((g ^ ((lane >> 1) & 3)) << 4));
          }
          if (res_pf || res_tma) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              f[g * 8 + 0] += bf16_lo(rcur[g].x);
              f[g * 8 + 1] += bf16_hi(rcur[g].x);
              f[g * 8 + 2] += bf16_lo(rcur[g].y);
              f[g * 8 + 3] += bf16_hi(rcur[g].y);
              f[g * 8 + 4] += bf16_lo(rcur[g].z);
              f[g * 8 + 5] += bf16_hi(rcur[g].z);
              f[g * 8 + 6] += bf16_lo(rcur[g].w);
              f[g * 8 + 7] += bf16_hi(rcur[g].w);
            }
          } else if (p.res_mode == ADB_RES_AVGPOOL2 && row_ok) {
            for (int sidx = 0; sidx < 4; ++sidx) {
              const int sy = 2 * y + (sidx >> 1), sx = 2 * x + (sidx & 1);
              const size_t pix = ((size_t)img * (2 * p.H) + sy) * (2 * p.W) + sx;
              const __nv_bfloat16* rp = p.residual + pix * p.cout + col0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (g * 8 < ncols) {
                  const uint4 r = __ldg(reinterpret_cast<const uint4*>(rp) + g);
                  f[g * 8 + 0] += 0.25f * bf16_lo(r.x);
                  f[g * 8 + 1] += 0.25f * bf16_hi(r.x);
                  f[g * 8 + 2] += 0.25f * bf16_lo(r.y);
                  f[g * 8 + 3] += 0.25f * bf16_hi(r.y);
                  f[g * 8 + 4] += 0.25f * bf16_lo(r.z);
                  f[g * 8 + 5] += 0.25f * bf16_hi(r.z);
                  f[g * 8 + 6] += 0.25f * bf16_lo(r.w);
                  f[g * 8 + 7] += 0.25f * bf16_hi(r.w);
                }
              }
            }
          }
          uint32_t ow[16];  // this row's 32 outputs rounded to bf16, packed in pairs (zeros past M / cout)
#pragma unroll
          for (int i = 0; i < 16; ++i)
            ow[i] = (row_ok && 2 * i < ncols) ? pack_bf16x2(f[2 * i], f[2 * i + 1]) : 0u;
          if (p.tma_epi) {
            // stage the warp's 32x32 bf16 sub-tile in shared memory (conflict-free with the 64-byte
            // swizzle) and let ONE TMA store write it: row-strided per-thread 16-byte stores cost 32 LSU
            // wavefronts per instruction and made small-K tiles epilogue-bound.
            if (lane == 0) tma_store_wait_read0();  // previous store has finished reading the buffer
            __syncwarp();
            uint8_t* orow = epi_out_gen + lane * 64;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(orow + ((g ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(ow[4 * g], ow[4 * g + 1], ow[4 * g + 2], ow[4 * g + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmOut, epi_out, col0, row0);  // rows >= M and columns >= cout are clipped
              tma_store_commit();
            }
          } else if (row_ok) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.cout + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g * 8 < ncols)
                reinterpret_cast<uint4*>(op)[g] = make_uint4(ow[4 * g], ow[4 * g + 1], ow[4 * g + 2], ow[4 * g + 3]);
            }
          }
          if (p.stats != nullptr) {
            // GroupNorm statistics of the tensor being written, for its consumer (nn.py:17-19): sum and
            // sum of squares of the STORED (bf16-rounded) values per (image, group of cpg channels).
            // Column sums over the warp's 32 rows (pixels of one image: P % 32 == 0): lane j owns column
            // col0 + j (read back from the staged tile, or by a transposing shuffle butterfly when the
            // tile is not staged). The columns go through a per-warp smem buffer and ONE lane per group
            // adds its group's columns into the warp's bin (no atomics: shared fp32 atomicAdd is a
            // CAS loop that serialises cpg-fold on these addresses).
            float sv[32], qv[32];
            if (p.tma_epi) {
              // the rounded tile is already staged in shared memory for the TMA store: lane j walks
              // column j down the 32 rows (a row is 64 contiguous bytes: conflict-free)
              float cs = 0.f, cq = 0.f;
              const uint32_t cch = (uint32_t)lane >> 3, cin8 = ((uint32_t)lane & 7u) * 2u;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint16_t hv = *reinterpret_cast<const uint16_t*>(
                    epi_out_gen + r * 64 + (((cch ^ ((uint32_t)(r >> 1) & 3u)) << 4) + cin8));
                const float val = __uint_as_float((uint32_t)hv << 16);
                cs += val;
                cq = fmaf(val, val, cq);
              }
              sv[0] = cs;
              qv[0] = cq;
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float val = (i & 1) ? bf16_hi(ow[i >> 1]) : bf16_lo(ow[i >> 1]);
                sv[i] = val;
                qv[i] = val * val;
              }
#pragma unroll
              for (int off = 16, cnt = 16; off > 0; off >>= 1, cnt >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < cnt; ++i) {
                  const float s_keep = upper ? sv[i + cnt] : sv[i];
                  const float s_send = upper ? sv[i] : sv[i + cnt];
                  const float q_keep = upper ? qv[i + cnt] : qv[i];
                  const float q_send = upper ? qv[i] : qv[i + cnt];
                  sv[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
                  qv[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
                }
              }
            }
            my_cols[lane] = sv[0];
            my_cols[32 + lane] = qv[0];
            __syncwarp();
            const int gfirst = col0 / p.cpg;
            const int gmine = gfirst + lane;                  // lane t sums group gfirst + t
            const int lo = max(gmine * p.cpg, col0) - col0;   // its columns inside this chunk
            const int hi = min((gmine + 1) * p.cpg, col0 + ncols) - col0;
            if (lo < hi) {
              float gs = 0.f, gq = 0.f;
              for (int j = lo; j < hi; ++j) {
                gs += my_cols[j];
                gq += my_cols[32 + j];
              }
              my_bins[2 * (gmine - g_lo)] += gs;
              my_bins[2 * (gmine - g_lo) + 1] += gq;
            }
            if (p.stats2 != nullptr) {
              const int c2 = p.choff2 + col0;  // this chunk's first column in the consumer's concat
              const int g2 = c2 / p.cpg2 + lane;
              const int lo2 = max(g2 * p.cpg2, c2) - c2;
              const int hi2 = min((g2 + 1) * p.cpg2, c2 + ncols) - c2;
              if (lo2 < hi2) {
                float gs = 0.f, gq = 0.f;
                for (int j = lo2; j < hi2; ++j) {
                  gs += my_cols[j];
                  gq += my_cols[32 + j];
                }
                my_bins2[2 * (g2 - g_lo2)] += gs;
                my_bins2[2 * (g2 - g_lo2) + 1] += gq;
              }
            }
            __syncwarp();
          }
        } else if (row_ok) {
          // fp32 NCHW: for a fixed channel, the warp's 32 pixels are contiguous along x
          float* op = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < ncols) op[((size_t)img * p.cout + (col0 + i)) * P + rem] = f[i];
          }
        }
      }
      }  // generic path
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2 && !is_leader) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
      if (p.stats != nullptr && cbeg < cend) {
        // flush this warp's bins: one fp64 atomic pair per group touched by this tile half
        const int col_last = min(n0 + cend, p.cout) - 1;
        const int ng = (col_last >= n0 + cbeg) ? (fast_div(col_last, p.cpg_magic) - g_lo + 1) : 0;
        const int img0 = __shfl_sync(0xffffffffu, img, 0);
        const bool any_row = __shfl_sync(0xffffffffu, (int)row_ok, 0) != 0;  // rows ascend: row 0 invalid => all invalid
        for (int gb = lane; gb < ng; gb += 32) {
          const float bs = my_bins[2 * gb], bq = my_bins[2 * gb + 1];
          my_bins[2 * gb] = 0.f;
          my_bins[2 * gb + 1] = 0.f;
          if (any_row) {
            double* sp = p.stats + ((size_t)img0 * 32 + (g_lo + gb)) * 2;
            atomicAdd(sp, (double)bs);
            atomicAdd(sp + 1, (double)bq);
          }
        }
        if (p.stats2 != nullptr) {
          const int ng2 = (col_last >= n0 + cbeg) ? (fast_div(p.choff2 + col_last, p.cpg2_magic) - g_lo2 + 1) : 0;
          for (int gb = lane; gb < ng2; gb += 32) {
            const float bs = my_bins2[2 * gb], bq = my_bins2[2 * gb + 1];
            my_bins2[2 * gb] = 0.f;
            my_bins2[2 * gb + 1] = 0.f;
            if (any_row) {
              double* sp = p.stats2 + ((size_t)img0 * 32 + (g_lo2 + gb)) * 2;
              atomicAdd(sp, (double)bs);
              atomicAdd(sp + 1, (double)bq);
            }
          }
        }
        __syncwarp();
      }
    }
  }

  if (warp < 8 && lane == 0 && p.tma_epi) tma_store_wait_all0();  // output stores of this warp are complete
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();  // the peer may still be reading our smem / TMEM
  if (warp == MMA_WARP) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N, int NCTA, bool GNB = false>
int launch(const ConvKParams& kp, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, NCTA>;
  static bool attr_set = false;
  if (!attr_set) {
    ADB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, NCTA, GNB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int items = ((kp.m_tiles + NCTA - 1) / NCTA) * kp.n_tiles;
  int grid = items * NCTA < num_sms() ? items * NCTA : num_sms();
  grid -= grid % NCTA;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ADB_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BLOCK_N, NCTA, GNB>, kp));
  return 1;
}

}  // namespace

// N tile for a given real Cout; weights are padded to a multiple of it.
int conv_block_n(int cout) {
  if (cout % 192 == 0) return 192;
  if (cout % 256 == 0) return 256;
  // 320-channel layers (Stable Diffusion): two 160-wide CTA-pair tiles cover them exactly; the 192-wide tile would
  // compute 384 columns (17 % wasted MMA work). ADB_CONV_NO160=1 restores the padded tiling (A/B testing).
  if (cout % 160 == 0 && cout % 128 != 0) {
    static int no160 = -1;
    if (no160 < 0) {
      const char* e = getenv("ADB_CONV_NO160");
      no160 = (e && e[0] == '1') ? 1 : 0;
    }
    if (!no160) return 160;
  }
  if (cout % 128 == 0) return 128;
  if (cout <= 16) return 16;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  return 192;
}

namespace {

// CTA pairs (cta_group::2) for the 192- and 256-wide N tiles unless ADB_CONV_1CTA=1 (A/B testing).
// Measured: a 128-wide pair tile is ~18 % SLOWER than the single-CTA 128 tile (the pair's handshakes buy only
// 8 KB less B traffic per stage). Round 1 fetched the 256-wide pair tile's B half (128 weight rows) as TWO
// 64-row boxes because a single cta_group::2 box of 112 or 128 rows "never completed its mbarrier transaction" in the
// kernel of that time. Chased in round 2: scripts/microbench/tma_2sm_box_rows.cu runs the same ring protocol stand-alone
// (64 / 96 / 112 / 128-row boxes, 1-4 stages, 1 or 74 clusters) and every load completes; with the present kernel the
// single box passes the parity tests and is 3 % faster (1490 -> 1531 TFLOP/s at 32x32 256 -> 256), so it is the default now
// (ADB_CONV_B128BOX=0 restores the two boxes). The round-1 stall came from that version's barrier protocol, not from TMA.
int conv_ncta(int block_n) {
  static int force1 = -1;
  if (force1 < 0) {
    const char* e = getenv("ADB_CONV_1CTA");
    force1 = (e && e[0] == '1') ? 1 : 0;
  }
  if (block_n == 256 || block_n == 160) return 2;
  return (block_n == 192 && !force1) ? 2 : 1;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

// the N tile conv_igemm_submit ends up with for an [M, cout_pad] output
static int effective_block_n(long long M, int cout, int cout_pad) {
  int block_n = conv_block_n(cout);
  // the 256-wide pair tile pays once there is at least one wave of them; small problems keep the 128-wide tile
  // (cout_pad is a multiple of 256, hence of 128 as well)
  if (block_n == 256 && ((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (cout_pad / 256) < num_sms() / 2) block_n = 128;
  return block_n;
}

int conv_gnb_supported(int n, int h, int w, int cout) {
  const long long P = (long long)h * w, M = P * n;
  if (n <= 0 || h <= 0 || w <= 0 || cout <= 0 || cout % 32 != 0 || P % 32 != 0 || M % BLOCK_M != 0) return 0;
  const int bn = conv_block_n(cout);
  if (cout % bn != 0) return 0;  // cout_pad == cout
  const int be = effective_block_n(M, cout, cout);
  return (be >= 64 && cout % be == 0) ? 1 : 0;
}

int conv_igemm_submit(adb_plan* plan, const adb_conv_desc* d, cudaStream_t stream) {
  ADB_REQUIRE(d != nullptr, "conv_igemm: null descriptor");
  ADB_REQUIRE(d->n > 0 && is_pow2(d->h) && is_pow2(d->w), "conv_igemm: n>0 and power-of-two h,w required (n=%d h=%d w=%d)", d->n, d->h, d->w);
  ADB_REQUIRE(d->nseg >= 1 && d->nseg <= 3, "conv_igemm: nseg must be 1..3");
  ADB_REQUIRE(d->cout > 0 && d->cout_pad >= d->cout && d->cout_pad % 16 == 0, "conv_igemm: cout_pad must be a multiple of 16 and >= cout");
  ADB_REQUIRE(d->weight && d->out, "conv_igemm: null weight/out");
  ADB_REQUIRE(d->out_mode == ADB_OUT_BF16_NHWC || d->out_mode == ADB_OUT_F32_NCHW, "conv_igemm: bad out_mode");
  if (d->out_mode == ADB_OUT_BF16_NHWC)
    ADB_REQUIRE(d->cout % 8 == 0, "conv_igemm: bf16 NHWC output needs cout %% 8 == 0");
  ADB_REQUIRE(d->res_mode == ADB_RES_NONE || (d->residual != nullptr && d->out_mode == ADB_OUT_BF16_NHWC),
              "conv_igemm: residual requires a source and bf16 output");
  if (d->res_mode == ADB_RES_NEAREST2) ADB_REQUIRE(d->h >= 2 && d->w >= 2, "conv_igemm: nearest2 needs h,w >= 2");

  ConvKParams kp;
  memset(&kp, 0, sizeof(kp));
  const long long P = (long long)d->h * d->w;
  const long long M = P * d->n;
  ADB_REQUIRE(M < (1ll << 31), "conv_igemm: too many pixels");
  // box geometry: 128 rows of the GEMM = bn images x bh rows x bw columns
  const int bw = d->w < BLOCK_M ? d->w : BLOCK_M;
  const int bh = (P < BLOCK_M) ? d->h : (BLOCK_M / bw);
  const int bn = (P < BLOCK_M) ? (int)(BLOCK_M / P) : 1;
  ADB_REQUIRE(bw * bh * bn == BLOCK_M, "conv_igemm: cannot tile %dx%d images into 128-row boxes", d->h, d->w);

  int ktot = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const adb_conv_seg& sg = d->seg[s];
    ADB_REQUIRE(sg.act != nullptr && sg.cin > 0 && sg.cin % 8 == 0, "conv_igemm: segment %d needs cin %% 8 == 0", s);
    ADB_REQUIRE(sg.taps == 1 || sg.taps == 9, "conv_igemm: taps must be 1 or 9");
    kp.seg_taps[s] = sg.taps;
    kp.seg_cin[s] = sg.cin;
    kp.seg_chunks[s] = (sg.cin + BLOCK_K - 1) / BLOCK_K;
    kp.seg_kbase[s] = ktot;
    ktot += sg.taps * sg.cin;
    kp.k_iters += sg.taps * kp.seg_chunks[s];
    kp.seg_stride[s] = d->seg_stride[s] == 2 ? 2 : 1;
    int r;
    if (kp.seg_stride[s] == 2) {
      ADB_REQUIRE(sg.taps == 9 && sg.cin % BLOCK_K == 0, "conv_igemm: a stride-2 segment needs taps = 9 and cin %% 64 == 0");
      // input [n, 2h, 2w, cin] viewed as [n, h, 2, w, 2*cin]: a pixel pair along x is one row of 2*cin channels
      const uint64_t dims[5] = {(uint64_t)2 * sg.cin, (uint64_t)d->w, 2, (uint64_t)d->h, (uint64_t)d->n};
      const uint64_t strides[4] = {(uint64_t)2 * sg.cin * 2, (uint64_t)2 * d->w * sg.cin * 2,
                                   (uint64_t)4 * d->w * sg.cin * 2, (uint64_t)4 * P * sg.cin * 2};
      const uint32_t box[5] = {(uint32_t)BLOCK_K, (uint32_t)bw, 1u, (uint32_t)bh, (uint32_t)bn};
      r = make_tmap_bf16(&kp.tmA[s], sg.act, 5, dims, strides, box);
    } else {
      const uint64_t dims[4] = {(uint64_t)sg.cin, (uint64_t)d->w, (uint64_t)d->h, (uint64_t)d->n};
      const uint64_t strides[3] = {(uint64_t)sg.cin * 2, (uint64_t)d->w * sg.cin * 2,
                                   (uint64_t)P * sg.cin * 2};
      const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
      r = make_tmap_bf16(&kp.tmA[s], sg.act, 4, dims, strides, box);
    }
    if (r != ADB_OK) return r;
  }
  const int block_n = effective_block_n(M, d->cout, d->cout_pad);
  const int ncta = conv_ncta(block_n);
  ADB_REQUIRE(d->cout_pad % block_n == 0, "conv_igemm: cout_pad (%d) must be a multiple of the N tile %d (adb_conv_block_n)", d->cout_pad, block_n);
  {
    const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->cout_pad};
    const uint64_t strides[1] = {(uint64_t)ktot * 2};
    static int b128 = -1;
    if (b128 < 0) {
      const char* e = getenv("ADB_CONV_B128BOX");
      b128 = (e && e[0] == '0') ? 0 : 1;
    }
    kp.b_one_box = b128;
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)(block_n == 256 && !b128 ? 64 : block_n / ncta)};
    int r = make_tmap_bf16(&kp.tmW, d->weight, 2, dims, strides, box);
    if (r != ADB_OK) return r;
  }
  kp.tma_epi = (d->out_mode == ADB_OUT_BF16_NHWC && block_n >= 32) ? 1 : 0;
  {
    static int no_tma_epi = -1;
    if (no_tma_epi < 0) {
      const char* e = getenv("ADB_CONV_NO_TMA_EPI");
      no_tma_epi = (e && e[0] == '1') ? 1 : 0;
    }
    if (no_tma_epi) kp.tma_epi = 0;
  }
  if (kp.tma_epi) {
    const uint64_t dims[2] = {(uint64_t)d->cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)d->cout * 2};
    const uint32_t box[2] = {32u, 32u};
    int r = make_tmap_bf16(&kp.tmOut, d->out, 2, dims, strides, box, 64);
    if (r != ADB_OK) return r;
    if (d->res_mode == ADB_RES_SAME) {
      r = make_tmap_bf16(&kp.tmRes, d->residual, 2, dims, strides, box, 64);
      if (r != ADB_OK) return r;
    }
  }
  kp.nseg = d->nseg;
  kp.M = (int)M;
  kp.cout = d->cout;
  kp.H = d->h;
  kp.W = d->w;
  kp.m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M);
  kp.n_tiles = (d->cout_pad + block_n - 1) / block_n;
  kp.bias = d->bias;
  kp.bias_stride = d->bias != nullptr ? d->bias_stride : 0;
  ADB_REQUIRE(kp.bias_stride == 0 || (kp.bias_stride >= d->cout && kp.bias_stride % 4 == 0 && P % 32 == 0),
              "conv_igemm: per-image bias needs bias_stride >= cout, %% 4 == 0 and h*w %% 32 == 0");
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  kp.res_mode = d->res_mode;
  kp.out = d->out;
  kp.out_mode = d->out_mode;
  kp.stats = d->stats_out;
  kp.cpg = d->cout / 32;
  kp.stats2 = d->stats2_out;
  kp.cpg2 = d->stats2_cpg;
  kp.choff2 = d->stats2_choff;
  kp.cpg_magic = kp.cpg > 0 ? ((1ull << 32) + (unsigned long long)kp.cpg - 1) / (unsigned long long)kp.cpg : 0;
  kp.cpg2_magic = kp.cpg2 > 0 ? ((1ull << 32) + (unsigned long long)kp.cpg2 - 1) / (unsigned long long)kp.cpg2 : 0;
  if (d->stats2_out != nullptr) {
    ADB_REQUIRE(d->stats_out != nullptr && d->stats2_cpg > 0 && d->stats2_choff >= 0 &&
                    d->stats2_choff + d->cout <= 32 * d->stats2_cpg && 96 / d->stats2_cpg + 2 <= STAT_BINS,
                "conv_igemm: stats2 needs stats_out and a concat grouping that contains this tensor");
  }
  if (d->stats_out != nullptr) {
    ADB_REQUIRE(d->out_mode == ADB_OUT_BF16_NHWC && d->cout % 32 == 0 && P % 32 == 0,
                "conv_igemm: stats_out needs bf16 output, cout %% 32 == 0 and h*w %% 32 == 0");
  }

  if (d->gnb_bstats != nullptr) {
    // GroupNorm-backward sums in the epilogue: only on the all-TMA epilogue path (full tiles, one image per 32-row slab)
    ADB_REQUIRE(d->gnb_x && d->gnb_stats && d->gnb_gamma && d->gnb_beta, "conv_igemm: gnb needs x, stats, gamma and beta");
    ADB_REQUIRE(kp.tma_epi && d->res_mode == ADB_RES_NONE && d->stats_out == nullptr && d->stats2_out == nullptr &&
                    M % BLOCK_M == 0 && P % 32 == 0 && d->cout % 32 == 0 && d->cout_pad == d->cout && d->cout % block_n == 0 &&
                    block_n >= 64,
                "conv_igemm: gnb needs bf16 output without residual or forward statistics, full tiles (n*h*w %% 128 == 0, "
                "cout %% N tile == 0) and h*w %% 32 == 0");
    const uint64_t dims[2] = {(uint64_t)d->cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)d->cout * 2};
    const uint32_t box[2] = {32u, 32u};
    int r = make_tmap_bf16(&kp.tmRes, d->gnb_x, 2, dims, strides, box, 64);
    if (r != ADB_OK) return r;
    kp.res_mode = ADB_RES_SAME;
    kp.gnb = 1;
    kp.gnb_inv_cnt = 1.0 / ((double)(d->cout / 32) * (double)P);
    kp.stats = d->gnb_bstats;
    kp.gnb_stats = d->gnb_stats;
    kp.gnb_gamma = d->gnb_gamma;
    kp.gnb_beta = d->gnb_beta;
    kp.gnb_ss = d->gnb_scale_shift;
    kp.gnb_ss_stride = d->gnb_ss_stride;
    kp.gnb_eps = d->gnb_eps;
    kp.gnb_silu = d->gnb_silu;
  }

  const double flops = 2.0 * (double)M * (double)d->cout * (double)ktot;
  const size_t gnb_bytes = kp.gnb ? (size_t)d->n * 32 * 2 * sizeof(double) : 0;
  // the GNB instantiation is reported under its own name: its epilogue carries the consumer GroupNorm's backward sums
  return submit(plan, stream, kp.gnb ? "conv_igemm_gnb" : "conv_igemm", flops, 0.0, [kp, block_n, ncta, gnb_bytes](cudaStream_t s) -> int {
    if (gnb_bytes) {
      ADB_CUDA(cudaMemsetAsync(kp.stats, 0, gnb_bytes, s));
      switch (block_n) {
        case 256: return launch<256, 2, true>(kp, s);
        case 192: return ncta == 2 ? launch<192, 2, true>(kp, s) : launch<192, 1, true>(kp, s);
        case 160: return launch<160, 2, true>(kp, s);
        case 128: return launch<128, 1, true>(kp, s);
        default: return launch<64, 1, true>(kp, s);
      }
    }
    switch (block_n) {
      case 256: return launch<256, 2>(kp, s);
      case 192: return ncta == 2 ? launch<192, 2>(kp, s) : launch<192, 1>(kp, s);
      case 160: return launch<160, 2>(kp, s);
      case 128: return launch<128, 1>(kp, s);
      case 64: return launch<64, 1>(kp, s);
      default: return launch<16, 1>(kp, s);
    }
  });
}

}  // namespace adb
