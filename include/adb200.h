/*
 * adb200.h — C-ABI of libadb200.so, the B200 (sm_100a) compute library under the
 * AutoDiffusion candidate evaluator.
 *
 * Plain pointers and sizes only: no torch types cross this boundary. Every device
 * pointer is borrowed for the duration of the call (or of the plan, for recorded
 * ops). All functions return 0 on success and a negative adb_status on failure;
 * adb_last_error() gives the message. Nothing here throws, exits or falls back to
 * the CPU.
 *
 * Layout conventions inside the library:
 *   - "act" tensors are bf16 NHWC: [n, h, w, c], c contiguous (pixels are GEMM rows).
 *   - weights for the implicit GEMM are bf16 [cout_pad, ktot], K contiguous, where
 *     K runs over (segment, tap=kh*3+kw, cin) in that order.
 *   - API-edge tensors (x_t, eps, samples) are fp32 NCHW exactly as the reference's.
 *
 * Every op takes `adb_plan* plan` first: NULL executes immediately on `stream`;
 * non-NULL records the op (with its TMA descriptors encoded once) for adb_plan_run.
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to /root/reference/examples/guided_diffusion/).
 */
#ifndef ADB200_H
#define ADB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* adb_stream;          /* cudaStream_t */
typedef struct adb_plan adb_plan;  /* opaque recorded op list */

enum adb_status {
  ADB_OK = 0,
  ADB_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  ADB_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  ADB_ERR_UNSUPPORTED = -3  /* device is not sm_100 */
};

const char* adb_last_error(void);
int adb_version(void);
/* 0 if the current device can run this library (compute capability 10.x). */
int adb_device_check(void);

/* ---- plans: recorded launch schedules (guided_diffusion/dynamic_unet.py:673-702 is
 * walked once per (batch, skip-mask); skipped blocks never enter the list) ---- */
adb_plan* adb_plan_create(void);
void adb_plan_destroy(adb_plan* plan);
int adb_plan_num_ops(const adb_plan* plan);
/* launches every recorded op in order on `stream`; returns #kernels launched (>=0) or <0 */
int adb_plan_run(adb_plan* plan, adb_stream stream);
/* what op i is ("conv_igemm", "groupnorm", ...) and its algorithmic work: flops = 2*MACs of
 * GEMM-shaped work, bytes = compulsory HBM bytes of memory-bound work (roofline numerators) */
int adb_plan_op_info(const adb_plan* plan, int i, const char** kind, double* flops, double* bytes);
/* like adb_plan_run, with a CUDA event pair around each op on `stream`; writes per-op device
 * milliseconds to ms_out[0..num_ops) and synchronises the stream. Returns num_ops or <0. */
int adb_plan_run_profiled(adb_plan* plan, adb_stream stream, float* ms_out, int capacity);

/* ---- implicit-GEMM convolution / k=1 GEMM on tcgen05+TMEM, operands by TMA ----
 * replaces nn.Conv2d 3x3 s1 p1, nn.Conv2d 1x1 and nn.Conv1d k=1
 * (guided_diffusion/nn.py:22-32 via dynamic_unet.py:194,220,231,303,311,502,653).
 * out[m, co] = bias[co] + sum_seg sum_tap sum_ci act_seg[n, h+dh, w+dw, ci] * W[co, k]
 *              (+ residual), zero padding, m = (n*h + y)*w + x.
 * Up to three K-segments let the ResBlock's second conv absorb the 1x1 skip
 * connection over the (never materialised) channel concat (dynamic_unet.py:271,699). */
enum { ADB_RES_NONE = 0, ADB_RES_SAME = 1, ADB_RES_AVGPOOL2 = 2, ADB_RES_NEAREST2 = 3 };
enum { ADB_OUT_BF16_NHWC = 0, ADB_OUT_F32_NCHW = 1 };

typedef struct {
  const void* act;  /* bf16 NHWC [n, h, w, cin] */
  int cin;          /* multiple of 8 */
  int taps;         /* 1 (1x1) or 9 (3x3, pad 1) */
} adb_conv_seg;

typedef struct {
  int n, h, w;          /* output geometry, h and w powers of two */
  int cout;             /* real output channels */
  int cout_pad;         /* rows of `weight`: cout rounded up to adb_conv_block_n(cout), zero rows */
  int nseg;             /* 1..3 */
  adb_conv_seg seg[3];
  const void* weight;   /* bf16 [cout_pad, ktot] */
  const float* bias;    /* fp32 [cout] or NULL */
  const void* residual; /* bf16 NHWC, geometry per res_mode, cout channels; or NULL */
  int res_mode;
  void* out;
  int out_mode;
  double* stats_out;    /* optional [n, 32, 2] fp64: the epilogue ADDS sum / sum-of-squares of the stored
                           output per (image, GroupNorm group of cout/32 channels); caller zeroes it */
  double* stats2_out;   /* optional second target (requires stats_out): the GroupNorm of a consumer that reads
                           this tensor as channels [stats2_choff, +cout) of a channel concat
                           (dynamic_unet.py:699) with groups of stats2_cpg channels */
  int stats2_cpg;
  int stats2_choff;
  /* ---- appended for the Stable-Diffusion UNet ("Stable Diffusion"/ldm/modules/diffusionmodules/openaimodel.py) ---- */
  int bias_stride;      /* 0: bias[co]. > 0: per-image bias rows, bias[img * bias_stride + co] - the ResBlock's
                           `h = h + emb_out[..., None, None]` (openaimodel.py:272) folded into the first conv's
                           epilogue (conv bias pre-added into the embedding Linear's bias) */
  int seg_stride[3];    /* 0 or 1: unit stride. 2: segment s is a 3x3 stride-2 pad-1 convolution (Downsample.op,
                           openaimodel.py:147-149): its `act` is [n, 2h, 2w, cin], output pixel (y, x) reads input
                           rows 2y-1..2y+1, columns 2x-1..2x+1 */
  /* ---- appended for classifier guidance: this conv is a DATA GRADIENT whose output dY feeds adb_gn_backward of a
   * GroupNorm with forward input gnb_x [n,h,w,cout]. With gnb_x != NULL the epilogue also reduces that backward's two
   * sums per (image, group) - sum dxh and sum dxh*xh with xh = (x - mean) rstd, dxh = dY act'(z) gamma (1+scale) - into
   * gnb_bstats (zeroed here), so that adb_gn_backward (bstats_ready = 1) reads x and dY once instead of twice.
   * Requires bf16 output, no residual, n*h*w % 128 == 0, h*w % 32 == 0, cout % adb_conv_block_n(cout) == 0. ---- */
  const void* gnb_x;           /* bf16 NHWC [n,h,w,cout] or NULL */
  const double* gnb_stats;     /* forward sums of gnb_x [n,32,2] */
  const float* gnb_gamma;      /* [cout] */
  const float* gnb_beta;       /* [cout] */
  const float* gnb_scale_shift; /* optional FiLM rows (scale | shift), row stride gnb_ss_stride floats */
  int gnb_ss_stride;
  float gnb_eps;
  int gnb_silu;
  double* gnb_bstats;          /* out [n,32,2] */
} adb_conv_desc;

/* N tile the kernel uses for `cout` output channels; `cout_pad` must be a multiple of it. */
int adb_conv_block_n(int cout);
/* 1 when adb_conv_igemm accepts gnb_* for an [n,h,w,cout] output (full tiles on the all-TMA epilogue), else 0 */
int adb_conv_gnb_supported(int n, int h, int w, int cout);
int adb_conv_igemm(adb_plan* plan, const adb_conv_desc* d, adb_stream stream);

/* ---- fused softmax attention, head dim 64 ----
 * replaces QKVAttention.forward / QKVAttentionLegacy.forward
 * (dynamic_unet.py:390-409 / 357-374): softmax_fp32((q*s)^T (k*s)) v, s = 64^-1/4.
 * qkv: bf16 [b*t, 3*heads*64] (the k=1 conv output, pixels as rows)
 * out: bf16 [b*t, heads*64]
 * legacy_order = 0: channels [q(all heads) | k | v]; 1: per head [q k v]. */
int adb_attention(adb_plan* plan, const void* qkv, void* out, int b, int t, int heads,
                  int legacy_order, adb_stream stream);

/* ---- GroupNorm(32) (+scale-shift) (+SiLU) (+2x resample) ----
 * replaces GroupNorm32.forward + nn.SiLU + the FiLM line + Upsample/Downsample on h
 * (nn.py:17-19, dynamic_unet.py:192-193,216-217,253-254,262-265,302,651-652).
 * The input may be the channel concat of two NHWC tensors (dynamic_unet.py:699). */
enum { ADB_RESAMPLE_NONE = 0, ADB_RESAMPLE_AVGPOOL2 = 1, ADB_RESAMPLE_NEAREST2 = 2 };

typedef struct {
  int n, h, w;              /* input geometry */
  const void* src0; int c0; /* bf16 NHWC */
  const void* src1; int c1; /* optional second source (channel concat), c1 = 0 if none */
  const float* gamma;       /* [c0+c1] */
  const float* beta;        /* [c0+c1] */
  float eps;
  const float* scale_shift; /* optional fp32 rows: scale = row[0:c], shift = row[c:2c] */
  int ss_stride;            /* floats between consecutive samples' rows */
  int silu;
  int resample;
  void* out;                /* bf16 NHWC at the resampled geometry, c0+c1 channels */
  double* stats;            /* [n, 32, 2] doubles: (sum, sum of squares) per (image, group) */
  int stats_ready;          /* 0: zero `stats` and compute them here (one extra read of the input);
                               1: `stats` were already accumulated by the producer (adb_conv_igemm
                               stats_out / adb_stem_conv) - only the normalise pass runs */
} adb_gn_desc;

int adb_groupnorm(adb_plan* plan, const adb_gn_desc* d, adb_stream stream);

/* 2x average pool / nearest upsample of a bf16 NHWC tensor
 * (x_upd of a *skipped* up/down ResBlock, dynamic_unet.py:246-249). */
int adb_resample2x(adb_plan* plan, const void* src, void* dst, int n, int h, int w, int c,
                   int mode, adb_stream stream);

/* ---- input stem: fp32 NCHW [n,cin,h,w] -> conv3x3 -> bf16 NHWC [n,h,w,cout]
 * (input_blocks.0.0, dynamic_unet.py:501-503, with x.type(dtype) at :693).
 * weight fp32 [cout, cin, 3, 3] (PyTorch layout), bias fp32 [cout]; cin <= 4. */
int adb_stem_conv(adb_plan* plan, const float* x, const float* weight, const float* bias,
                  void* out, int n, int cin, int h, int w, int cout, adb_stream stream);

/* ---- timestep / label embedding path (nn.py:103-121, dynamic_unet.py:490-498,687-691)
 * sinusoid: out[b, :] = [cos(t_b f_k) | sin(t_b f_k)]; freqs[k] = exp(-ln(1e4) k / (dim/2)) is a
 * constant fp32 table of dim/2 entries supplied by the host. */
int adb_timestep_embedding(adb_plan* plan, const int64_t* t, const float* freqs, float* out, int b,
                           int dim, adb_stream stream);
/* out[b, j] = bias[j] + sum_k act(x[b,k]) W[j,k] (+ table[idx[b], j]);
 * act = SiLU if silu_in else identity. fp32 throughout (the reference keeps these fp32,
 * fp16_util.py:15-22). Covers time_embed.{0,2}, label_emb add and all ResBlock
 * emb_layers batched into one [b,768]x[768, sum 2*cout] product. */
int adb_linear(adb_plan* plan, const float* x, const float* w, const float* bias, float* out,
               int b, int k, int nout, int silu_in, const float* table, const int64_t* idx,
               adb_stream stream);

/* Tensor-core form of adb_stem_conv for cin == 3: writes the 3x3 zero-padded neighbourhood of every pixel as a
 * bf16 row [27 values (tap-major, channel-minor) | 5 zeros | their 27 bf16 rounding residuals | 5 zeros] of
 * out [n,h,w,64]; adb_conv_igemm over it (taps = 1, cin = 64) with the weight [cout, 27|0|27|0] is the stem conv. */
int adb_stem_im2col(adb_plan* plan, const float* x, void* out, int n, int h, int w, adb_stream stream);

/* hi = bf16(act(x)), lo = bf16(act(x) - hi), act = SiLU if silu_in else identity; x fp32 [total].
 * Feeds the wide emb_layers product (dynamic_unet.py:208-214,259: [B,768] x [768, sum 2*cout]) to adb_conv_igemm as
 * three K-segments [hi | lo | hi] x [W_hi | W_hi | W_lo] with fp32 output: an fp32-grade Linear on tensor cores. */
int adb_split_bf16(adb_plan* plan, const float* x, void* hi, void* lo, size_t total, int silu_in, adb_stream stream);

/* ---- fused guidance + DDIM update (gaussian_diffusion.py:328-349,371-393,536-584)
 * coef = {sqrt_recip_acp, sqrt_recipm1_acp, sqrt(1-acp), sqrt(acp_prev), sqrt(1-acp_prev)}
 * (fp32, as the reference rounds them at gather, :920). eps is read from the first 3 of
 * eps_channels channels of model_out. grad may be NULL (no cond_fn). eta = 0 only. */
int adb_ddim_step(adb_plan* plan, const float* x, const float* model_out, int eps_channels,
                  const float* grad, float* x_prev, float* pred_xstart /* may be NULL */,
                  int n, int c, int hw, const float coef[5], int clip_denoised,
                  adb_stream stream);

/* ((s+1)*127.5).clamp(0,255).to(uint8) NCHW -> NHWC
 * (search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:421-423). */
int adb_pack_uint8(adb_plan* plan, const float* sample, uint8_t* out, int n, int c, int hw,
                   adb_stream stream);

/* ---- FID moments (evaluations/evaluator_v1.py:218-221): accumulates
 * sum_x[d] += sum_i f[i,d], sum_xx[d,e] += sum_i f[i,d] f[i,e] in fp64. */
int adb_moments_accumulate(adb_plan* plan, const float* feats, int n, int d, double* sum_x,
                           double* sum_xx, adb_stream stream);

/* zero `bytes` bytes at `ptr` (recorded memset node) */
int adb_memset0(adb_plan* plan, void* ptr, size_t bytes, adb_stream stream);

/* ==== classifier guidance: forward extras and the input-gradient (backward-data) path ====
 * The search's cond_fn (search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-390)
 * is th.autograd.grad(log_softmax(classifier(x_t, t))[range(B), y].sum(), x_t) * classifier_scale
 * with classifier = EncoderUNetModel (guided_diffusion/unet.py:685-896, pool="attention").
 * Its convolutions' data gradients are adb_conv_igemm calls with transposed / flipped weights;
 * the entries below are the remaining pieces. */

/* adb_attention that also writes lse[b*heads, t] (fp32, log2 domain: P_ij = 2^(s_ij*log2(e)/8 - lse_i)),
 * which adb_attention_backward recomputes the probabilities from. */
int adb_attention_lse(adb_plan* plan, const void* qkv, void* out, float* lse, int b, int t, int heads,
                      int legacy_order, adb_stream stream);

/* Gradient of adb_attention w.r.t. qkv (autograd of QKVAttention(Legacy).forward, unet.py:325-371).
 * out / dout: bf16 [b*t, heads*64] (forward output and its gradient); dqkv: bf16 [b*t, 3*heads*64], same
 * column layout as qkv; dsum: fp32 [b*heads, t] scratch (row sums of dout*out). */
int adb_attention_backward(adb_plan* plan, const void* qkv, const void* out, const void* dout, const float* lse,
                           float* dsum, void* dqkv, int b, int t, int heads, int legacy_order,
                           adb_stream stream);
/* Same, with an fp32 [b*t, heads*64] workspace `dq_ws`: when t is a multiple of 128 the single-pass kernel runs (one
 * exponential and five tile products per score instead of two and seven; dQ partials are reduced into dq_ws with
 * TMA reductions (cp.reduce.async.bulk.tensor) and rounded into dqkv by a second kernel). dq_ws == NULL, or the
 * deterministic mode (see adb_set_attention_backward_fused), behaves like adb_attention_backward. */
int adb_attention_backward_ws(adb_plan* plan, const void* qkv, const void* out, const void* dout, const float* lse,
                              float* dsum, void* dqkv, float* dq_ws, int b, int t, int heads, int legacy_order,
                              adb_stream stream);
/* Selects what adb_attention_backward_ws records from now on: 1 = the single-pass kernel (faster; dQ partials are summed
 * by L2 reductions in no fixed order, so results are reproducible to fp32 rounding only), 0 = the deterministic
 * two-kernel form (default; also ADB_ATTN_BWD_FUSED=1 in the environment). on < 0 only queries. Returns the mode in force. */
int adb_set_attention_backward_fused(int on);

/* Gradient of adb_groupnorm (single source) w.r.t. its input: GroupNorm32 (+FiLM) (+SiLU) (+2x average
 * pool) backward (nn.py:17-19, unet.py:236-258 under autograd). */
typedef struct {
  int n, h, w, c;           /* geometry of x, the forward op's INPUT */
  const void* x;            /* bf16 NHWC [n,h,w,c] */
  const double* stats;      /* the forward's [n,32,2] (sum, sum of squares) of x */
  const float* gamma;
  const float* beta;
  float eps;
  const float* scale_shift; /* as in adb_gn_desc, or NULL */
  int ss_stride;
  int silu;
  int resample;             /* ADB_RESAMPLE_NONE, or ADB_RESAMPLE_AVGPOOL2: dout is [n,h/2,w/2,c] */
  const void* dout;         /* bf16 NHWC gradient w.r.t. the forward op's output */
  const void* add;          /* optional bf16 gradient added to dx (the block's skip path) */
  int add_mode;             /* ADB_RES_NONE; ADB_RES_SAME: add is [n,h,w,c]; ADB_RES_AVGPOOL2: the skip path
                               average-pooled x, add is [n,h/2,w/2,c] and contributes add/4 */
  void* dx;                 /* bf16 NHWC [n,h,w,c] */
  double* bstats;           /* scratch [n,32,2]; with bstats_ready: the sums, already reduced by the producing conv */
  int bstats_ready;         /* 1: bstats was filled by adb_conv_igemm (gnb_*): only the apply pass runs */
} adb_gn_bwd_desc;
int adb_gn_backward(adb_plan* plan, const adb_gn_bwd_desc* d, adb_stream stream);

/* AttentionPool2d (unet.py:22-51) restricted to the token it returns (index 0 = the spatial mean):
 *   prepare : xp[n,p,:] = h[n,p,:] + pos[:,1+p] (bf16), mean[n,:] = mean_p h[n,p,:] + pos[:,0] (fp32);
 *             h bf16 [n,P,C] (NHWC pixels), pos fp32 [C, P+1]
 *   the k/v rows of qkv_proj over xp run on adb_conv_igemm, the mean token's q/k/v on adb_linear
 *   attention: out0[n,:] = softmax(q0 . k_j / 8) v_j over the P+1 tokens, per 64-channel head;
 *             qkv0 fp32 [n,3C] = (q|k|v) of the mean token, kv bf16 [n,P,2C] = (k|v) of the pixels;
 *             probs fp32 [n, C/64, P+1] is kept for the backward
 *   backward: dout0 fp32 [n,C] -> dqkv0 fp32 [n,3C], dkv bf16 [n,P,2C]
 *   merge   : dh[n,p,:] = dxp[n,p,:] + dmean[n,:] / P */
int adb_pool_prepare(adb_plan* plan, const void* h, const float* pos, void* xp, float* mean, int n, int p, int c,
                     adb_stream stream);
int adb_pool_attention(adb_plan* plan, const float* qkv0, const void* kv, float* out0, float* probs, int n, int p,
                       int c, adb_stream stream);
int adb_pool_attention_backward(adb_plan* plan, const float* dout0, const float* probs, const float* qkv0,
                                const void* kv, float* dqkv0, void* dkv, int n, int p, int c, adb_stream stream);
int adb_pool_merge(adb_plan* plan, const void* dxp, const float* dmean, void* dh, int n, int p, int c,
                   adb_stream stream);

/* dlogits[n,c] = scale * ((c == y[n]) - softmax(logits[n])[c]): the gradient of
 * log_softmax(logits)[range(n), y].sum() * scale (…progressive.py:387-390). */
int adb_logsoftmax_grad(adb_plan* plan, const float* logits, const int64_t* y, float* dlogits, int n, int k,
                        float scale, adb_stream stream);

/* ==== Stable-Diffusion-v1 family (BASELINE configs[4]); paths relative to
 * /root/reference/examples/"Stable Diffusion"/ ==== */

/* Fused softmax attention between the projections of CrossAttention.forward (ldm/modules/attention.py:170-194):
 * out = softmax(q k^T * d_head^-0.5) v per (batch, head), self- or cross-attention.
 * Every head occupies d_pad (multiple of 64, <= 192) columns of the q / k / v matrices and of `out`; columns
 * [d_head, d_pad) of q, k, v must be zero (zero-padded projection weights) and come out zero.
 * q   : bf16 [b*tq, q_width], head h at columns q_col0 + h*d_pad
 * kv  : bf16 [b*tk_rows, kv_width], K of head h at k_col0 + h*d_pad, V at v_col0 + h*d_pad; only the first
 *       tk_valid of the tk_rows rows of a batch element are keys (77 context tokens in a 128-row padded buffer)
 * out : bf16 [b*tq, heads*d_pad] */
typedef struct {
  const void* q;
  int q_width, q_col0;
  const void* kv;
  int kv_width, k_col0, v_col0;
  void* out;
  int b, heads, d_head, d_pad;
  int tq, tk_rows, tk_valid;
  int v_ones;  /* 1: column d_head of every V row holds 1.0 (put there by the V projection's bias; needs d_pad > d_head):
                  the kernel then takes the softmax denominator from the PV product instead of summing P itself */
} adb_attn_sd_desc;
int adb_attention_sd(adb_plan* plan, const adb_attn_sd_desc* d, adb_stream stream);

/* nn.LayerNorm(c) over the channels of every token (BasicTransformerBlock.norm1/2/3, attention.py:205-207):
 * x, out bf16 [rows, c]; gamma, beta fp32 [c]; c % 8 == 0, c <= 2048. Two-pass statistics in fp32 registers. */
int adb_layernorm(adb_plan* plan, const void* x, const float* gamma, const float* beta, void* out, int rows, int c,
                  float eps, adb_stream stream);

/* GEGLU gate (attention.py:37-44): out[r, j] = x[r, j] * gelu(x[r, inner + j]), exact (erf) GELU;
 * x bf16 [rows, 2*inner], out bf16 [rows, inner]; inner % 8 == 0. */
int adb_geglu(adb_plan* plan, const void* x, void* out, int rows, int inner, adb_stream stream);

/* Classifier-free-guidance combine + DDIM update of p_sample_ddim with eta = 0 (ldm/models/diffusion/ddim.py:
 * 184-216): e_t = e_u + scale * (e_c - e_u); pred_x0 = (x - coef[0] * e_t) / coef[1];
 * x_prev = coef[2] * pred_x0 + coef[3] * e_t, with coef = {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)}
 * in fp32 and every intermediate rounded as the reference's separate fp32 tensor ops round it (bit-exact given eps).
 * eps: fp32 [2n, chw], unconditional half first (torch.cat([uncond, cond]), :187-189); with scale == 1 or
 * cfg == 0 eps is [n, chw] and used as is. x_prev may alias x. pred_x0 may be NULL. */
int adb_cfg_ddim_step(adb_plan* plan, const float* x, const float* eps, float* x_prev, float* pred_x0, int n,
                      int chw, int cfg, float scale, const float coef[4], adb_stream stream);

/* PLMS (ldm/models/diffusion/plms.py:190-257). adb_cfg_combine: e_out = e_u + scale * (e_c - e_u) over `total`
 * elements (eps is [2*total] with the unconditional half first when cfg, else [total] copied), the tensor p_sample_plms
 * returns as e_t and keeps in `old_eps`. adb_plms_update: e' from e_t and up to three older eps tensors -
 * mode 0: e_t; 1: (e_t + o1) / 2 with o1 = e_t_next; 2: (3 e_t - o1) / 2; 3: (23 e_t - 16 o1 + 5 o2) / 12;
 * 4: (55 e_t - 59 o1 + 37 o2 - 9 o3) / 24 (o1 = newest) - followed by the eta = 0 update of adb_cfg_ddim_step with e'.
 * Bit-exact against the reference's fp32 tensor expressions. x_prev may alias x. */
int adb_cfg_combine(adb_plan* plan, const float* eps, float* e_out, size_t total, int cfg, float scale, adb_stream stream);
int adb_plms_update(adb_plan* plan, const float* x, const float* e_t, const float* o1, const float* o2, const float* o3,
                    int mode, const float coef[4], float* x_prev, float* pred_x0, size_t total, adb_stream stream);

/* DPM-Solver++(2M) (ldm/models/diffusion/dpm_solver/dpm_solver.py). adb_timestep_embedding_f32: the sinusoid of
 * fractional model timesteps (t_continuous - 1/N) * 1000 (:278-286). adb_dpm_x0: classifier-free combine (:336-343)
 * + data prediction x0 = (x - sigma_t * noise) / alpha_t (:386-391) over `total` elements (eps is [2*total],
 * unconditional half first, when cfg). adb_dpm_update: order 1 x_t = c0 x - c1 m0 (:519-533); order 2
 * x_t = c0 x - c1 m0 - c2 * (inv_r0 * (m0 - m1)) (:770-790); the scalars are the reference's fp32 schedule
 * expressions evaluated on the host. x_out may alias x. */
int adb_timestep_embedding_f32(adb_plan* plan, const float* t, const float* freqs, float* out, int b, int dim,
                               adb_stream stream);
int adb_dpm_x0(adb_plan* plan, const float* x, const float* eps, float* x0, size_t total, int cfg, float scale,
               float sigma, float alpha, adb_stream stream);
int adb_dpm_update(adb_plan* plan, const float* x, const float* m0, const float* m1, float* x_out, size_t total,
                   int order, float c0, float c1, float c2, float inv_r0, adb_stream stream);

/* fp32 [n, t, c] -> bf16 [n, t_pad, c_pad] zero-padded: the text context (77 x 768) into the 128-row buffer the
 * K/V projection GEMMs and adb_attention_sd read. */
int adb_pad_context(adb_plan* plan, const float* ctx, void* out, int n, int t, int c, int t_pad, adb_stream stream);

/* ---- Inception-V3 pool_3 feature extractor on the device ----
 * Replaces the TensorFlow graph evaluation of evaluations/evaluator_v1.py:252-280, 665-679 (pool_3 of
 * classify_image_graph_def.pb) and pytorch-fid's InceptionV3 of "Stable Diffusion"/scripts/search_ea.py:95-127, 171-182.
 * Every convolution = adb_gather_patches + adb_conv_igemm (1 tap over the patch rows, BatchNorm folded into the
 * packed weights and bias); a conv writes its pre-activation and consumers apply ReLU on load.
 * "Sources" are up to 4 bf16 NHWC tensors of equal n, h, w read as one channel concatenation (th.cat is never
 * materialised): ptrs[i], chans[i] (multiples of 8), relu[i] != 0 -> max(x, 0) on load. */

/* uint8 NHWC [n, h, w, 3] -> bf16 NHWC [n, oh, ow, 8] (channels 3..7 zero), bilinear with half-pixel centres
 * (F.interpolate(..., mode="bilinear", align_corners=False), as pytorch-fid resizes to 299 x 299), then x / 127.5 - 1. */
int adb_resize_bilinear_u8(adb_plan* plan, const uint8_t* in, void* out, int n, int h, int w, int oh, int ow,
                           adb_stream stream);
/* im2col: out bf16 [n * ho * wo, k_pad], row = (ky, kx) taps x concatenated channels, zero tail up to k_pad (a multiple
 * of 8), taps outside the image zero; ho = (h + 2 ph - kh) / stride + 1 (wo likewise). */
int adb_gather_patches(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out,
                       int n, int h, int w, int kh, int kw, int stride, int ph, int pw, int k_pad, adb_stream stream);
/* 3x3 pooling, stride 1 or 2, padding 0 or 1. mode 0: max; 1: average over 9 (count_include_pad, torchvision);
 * 2: average over the in-image taps (the FID Inception). out bf16 [n, ho, wo, sum chans]. */
int adb_pool3x3(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out, int n,
                int h, int w, int stride, int pad, int mode, adb_stream stream);
/* global average pool over hw pixels: out fp32 [n, sum chans] (pool_3). */
int adb_global_avgpool(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, float* out,
                       int n, int hw, adb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* ADB200_H */
