// host.cu — host-side runtime of libadb200: error strings, TMA descriptor encoding,
// plan record/replay and the extern "C" entry points declared in include/adb200.h.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace adb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return ADB_ERR_CUDA;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// The driver entry point is resolved through the runtime so the library needs no direct
// libcuda link (nvcc's static cudart dlopens the driver).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return ADB_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return ADB_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("TMA box dim %d = %u out of range", i, box[i]);
      return ADB_ERR_INVALID;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) {
      set_error("TMA stride %d = %llu bytes is not a multiple of 16", i,
                (unsigned long long)strides_bytes[i]);
      return ADB_ERR_INVALID;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return ADB_ERR_CUDA;
  }
  return ADB_OK;
}

// fp32 tensor (e.g. the dQ workspace the fused attention backward reduces into with cp.reduce.async.bulk.tensor)
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return ADB_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return ADB_ERR_INVALID;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) {
      set_error("TMA box dim %d = %u out of range", i, box[i]);
      return ADB_ERR_INVALID;
    }
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) {
      set_error("TMA stride %d = %llu bytes is not a multiple of 16", i,
                (unsigned long long)strides_bytes[i]);
      return ADB_ERR_INVALID;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return ADB_ERR_CUDA;
  }
  return ADB_OK;
}

// implemented in the kernel translation units
int conv_block_n(int cout);
int conv_gnb_supported(int n, int h, int w, int cout);
int conv_igemm_submit(adb_plan*, const adb_conv_desc*, cudaStream_t);
int attention_submit(adb_plan*, const void*, void*, float*, int, int, int, int, cudaStream_t);
int attention_backward_submit(adb_plan*, const void*, const void*, const void*, const float*, float*, void*, float*, int, int,
                              int, int, cudaStream_t);
int attention_backward_fused_mode(int);
int gn_backward_submit(adb_plan*, const adb_gn_bwd_desc*, cudaStream_t);
int pool_prepare_submit(adb_plan*, const void*, const float*, void*, float*, int, int, int, cudaStream_t);
int pool_attention_submit(adb_plan*, const float*, const void*, float*, float*, int, int, int, cudaStream_t);
int pool_attention_backward_submit(adb_plan*, const float*, const float*, const float*, const void*, float*, void*, int,
                                   int, int, cudaStream_t);
int pool_merge_submit(adb_plan*, const void*, const float*, void*, int, int, int, cudaStream_t);
int logsoftmax_grad_submit(adb_plan*, const float*, const int64_t*, float*, int, int, float, cudaStream_t);
int groupnorm_submit(adb_plan*, const adb_gn_desc*, cudaStream_t);
int resample2x_submit(adb_plan*, const void*, void*, int, int, int, int, int, cudaStream_t);
int stem_conv_submit(adb_plan*, const float*, const float*, const float*, void*, int, int, int, int,
                     int, cudaStream_t);
int timestep_embedding_submit(adb_plan*, const int64_t*, const float*, float*, int, int, cudaStream_t);
int linear_submit(adb_plan*, const float*, const float*, const float*, float*, int, int, int, int,
                  const float*, const int64_t*, cudaStream_t);
int split_bf16_submit(adb_plan*, const float*, void*, void*, size_t, int, cudaStream_t);
int stem_im2col_submit(adb_plan*, const float*, void*, int, int, int, cudaStream_t);
int ddim_step_submit(adb_plan*, const float*, const float*, int, const float*, float*, float*, int,
                     int, int, const float*, int, cudaStream_t);
int pack_uint8_submit(adb_plan*, const float*, uint8_t*, int, int, int, cudaStream_t);
int moments_submit(adb_plan*, const float*, int, int, double*, double*, cudaStream_t);
int attention_sd_submit(adb_plan*, const adb_attn_sd_desc*, cudaStream_t);
int layernorm_submit(adb_plan*, const void*, const float*, const float*, void*, int, int, float, cudaStream_t);
int geglu_submit(adb_plan*, const void*, void*, int, int, cudaStream_t);
int cfg_ddim_step_submit(adb_plan*, const float*, const float*, float*, float*, int, int, int, float, const float*,
                         cudaStream_t);
int pad_context_submit(adb_plan*, const float*, void*, int, int, int, int, cudaStream_t);
int cfg_combine_submit(adb_plan*, const float*, float*, size_t, int, float, cudaStream_t);
int timestep_embedding_f32_submit(adb_plan*, const float*, const float*, float*, int, int, cudaStream_t);
int dpm_x0_submit(adb_plan*, const float*, const float*, float*, size_t, int, float, float, float, cudaStream_t);
int dpm_update_submit(adb_plan*, const float*, const float*, const float*, float*, size_t, int, float, float, float, float,
                      cudaStream_t);
int gather_patches_submit(adb_plan*, const void* const*, const int*, const int*, int, void*, int, int, int, int, int, int, int,
                          int, int, cudaStream_t);
int pool3x3_submit(adb_plan*, const void* const*, const int*, const int*, int, void*, int, int, int, int, int, int, cudaStream_t);
int global_avgpool_submit(adb_plan*, const void* const*, const int*, const int*, int, float*, int, int, cudaStream_t);
int resize_bilinear_u8_submit(adb_plan*, const uint8_t*, void*, int, int, int, int, int, cudaStream_t);
int plms_update_submit(adb_plan*, const float*, const float*, const float*, const float*, const float*, int, const float*,
                       float*, float*, size_t, cudaStream_t);

}  // namespace adb

using namespace adb;

extern "C" {

const char* adb_last_error(void) { return g_err; }
int adb_version(void) { return 100; }

int adb_device_check(void) {
  int dev = 0;
  ADB_CUDA(cudaGetDevice(&dev));
  int major = 0;
  ADB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100 (B200)", major);
    return ADB_ERR_UNSUPPORTED;
  }
  return ADB_OK;
}

adb_plan* adb_plan_create(void) { return new (std::nothrow) adb_plan(); }
void adb_plan_destroy(adb_plan* plan) { delete plan; }
int adb_plan_num_ops(const adb_plan* plan) { return plan ? (int)plan->ops.size() : 0; }

int adb_plan_run(adb_plan* plan, adb_stream stream) {
  if (!plan) {
    set_error("adb_plan_run: null plan");
    return ADB_ERR_INVALID;
  }
  int launches = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (auto& op : plan->ops) {
    int r = op(s);
    if (r < 0) return r;
    launches += r;
  }
  return launches;
}

int adb_plan_op_info(const adb_plan* plan, int i, const char** kind, double* flops, double* bytes) {
  if (!plan || i < 0 || i >= (int)plan->info.size()) {
    set_error("adb_plan_op_info: index out of range");
    return ADB_ERR_INVALID;
  }
  if (kind) *kind = plan->info[i].kind;
  if (flops) *flops = plan->info[i].flops;
  if (bytes) *bytes = plan->info[i].bytes;
  return ADB_OK;
}

// Runs the plan with a CUDA event pair around every op (on `stream`, the stream the kernels are
// launched on) and returns each op's device time in milliseconds. Synchronises the stream.
int adb_plan_run_profiled(adb_plan* plan, adb_stream stream, float* ms_out, int capacity) {
  if (!plan || !ms_out || capacity < (int)plan->ops.size()) {
    set_error("adb_plan_run_profiled: bad arguments");
    return ADB_ERR_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = plan->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) ADB_CUDA(cudaEventCreate(&e));
  int rc = ADB_OK;
  ADB_CUDA(cudaEventRecord(ev[0], s));
  for (size_t i = 0; i < n; ++i) {
    int r = plan->ops[i](s);
    if (r < 0) {
      rc = r;
      break;
    }
    ADB_CUDA(cudaEventRecord(ev[i + 1], s));
  }
  if (rc == ADB_OK) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc == ADB_OK) {
    for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc == ADB_OK ? (int)n : rc;
}

int adb_conv_block_n(int cout) { return conv_block_n(cout); }
int adb_conv_gnb_supported(int n, int h, int w, int cout) { return conv_gnb_supported(n, h, w, cout); }

int adb_conv_igemm(adb_plan* plan, const adb_conv_desc* d, adb_stream stream) {
  return conv_igemm_submit(plan, d, static_cast<cudaStream_t>(stream));
}

int adb_attention(adb_plan* plan, const void* qkv, void* out, int b, int t, int heads,
                  int legacy_order, adb_stream stream) {
  return attention_submit(plan, qkv, out, nullptr, b, t, heads, legacy_order, static_cast<cudaStream_t>(stream));
}

int adb_attention_lse(adb_plan* plan, const void* qkv, void* out, float* lse, int b, int t, int heads,
                      int legacy_order, adb_stream stream) {
  if (!lse) {
    set_error("adb_attention_lse: null lse");
    return ADB_ERR_INVALID;
  }
  return attention_submit(plan, qkv, out, lse, b, t, heads, legacy_order, static_cast<cudaStream_t>(stream));
}

int adb_attention_backward(adb_plan* plan, const void* qkv, const void* out, const void* dout, const float* lse,
                           float* dsum, void* dqkv, int b, int t, int heads, int legacy_order,
                           adb_stream stream) {
  return attention_backward_submit(plan, qkv, out, dout, lse, dsum, dqkv, nullptr, b, t, heads, legacy_order,
                                   static_cast<cudaStream_t>(stream));
}

int adb_set_attention_backward_fused(int on) { return attention_backward_fused_mode(on); }

int adb_attention_backward_ws(adb_plan* plan, const void* qkv, const void* out, const void* dout, const float* lse,
                              float* dsum, void* dqkv, float* dq_ws, int b, int t, int heads, int legacy_order,
                              adb_stream stream) {
  return attention_backward_submit(plan, qkv, out, dout, lse, dsum, dqkv, dq_ws, b, t, heads, legacy_order,
                                   static_cast<cudaStream_t>(stream));
}

int adb_gn_backward(adb_plan* plan, const adb_gn_bwd_desc* d, adb_stream stream) {
  return gn_backward_submit(plan, d, static_cast<cudaStream_t>(stream));
}

int adb_pool_prepare(adb_plan* plan, const void* h, const float* pos, void* xp, float* mean, int n, int p, int c,
                     adb_stream stream) {
  return pool_prepare_submit(plan, h, pos, xp, mean, n, p, c, static_cast<cudaStream_t>(stream));
}

int adb_pool_attention(adb_plan* plan, const float* qkv0, const void* kv, float* out0, float* probs, int n, int p,
                       int c, adb_stream stream) {
  return pool_attention_submit(plan, qkv0, kv, out0, probs, n, p, c, static_cast<cudaStream_t>(stream));
}

int adb_pool_attention_backward(adb_plan* plan, const float* dout0, const float* probs, const float* qkv0,
                                const void* kv, float* dqkv0, void* dkv, int n, int p, int c, adb_stream stream) {
  return pool_attention_backward_submit(plan, dout0, probs, qkv0, kv, dqkv0, dkv, n, p, c,
                                        static_cast<cudaStream_t>(stream));
}

int adb_pool_merge(adb_plan* plan, const void* dxp, const float* dmean, void* dh, int n, int p, int c,
                   adb_stream stream) {
  return pool_merge_submit(plan, dxp, dmean, dh, n, p, c, static_cast<cudaStream_t>(stream));
}

int adb_logsoftmax_grad(adb_plan* plan, const float* logits, const int64_t* y, float* dlogits, int n, int k,
                        float scale, adb_stream stream) {
  return logsoftmax_grad_submit(plan, logits, y, dlogits, n, k, scale, static_cast<cudaStream_t>(stream));
}

int adb_groupnorm(adb_plan* plan, const adb_gn_desc* d, adb_stream stream) {
  return groupnorm_submit(plan, d, static_cast<cudaStream_t>(stream));
}

int adb_resample2x(adb_plan* plan, const void* src, void* dst, int n, int h, int w, int c, int mode,
                   adb_stream stream) {
  return resample2x_submit(plan, src, dst, n, h, w, c, mode, static_cast<cudaStream_t>(stream));
}

int adb_stem_conv(adb_plan* plan, const float* x, const float* weight, const float* bias, void* out,
                  int n, int cin, int h, int w, int cout, adb_stream stream) {
  return stem_conv_submit(plan, x, weight, bias, out, n, cin, h, w, cout,
                          static_cast<cudaStream_t>(stream));
}

int adb_timestep_embedding(adb_plan* plan, const int64_t* t, const float* freqs, float* out, int b,
                           int dim, adb_stream stream) {
  return timestep_embedding_submit(plan, t, freqs, out, b, dim, static_cast<cudaStream_t>(stream));
}

int adb_linear(adb_plan* plan, const float* x, const float* w, const float* bias, float* out, int b,
               int k, int nout, int silu_in, const float* table, const int64_t* idx,
               adb_stream stream) {
  return linear_submit(plan, x, w, bias, out, b, k, nout, silu_in, table, idx,
                       static_cast<cudaStream_t>(stream));
}

int adb_stem_im2col(adb_plan* plan, const float* x, void* out, int n, int h, int w, adb_stream stream) {
  return stem_im2col_submit(plan, x, out, n, h, w, static_cast<cudaStream_t>(stream));
}

int adb_split_bf16(adb_plan* plan, const float* x, void* hi, void* lo, size_t total, int silu_in, adb_stream stream) {
  return split_bf16_submit(plan, x, hi, lo, total, silu_in, static_cast<cudaStream_t>(stream));
}

int adb_ddim_step(adb_plan* plan, const float* x, const float* model_out, int eps_channels,
                  const float* grad, float* x_prev, float* pred_xstart, int n, int c, int hw,
                  const float coef[5], int clip_denoised, adb_stream stream) {
  return ddim_step_submit(plan, x, model_out, eps_channels, grad, x_prev, pred_xstart, n, c, hw, coef,
                          clip_denoised, static_cast<cudaStream_t>(stream));
}

int adb_pack_uint8(adb_plan* plan, const float* sample, uint8_t* out, int n, int c, int hw,
                   adb_stream stream) {
  return pack_uint8_submit(plan, sample, out, n, c, hw, static_cast<cudaStream_t>(stream));
}

int adb_moments_accumulate(adb_plan* plan, const float* feats, int n, int d, double* sum_x,
                           double* sum_xx, adb_stream stream) {
  return moments_submit(plan, feats, n, d, sum_x, sum_xx, static_cast<cudaStream_t>(stream));
}

int adb_attention_sd(adb_plan* plan, const adb_attn_sd_desc* d, adb_stream stream) {
  return attention_sd_submit(plan, d, static_cast<cudaStream_t>(stream));
}

int adb_layernorm(adb_plan* plan, const void* x, const float* gamma, const float* beta, void* out, int rows, int c,
                  float eps, adb_stream stream) {
  return layernorm_submit(plan, x, gamma, beta, out, rows, c, eps, static_cast<cudaStream_t>(stream));
}

int adb_geglu(adb_plan* plan, const void* x, void* out, int rows, int inner, adb_stream stream) {
  return geglu_submit(plan, x, out, rows, inner, static_cast<cudaStream_t>(stream));
}

int adb_cfg_ddim_step(adb_plan* plan, const float* x, const float* eps, float* x_prev, float* pred_x0, int n,
                      int chw, int cfg, float scale, const float coef[4], adb_stream stream) {
  return cfg_ddim_step_submit(plan, x, eps, x_prev, pred_x0, n, chw, cfg, scale, coef, static_cast<cudaStream_t>(stream));
}

int adb_cfg_combine(adb_plan* plan, const float* eps, float* e_out, size_t total, int cfg, float scale, adb_stream stream) {
  return cfg_combine_submit(plan, eps, e_out, total, cfg, scale, static_cast<cudaStream_t>(stream));
}

int adb_plms_update(adb_plan* plan, const float* x, const float* e_t, const float* o1, const float* o2, const float* o3,
                    int mode, const float coef[4], float* x_prev, float* pred_x0, size_t total, adb_stream stream) {
  return plms_update_submit(plan, x, e_t, o1, o2, o3, mode, coef, x_prev, pred_x0, total, static_cast<cudaStream_t>(stream));
}

int adb_timestep_embedding_f32(adb_plan* plan, const float* t, const float* freqs, float* out, int b, int dim,
                               adb_stream stream) {
  return timestep_embedding_f32_submit(plan, t, freqs, out, b, dim, static_cast<cudaStream_t>(stream));
}

int adb_dpm_x0(adb_plan* plan, const float* x, const float* eps, float* x0, size_t total, int cfg, float scale,
               float sigma, float alpha, adb_stream stream) {
  return dpm_x0_submit(plan, x, eps, x0, total, cfg, scale, sigma, alpha, static_cast<cudaStream_t>(stream));
}

int adb_dpm_update(adb_plan* plan, const float* x, const float* m0, const float* m1, float* x_out, size_t total,
                   int order, float c0, float c1, float c2, float inv_r0, adb_stream stream) {
  return dpm_update_submit(plan, x, m0, m1, x_out, total, order, c0, c1, c2, inv_r0, static_cast<cudaStream_t>(stream));
}

int adb_pad_context(adb_plan* plan, const float* ctx, void* out, int n, int t, int c, int t_pad, adb_stream stream) {
  return pad_context_submit(plan, ctx, out, n, t, c, t_pad, static_cast<cudaStream_t>(stream));
}

int adb_resize_bilinear_u8(adb_plan* plan, const uint8_t* in, void* out, int n, int h, int w, int oh, int ow,
                           adb_stream stream) {
  return resize_bilinear_u8_submit(plan, in, out, n, h, w, oh, ow, static_cast<cudaStream_t>(stream));
}

int adb_gather_patches(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out,
                       int n, int h, int w, int kh, int kw, int stride, int ph, int pw, int k_pad, adb_stream stream) {
  return gather_patches_submit(plan, ptrs, chans, relu, nsrc, out, n, h, w, kh, kw, stride, ph, pw, k_pad,
                               static_cast<cudaStream_t>(stream));
}

int adb_pool3x3(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, void* out, int n,
                int h, int w, int stride, int pad, int mode, adb_stream stream) {
  return pool3x3_submit(plan, ptrs, chans, relu, nsrc, out, n, h, w, stride, pad, mode, static_cast<cudaStream_t>(stream));
}

int adb_global_avgpool(adb_plan* plan, const void* const* ptrs, const int* chans, const int* relu, int nsrc, float* out,
                       int n, int hw, adb_stream stream) {
  return global_avgpool_submit(plan, ptrs, chans, relu, nsrc, out, n, hw, static_cast<cudaStream_t>(stream));
}

int adb_memset0(adb_plan* plan, void* ptr, size_t bytes, adb_stream stream) {
  if (!ptr && bytes) {
    set_error("adb_memset0: null pointer");
    return ADB_ERR_INVALID;
  }
  return submit(plan, static_cast<cudaStream_t>(stream), "memset", 0.0, 0.0, [ptr, bytes](cudaStream_t s) -> int {
    ADB_CUDA(cudaMemsetAsync(ptr, 0, bytes, s));
    return 1;
  });
}

}  // extern "C"
