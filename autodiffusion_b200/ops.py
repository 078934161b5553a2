"""Torch-tensor front end of the C-ABI ops (include/adb200.h).

PyTorch is used here only for device memory and the current CUDA stream: every function
hands raw device pointers to libadb200.so. Tensors must live on a CUDA device; a CPU tensor
raises — there is no fallback.

Activation layout: bf16 NHWC `[n, h, w, c]` contiguous. Passing `plan=` records the op
into a `Plan` (replayed with `Plan.run()`, capturable into a CUDA graph) instead of
launching it.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (
    OUT_BF16_NHWC,
    OUT_F32_NCHW,
    RES_AVGPOOL2,
    RES_NEAREST2,
    RES_NONE,
    RES_SAME,
    RESAMPLE_AVGPOOL2,
    RESAMPLE_NEAREST2,
    RESAMPLE_NONE,
)

__all__ = [
    "Plan", "conv_block_n", "pack_conv_weight", "conv_igemm", "attention", "groupnorm", "resample2x",
    "stem_conv", "timestep_embedding", "linear", "ddim_step", "pack_uint8", "moments_accumulate",
    "memset0", "nchw_to_nhwc_bf16", "nhwc_to_nchw_f32",
    "attention_backward", "gn_backward", "pool_prepare", "pool_attention", "pool_attention_backward", "pool_merge",
    "logsoftmax_grad", "pack_conv_weight_dgrad", "pack_linear_weight_split", "linear_tc",
    "pack_stem_weight", "stem_conv_tc",
    "attention_sd", "layernorm", "geglu", "cfg_ddim_step", "pad_context", "cfg_combine", "plms_update", "dpm_x0", "dpm_update",
    "resize_bilinear_u8", "gather_patches", "pool3x3", "global_avgpool",
]


def _dev(t: torch.Tensor, name: str, dtype=None) -> int:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (autodiffusion_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t.data_ptr()


def _opt(t: Optional[torch.Tensor], name: str, dtype=None) -> Optional[int]:
    return None if t is None else _dev(t, name, dtype)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Plan:
    """A recorded launch schedule (adb_plan). Holds references to every tensor recorded into it."""

    def __init__(self):
        self._lib = _lib.lib()
        self._h = self._lib.adb_plan_create()
        if not self._h:
            raise MemoryError("adb_plan_create failed")
        self._keep = []
        self.launches_per_run = None

    def keep(self, *tensors):
        self._keep.extend(t for t in tensors if t is not None)

    @property
    def handle(self):
        return self._h

    def num_ops(self) -> int:
        return self._lib.adb_plan_num_ops(self._h)

    def run(self) -> int:
        n = _lib.check(self._lib.adb_plan_run(self._h, _stream()), "adb_plan_run")
        self.launches_per_run = n
        return n

    def op_info(self):
        """[(kind, flops, bytes)] per recorded op."""
        out = []
        kind, fl, by = C.c_char_p(), C.c_double(), C.c_double()
        for i in range(self.num_ops()):
            _lib.check(self._lib.adb_plan_op_info(self._h, i, C.byref(kind), C.byref(fl), C.byref(by)), "adb_plan_op_info")
            out.append((kind.value.decode(), fl.value, by.value))
        return out

    def run_profiled(self):
        """Run with CUDA events around every op; returns per-op milliseconds (syncs the stream)."""
        n = self.num_ops()
        ms = (C.c_float * n)()
        _lib.check(self._lib.adb_plan_run_profiled(self._h, _stream(), ms, n), "adb_plan_run_profiled")
        return list(ms)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.adb_plan_destroy(h)
            except Exception:
                pass


def _ph(plan: Optional[Plan]):
    return None if plan is None else plan.handle


def conv_block_n(cout: int) -> int:
    return _lib.lib().adb_conv_block_n(int(cout))


def pack_conv_weight(weights: Sequence[torch.Tensor], device=None) -> torch.Tensor:
    """Pack PyTorch conv weights into the implicit-GEMM K-major bf16 matrix.

    Each `[cout, cin, kh, kw]` (or `[cout, cin, 1]`) weight becomes `[cout, kh*kw*cin]` with K
    ordered (tap, cin); several weights (K-segments) are concatenated along K. Rows are
    zero-padded up to a multiple of the kernel's N tile.
    """
    mats = []
    cout = weights[0].shape[0]
    for w in weights:
        if w.dim() == 3:
            w = w.unsqueeze(-1)
        assert w.shape[0] == cout
        mats.append(w.detach().float().permute(0, 2, 3, 1).reshape(cout, -1))
    m = torch.cat(mats, dim=1)
    bn = conv_block_n(cout)
    pad = (-cout) % bn
    if pad:
        m = torch.cat([m, m.new_zeros(pad, m.shape[1])], dim=0)
    m = m.to(torch.bfloat16).contiguous()
    return m.to(device) if device is not None else m


def conv_igemm(
    segs: Sequence[tuple],
    weight: torch.Tensor,
    bias: Optional[torch.Tensor],
    cout: int,
    out: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    res_mode: int = RES_NONE,
    out_mode: int = OUT_BF16_NHWC,
    plan: Optional[Plan] = None,
    stats_out: Optional[torch.Tensor] = None,
    stats2: Optional[tuple] = None,
    gnb: Optional[dict] = None,
) -> torch.Tensor:
    """segs: [(act[n,h,w,cin] bf16, taps), ...]; weight from `pack_conv_weight`.
    stats_out: optional zeroed fp64 [n,32,2]; the epilogue adds the output's GroupNorm sums to it.
    stats2: optional (fp64 [n,32,2], cpg, channel offset) - sums for a consumer GroupNorm over a concat.
    gnb: this conv is a data gradient consumed by `gn_backward` of the GroupNorm described by the dict (keys x, stats,
      gamma, beta, bstats and optionally scale_shift, ss_stride, eps, silu - the arguments of `gn_backward`): the epilogue
      reduces that backward's two per-(image, group) sums into `bstats`; call `gn_backward(..., bstats_ready=True)`."""
    act0 = segs[0][0]
    s0 = segs[0][2] if len(segs[0]) > 2 else 1  # a third tuple entry is the segment's stride (1 or 2)
    n, h, w = act0.shape[0], act0.shape[1] // s0, act0.shape[2] // s0
    d = _lib.ConvDesc()
    d.n, d.h, d.w = n, h, w
    d.cout = cout
    d.cout_pad = weight.shape[0]
    d.nseg = len(segs)
    ktot = 0
    for i, sg in enumerate(segs):
        act, taps = sg[0], sg[1]
        stride = sg[2] if len(sg) > 2 else 1
        assert act.shape[:3] == (n, h * stride, w * stride), "all K-segments share the output geometry"
        d.seg[i].act = _dev(act, f"seg{i}.act", torch.bfloat16)
        d.seg[i].cin = act.shape[3]
        d.seg[i].taps = taps
        d.seg_stride[i] = stride
        ktot += taps * act.shape[3]
    if weight.shape[1] != ktot:
        raise ValueError(f"packed weight K={weight.shape[1]} does not match segments K={ktot}")
    d.weight = _dev(weight, "weight", torch.bfloat16)
    if bias is not None and bias.dim() == 2:  # per-image bias rows [n, cout]; may be a column slice (row stride > cout)
        if not bias.is_cuda or bias.dtype != torch.float32 or bias.stride(1) != 1 or bias.shape[0] != n:
            raise ValueError("per-image bias must be a CUDA fp32 [n, cout] tensor or row-strided view of one")
        d.bias = bias.data_ptr()
        d.bias_stride = bias.stride(0)
    else:
        d.bias = _opt(bias, "bias", torch.float32)
    d.residual = _opt(residual, "residual", torch.bfloat16)
    d.res_mode = res_mode
    if out is None:
        if out_mode == OUT_BF16_NHWC:
            out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=act0.device)
        else:
            out = torch.empty((n, cout, h, w), dtype=torch.float32, device=act0.device)
    d.out = _dev(out, "out")
    d.out_mode = out_mode
    d.stats_out = _opt(stats_out, "stats_out", torch.float64)
    if stats2 is not None:
        d.stats2_out = _dev(stats2[0], "stats2", torch.float64)
        d.stats2_cpg, d.stats2_choff = int(stats2[1]), int(stats2[2])
    gnb_keep = ()
    if gnb is not None:
        gx = gnb["x"]
        assert tuple(gx.shape) == (n, h, w, cout), f"gnb x shape {tuple(gx.shape)} != {(n, h, w, cout)}"
        d.gnb_x = _dev(gx, "gnb.x", torch.bfloat16)
        d.gnb_stats = _dev(gnb["stats"], "gnb.stats", torch.float64)
        d.gnb_gamma = _dev(gnb["gamma"], "gnb.gamma", torch.float32)
        d.gnb_beta = _dev(gnb["beta"], "gnb.beta", torch.float32)
        ss = gnb.get("scale_shift")
        if isinstance(ss, tuple):
            d.gnb_scale_shift = _dev(ss[0], "gnb.scale_shift", torch.float32) + 4 * int(ss[1])
            ss = ss[0]
        else:
            d.gnb_scale_shift = _opt(ss, "gnb.scale_shift", torch.float32)
        d.gnb_ss_stride = int(gnb.get("ss_stride", 0))
        d.gnb_eps = float(gnb.get("eps", 1e-5))
        d.gnb_silu = int(bool(gnb.get("silu", True)))
        d.gnb_bstats = _dev(gnb["bstats"], "gnb.bstats", torch.float64)
        gnb_keep = (gx, gnb["stats"], gnb["gamma"], gnb["beta"], ss, gnb["bstats"])
    _lib.check(_lib.lib().adb_conv_igemm(_ph(plan), C.byref(d), _stream()), "adb_conv_igemm")
    if plan is not None:
        plan.keep(*[s[0] for s in segs], weight, bias, residual, out, stats_out, stats2[0] if stats2 else None, *gnb_keep)
    return out


def conv_gnb_supported(n: int, h: int, w: int, cout: int) -> bool:
    """Whether `conv_igemm(..., gnb=...)` accepts an [n,h,w,cout] output (full tiles on the all-TMA epilogue)."""
    return bool(_lib.lib().adb_conv_gnb_supported(int(n), int(h), int(w), int(cout)))


def attention(qkv: torch.Tensor, b: int, t: int, heads: int, legacy_order: bool,
              out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None,
              lse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv: bf16 [b*t, 3*heads*64] -> bf16 [b*t, heads*64]. lse: optional fp32 [b*heads, t] output
    (log2-domain log-sum-exp of the scaled scores) consumed by `attention_backward`."""
    c = heads * 64
    assert qkv.numel() == b * t * 3 * c
    if out is None:
        out = torch.empty((b * t, c), dtype=torch.bfloat16, device=qkv.device)
    if lse is None:
        _lib.check(
            _lib.lib().adb_attention(_ph(plan), _dev(qkv, "qkv", torch.bfloat16), _dev(out, "out", torch.bfloat16),
                                     b, t, heads, int(bool(legacy_order)), _stream()),
            "adb_attention",
        )
    else:
        assert lse.numel() == b * heads * t
        _lib.check(
            _lib.lib().adb_attention_lse(_ph(plan), _dev(qkv, "qkv", torch.bfloat16), _dev(out, "out", torch.bfloat16),
                                         _dev(lse, "lse", torch.float32), b, t, heads, int(bool(legacy_order)), _stream()),
            "adb_attention_lse",
        )
    if plan is not None:
        plan.keep(qkv, out, lse)
    return out


def attention_backward(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, b: int, t: int,
                       heads: int, legacy_order: bool, dqkv: Optional[torch.Tensor] = None,
                       dsum: Optional[torch.Tensor] = None, plan: Optional[Plan] = None,
                       dq_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of `attention` w.r.t. qkv. out/dout bf16 [b*t, heads*64] -> dqkv bf16 [b*t, 3*heads*64].
    dq_ws: fp32 [b*t, heads*64] workspace of the single-pass kernel (allocated here when not given and t % 128 == 0)."""
    c = heads * 64
    assert qkv.numel() == b * t * 3 * c and out.numel() == b * t * c and dout.numel() == b * t * c
    if dqkv is None:
        dqkv = torch.empty((b * t, 3 * c), dtype=torch.bfloat16, device=qkv.device)
    if dsum is None:
        dsum = torch.empty((b * heads, t), dtype=torch.float32, device=qkv.device)
    if dq_ws is None and t % 128 == 0:
        dq_ws = torch.empty((b * t, c), dtype=torch.float32, device=qkv.device)
    if dq_ws is not None:
        assert dq_ws.numel() == b * t * c
    _lib.check(
        _lib.lib().adb_attention_backward_ws(_ph(plan), _dev(qkv, "qkv", torch.bfloat16), _dev(out, "out", torch.bfloat16),
                                             _dev(dout, "dout", torch.bfloat16), _dev(lse, "lse", torch.float32),
                                             _dev(dsum, "dsum", torch.float32), _dev(dqkv, "dqkv", torch.bfloat16),
                                             _opt(dq_ws, "dq_ws", torch.float32), b, t, heads, int(bool(legacy_order)),
                                             _stream()),
        "adb_attention_backward_ws",
    )
    if plan is not None:
        plan.keep(qkv, out, dout, lse, dsum, dqkv, dq_ws)
    return dqkv


def set_attention_backward_fused(on: Optional[bool]) -> bool:
    """Choose the attention-backward form recorded from now on (include/adb200.h: adb_set_attention_backward_fused):
    True = single-pass kernel (faster, fp32-rounding-level run-to-run differences), False = deterministic two-kernel form
    (default). None only queries. Returns the mode in force."""
    return bool(_lib.lib().adb_set_attention_backward_fused(-1 if on is None else int(bool(on))))


def gn_backward(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dout: torch.Tensor,
                scale_shift=None, ss_stride: int = 0, silu: bool = True, resample: int = RESAMPLE_NONE,
                add: Optional[torch.Tensor] = None, add_mode: int = RES_NONE, eps: float = 1e-5,
                dx: Optional[torch.Tensor] = None, bstats: Optional[torch.Tensor] = None,
                plan: Optional[Plan] = None, bstats_ready: bool = False) -> torch.Tensor:
    """Gradient of `groupnorm` (single source) w.r.t. its input x [n,h,w,c]; `stats` are the forward sums of x.
    bstats_ready: `bstats` already holds the two backward sums (`conv_igemm(..., gnb=...)` produced `dout`), so x and
    dout are read once instead of twice."""
    n, h, w, c = x.shape
    if dx is None:
        dx = torch.empty_like(x)
    if bstats is None:
        bstats = torch.empty((n, 32, 2), dtype=torch.float64, device=x.device)
    d = _lib.GnBwdDesc()
    d.n, d.h, d.w, d.c = n, h, w, c
    d.x = _dev(x, "x", torch.bfloat16)
    d.stats = _dev(stats, "stats", torch.float64)
    d.gamma = _dev(gamma, "gamma", torch.float32)
    d.beta = _dev(beta, "beta", torch.float32)
    d.eps = eps
    ss_base = None
    if isinstance(scale_shift, tuple):
        ss_base, ss_off = scale_shift
        d.scale_shift = _dev(ss_base, "scale_shift", torch.float32) + 4 * int(ss_off)
    else:
        d.scale_shift = _opt(scale_shift, "scale_shift", torch.float32)
        ss_base = scale_shift
    d.ss_stride = ss_stride
    d.silu = int(bool(silu))
    d.resample = resample
    exp = (n, h // 2, w // 2, c) if resample == RESAMPLE_AVGPOOL2 else (n, h, w, c)
    assert tuple(dout.shape) == exp, f"dout shape {tuple(dout.shape)} != {exp}"
    d.dout = _dev(dout, "dout", torch.bfloat16)
    d.add = _opt(add, "add", torch.bfloat16)
    d.add_mode = add_mode
    if add is not None:
        exp = (n, h // 2, w // 2, c) if add_mode == RES_AVGPOOL2 else (n, h, w, c)
        assert tuple(add.shape) == exp, f"add shape {tuple(add.shape)} != {exp}"
    d.dx = _dev(dx, "dx", torch.bfloat16)
    d.bstats = _dev(bstats, "bstats", torch.float64)
    d.bstats_ready = int(bool(bstats_ready))
    if bstats_ready and resample != RESAMPLE_NONE:
        raise ValueError("gn_backward: bstats_ready needs dout at x's resolution")
    _lib.check(_lib.lib().adb_gn_backward(_ph(plan), C.byref(d), _stream()), "adb_gn_backward")
    if plan is not None:
        plan.keep(x, stats, gamma, beta, ss_base, dout, add, dx, bstats)
    return dx


def pack_conv_weight_dgrad(weight: torch.Tensor, device=None) -> torch.Tensor:
    """Operand of the data-gradient of a conv: dx = conv(dy, W^T flipped). weight [cout, cin, kh, kw] (or
    [cout, cin, 1]) -> packed matrix of the transposed conv, rows = cin, K = (tap, cout)."""
    w = weight.detach()
    if w.dim() == 3:
        w = w.unsqueeze(-1)
    return pack_conv_weight([w.transpose(0, 1).flip(2, 3).contiguous()], device)


def pool_prepare(h: torch.Tensor, pos: torch.Tensor, plan: Optional[Plan] = None):
    """h bf16 [n, hh, ww, c], pos fp32 [c, hh*ww+1] -> (xp bf16 like h, mean fp32 [n, c])."""
    n, hh, ww, c = h.shape
    P = hh * ww
    assert tuple(pos.shape) == (c, P + 1)
    xp = torch.empty_like(h)
    mean = torch.empty((n, c), dtype=torch.float32, device=h.device)
    _lib.check(_lib.lib().adb_pool_prepare(_ph(plan), _dev(h, "h", torch.bfloat16), _dev(pos, "pos", torch.float32),
                                           _dev(xp, "xp", torch.bfloat16), _dev(mean, "mean", torch.float32), n, P, c,
                                           _stream()), "adb_pool_prepare")
    if plan is not None:
        plan.keep(h, pos, xp, mean)
    return xp, mean


def pool_attention(qkv0: torch.Tensor, kv: torch.Tensor, plan: Optional[Plan] = None):
    """qkv0 fp32 [n, 3c], kv bf16 [n, hh, ww, 2c] -> (out0 fp32 [n, c], probs fp32 [n, c/64, P+1])."""
    n, hh, ww, c2 = kv.shape
    c, P = c2 // 2, hh * ww
    out0 = torch.empty((n, c), dtype=torch.float32, device=kv.device)
    probs = torch.empty((n, c // 64, P + 1), dtype=torch.float32, device=kv.device)
    _lib.check(_lib.lib().adb_pool_attention(_ph(plan), _dev(qkv0, "qkv0", torch.float32), _dev(kv, "kv", torch.bfloat16),
                                             _dev(out0, "out0", torch.float32), _dev(probs, "probs", torch.float32),
                                             n, P, c, _stream()), "adb_pool_attention")
    if plan is not None:
        plan.keep(qkv0, kv, out0, probs)
    return out0, probs


def pool_attention_backward(dout0: torch.Tensor, probs: torch.Tensor, qkv0: torch.Tensor, kv: torch.Tensor,
                            plan: Optional[Plan] = None):
    """-> (dqkv0 fp32 [n, 3c], dkv bf16 like kv)."""
    n, hh, ww, c2 = kv.shape
    c, P = c2 // 2, hh * ww
    dqkv0 = torch.empty((n, 3 * c), dtype=torch.float32, device=kv.device)
    dkv = torch.empty_like(kv)
    _lib.check(_lib.lib().adb_pool_attention_backward(
        _ph(plan), _dev(dout0, "dout0", torch.float32), _dev(probs, "probs", torch.float32),
        _dev(qkv0, "qkv0", torch.float32), _dev(kv, "kv", torch.bfloat16), _dev(dqkv0, "dqkv0", torch.float32),
        _dev(dkv, "dkv", torch.bfloat16), n, P, c, _stream()), "adb_pool_attention_backward")
    if plan is not None:
        plan.keep(dout0, probs, qkv0, kv, dqkv0, dkv)
    return dqkv0, dkv


def pool_merge(dxp: torch.Tensor, dmean: torch.Tensor, plan: Optional[Plan] = None) -> torch.Tensor:
    """dh[n,p,:] = dxp[n,p,:] + dmean[n,:] / P."""
    n, hh, ww, c = dxp.shape
    dh = torch.empty_like(dxp)
    _lib.check(_lib.lib().adb_pool_merge(_ph(plan), _dev(dxp, "dxp", torch.bfloat16), _dev(dmean, "dmean", torch.float32),
                                         _dev(dh, "dh", torch.bfloat16), n, hh * ww, c, _stream()), "adb_pool_merge")
    if plan is not None:
        plan.keep(dxp, dmean, dh)
    return dh


def logsoftmax_grad(logits: torch.Tensor, y: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None,
                    plan: Optional[Plan] = None) -> torch.Tensor:
    """d/dlogits of log_softmax(logits)[range(n), y].sum() * scale."""
    n, k = logits.shape
    if out is None:
        out = torch.empty_like(logits)
    _lib.check(_lib.lib().adb_logsoftmax_grad(_ph(plan), _dev(logits, "logits", torch.float32), _dev(y, "y", torch.int64),
                                              _dev(out, "out", torch.float32), n, k, float(scale), _stream()),
               "adb_logsoftmax_grad")
    if plan is not None:
        plan.keep(logits, y, out)
    return out


def groupnorm(
    src0: torch.Tensor,
    gamma: torch.Tensor,
    beta: torch.Tensor,
    src1: Optional[torch.Tensor] = None,
    scale_shift: Optional[torch.Tensor] = None,
    ss_stride: int = 0,
    silu: bool = True,
    resample: int = RESAMPLE_NONE,
    eps: float = 1e-5,
    out: Optional[torch.Tensor] = None,
    stats: Optional[torch.Tensor] = None,
    plan: Optional[Plan] = None,
    stats_ready: bool = False,
) -> torch.Tensor:
    """stats_ready: `stats` already holds the producer-accumulated sums (conv_igemm stats_out)."""
    n, h, w, c0 = src0.shape
    c1 = 0 if src1 is None else src1.shape[3]
    c = c0 + c1
    ho, wo = (h // 2, w // 2) if resample == RESAMPLE_AVGPOOL2 else ((h * 2, w * 2) if resample == RESAMPLE_NEAREST2 else (h, w))
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=src0.device)
    if stats is None:
        stats = torch.empty((n, 32, 2), dtype=torch.float64, device=src0.device)
    d = _lib.GnDesc()
    d.n, d.h, d.w = n, h, w
    d.src0, d.c0 = _dev(src0, "src0", torch.bfloat16), c0
    d.src1, d.c1 = _opt(src1, "src1", torch.bfloat16), c1
    d.gamma = _dev(gamma, "gamma", torch.float32)
    d.beta = _dev(beta, "beta", torch.float32)
    d.eps = eps
    ss_base = None
    if isinstance(scale_shift, tuple):  # (base tensor, element offset of this block's columns)
        ss_base, ss_off = scale_shift
        d.scale_shift = _dev(ss_base, "scale_shift", torch.float32) + 4 * int(ss_off)
    else:
        d.scale_shift = _opt(scale_shift, "scale_shift", torch.float32)
        ss_base = scale_shift
    d.ss_stride = ss_stride
    d.silu = int(bool(silu))
    d.resample = resample
    d.out = _dev(out, "out", torch.bfloat16)
    d.stats = _dev(stats, "stats", torch.float64)
    d.stats_ready = int(bool(stats_ready))
    _lib.check(_lib.lib().adb_groupnorm(_ph(plan), C.byref(d), _stream()), "adb_groupnorm")
    if plan is not None:
        plan.keep(src0, src1, gamma, beta, ss_base, out, stats)
    return out


def resample2x(src: torch.Tensor, mode: int, out: Optional[torch.Tensor] = None,
               plan: Optional[Plan] = None) -> torch.Tensor:
    n, h, w, c = src.shape
    ho, wo = (h // 2, w // 2) if mode == RESAMPLE_AVGPOOL2 else (h * 2, w * 2)
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=src.device)
    _lib.check(
        _lib.lib().adb_resample2x(_ph(plan), _dev(src, "src", torch.bfloat16), _dev(out, "out", torch.bfloat16),
                                  n, h, w, c, mode, _stream()),
        "adb_resample2x",
    )
    if plan is not None:
        plan.keep(src, out)
    return out


def stem_conv(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
              out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """x fp32 NCHW, weight fp32 [cout,cin,3,3] -> bf16 NHWC."""
    n, cin, h, w = x.shape
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=x.device)
    _lib.check(
        _lib.lib().adb_stem_conv(_ph(plan), _dev(x, "x", torch.float32), _dev(weight, "weight", torch.float32),
                                 _opt(bias, "bias", torch.float32), _dev(out, "out", torch.bfloat16),
                                 n, cin, h, w, cout, _stream()),
        "adb_stem_conv",
    )
    if plan is not None:
        plan.keep(x, weight, bias, out)
    return out


def pack_stem_weight(weight: torch.Tensor, device=None) -> torch.Tensor:
    """Stem conv weight [cout, 3, 3, 3] -> bf16 [cout_pad, 64] = [W (tap-major, channel-minor) | 0 x5 | W | 0 x5],
    the operand matching `adb_stem_im2col`'s [hi | lo] rows."""
    cout, cin = weight.shape[0], weight.shape[1]
    assert cin == 3 and tuple(weight.shape[2:]) == (3, 3)
    w = weight.detach().float().permute(0, 2, 3, 1).reshape(cout, 27)
    z = w.new_zeros(cout, 5)
    m = torch.cat([w, z, w, z], dim=1)
    bn = conv_block_n(cout)
    pad = (-cout) % bn
    if pad:
        m = torch.cat([m, m.new_zeros(pad, 64)], dim=0)
    m = m.to(torch.bfloat16).contiguous()
    return m.to(device) if device is not None else m


def stem_conv_tc(x: torch.Tensor, w_stem: torch.Tensor, bias: Optional[torch.Tensor], cout: int,
                 out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None,
                 stats_out: Optional[torch.Tensor] = None, stats2: Optional[tuple] = None) -> torch.Tensor:
    """x fp32 NCHW [n,3,h,w] -> bf16 NHWC [n,h,w,cout]: im2col (hi | lo split of the input) + one tensor-core k-step;
    the epilogue can accumulate the consumer GroupNorm's sums like any other conv (`stats_out`, `stats2`)."""
    n, cin, h, w = x.shape
    assert cin == 3
    col = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().adb_stem_im2col(_ph(plan), _dev(x, "x", torch.float32), _dev(col, "col", torch.bfloat16),
                                          n, h, w, _stream()), "adb_stem_im2col")
    if plan is not None:
        plan.keep(x, col)
    return conv_igemm([(col, 1)], w_stem, bias, cout, out=out, plan=plan, stats_out=stats_out, stats2=stats2)


_FREQS = {}


def timestep_freqs(dim: int, device, max_period: int = 10000) -> torch.Tensor:
    """The constant table of nn.py:112-114, evaluated once on the host with the reference's own
    expression (fp32) so that t * freqs is bit-identical to the reference's `args`."""
    key = (dim, str(device), max_period)
    if key not in _FREQS:
        half = dim // 2
        f = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
        _FREQS[key] = f.to(device)
    return _FREQS[key]


def timestep_embedding(t: torch.Tensor, dim: int, out: Optional[torch.Tensor] = None,
                       plan: Optional[Plan] = None) -> torch.Tensor:
    b = t.shape[0]
    if out is None:
        out = torch.empty((b, dim), dtype=torch.float32, device=t.device)
    freqs = timestep_freqs(dim, t.device)
    if t.dtype == torch.float32:  # fractional model timesteps (DPM-Solver)
        _lib.check(
            _lib.lib().adb_timestep_embedding_f32(_ph(plan), _dev(t, "t", torch.float32), _dev(freqs, "freqs", torch.float32),
                                                  _dev(out, "out", torch.float32), b, dim, _stream()),
            "adb_timestep_embedding_f32",
        )
    else:
        _lib.check(
            _lib.lib().adb_timestep_embedding(_ph(plan), _dev(t, "t", torch.int64), _dev(freqs, "freqs", torch.float32),
                                              _dev(out, "out", torch.float32), b, dim, _stream()),
            "adb_timestep_embedding",
        )
    if plan is not None:
        plan.keep(t, freqs, out)
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], silu_in: bool = False,
           table: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    b, k = x.shape
    nout = weight.shape[0]
    assert weight.shape[1] == k
    if out is None:
        out = torch.empty((b, nout), dtype=torch.float32, device=x.device)
    _lib.check(
        _lib.lib().adb_linear(_ph(plan), _dev(x, "x", torch.float32), _dev(weight, "weight", torch.float32),
                              _opt(bias, "bias", torch.float32), _dev(out, "out", torch.float32), b, k, nout,
                              int(bool(silu_in)), _opt(table, "table", torch.float32),
                              _opt(idx, "idx", torch.int64), _stream()),
        "adb_linear",
    )
    if plan is not None:
        plan.keep(x, weight, bias, table, idx, out)
    return out


def pack_linear_weight_split(weight: torch.Tensor, device=None) -> torch.Tensor:
    """fp32 Linear weight [nout, k] -> the K-major bf16 operand [nout_pad, 3k] = [W_hi | W_hi | W_lo] of `linear_tc`."""
    w = weight.detach().float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    m = torch.cat([hi, hi, lo], dim=1)
    bn = conv_block_n(w.shape[0])
    pad = (-w.shape[0]) % bn
    if pad:
        m = torch.cat([m, m.new_zeros(pad, m.shape[1])], dim=0)
    m = m.contiguous()
    return m.to(device) if device is not None else m


def linear_tc(x: torch.Tensor, w_split: torch.Tensor, bias: Optional[torch.Tensor], nout: int, silu_in: bool = False,
              out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None, hi_lo: Optional[tuple] = None) -> torch.Tensor:
    """out[b, :] = act(x[b, :]) W^T + bias on tensor cores at fp32-grade accuracy: act(x) is split into bf16
    hi + lo, the weight was split by `pack_linear_weight_split`, and the three cross products that matter run as
    one implicit GEMM over three K-segments with fp32 accumulation and fp32 output. x fp32 [b, k], k % 8 == 0."""
    b, k = x.shape
    assert w_split.shape[1] == 3 * k and k % 8 == 0
    if hi_lo is not None:  # caller-provided scratch for the bf16 hi / lo halves (bf16 [b, 1, 1, k] each)
        hi, lo = hi_lo
        assert tuple(hi.shape) == (b, 1, 1, k) and tuple(lo.shape) == (b, 1, 1, k)
    else:
        hi = torch.empty((b, 1, 1, k), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty_like(hi)
    _lib.check(_lib.lib().adb_split_bf16(_ph(plan), _dev(x, "x", torch.float32), _dev(hi, "hi", torch.bfloat16),
                                         _dev(lo, "lo", torch.bfloat16), b * k, int(bool(silu_in)), _stream()),
               "adb_split_bf16")
    if plan is not None:
        plan.keep(x, hi, lo)
    if out is None:
        out = torch.empty((b, nout), dtype=torch.float32, device=x.device)
    conv_igemm([(hi, 1), (lo, 1), (hi, 1)], w_split, bias, nout, out=out.view(b, nout, 1, 1), out_mode=OUT_F32_NCHW,
               plan=plan)
    return out


def ddim_step(x: torch.Tensor, model_out: torch.Tensor, grad: Optional[torch.Tensor], coef: Sequence[float],
              clip_denoised: bool = True, x_prev: Optional[torch.Tensor] = None,
              pred_xstart: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """x fp32 [n,c,h,w]; model_out fp32 [n,>=c,h,w] (eps in the first c channels)."""
    n, c = x.shape[0], x.shape[1]
    hw = x[0, 0].numel()
    if x_prev is None:
        x_prev = torch.empty_like(x)
    cf = (C.c_float * 5)(*[float(v) for v in coef])
    _lib.check(
        _lib.lib().adb_ddim_step(_ph(plan), _dev(x, "x", torch.float32), _dev(model_out, "model_out", torch.float32),
                                 model_out.shape[1], _opt(grad, "grad", torch.float32),
                                 _dev(x_prev, "x_prev", torch.float32), _opt(pred_xstart, "pred_xstart", torch.float32),
                                 n, c, hw, cf, int(bool(clip_denoised)), _stream()),
        "adb_ddim_step",
    )
    if plan is not None:
        plan.keep(x, model_out, grad, x_prev, pred_xstart)
    return x_prev


def pack_uint8(sample: torch.Tensor, out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    n, c, h, w = sample.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.uint8, device=sample.device)
    _lib.check(
        _lib.lib().adb_pack_uint8(_ph(plan), _dev(sample, "sample", torch.float32), _dev(out, "out", torch.uint8),
                                  n, c, h * w, _stream()),
        "adb_pack_uint8",
    )
    if plan is not None:
        plan.keep(sample, out)
    return out


def moments_accumulate(feats: torch.Tensor, sum_x: torch.Tensor, sum_xx: torch.Tensor,
                       plan: Optional[Plan] = None) -> None:
    n, d = feats.shape
    assert sum_x.shape == (d,) and sum_xx.shape == (d, d)
    _lib.check(
        _lib.lib().adb_moments_accumulate(_ph(plan), _dev(feats, "feats", torch.float32), n, d,
                                          _dev(sum_x, "sum_x", torch.float64), _dev(sum_xx, "sum_xx", torch.float64),
                                          _stream()),
        "adb_moments_accumulate",
    )
    if plan is not None:
        plan.keep(feats, sum_x, sum_xx)


def memset0(t: torch.Tensor, plan: Optional[Plan] = None) -> None:
    _lib.check(_lib.lib().adb_memset0(_ph(plan), _dev(t, "t"), t.numel() * t.element_size(), _stream()), "adb_memset0")
    if plan is not None:
        plan.keep(t)


# layout helpers for tests / API edges (plain torch data movement, not compute)
def nchw_to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nhwc_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    return x.float().permute(0, 3, 1, 2).contiguous()


# ---- Stable-Diffusion-v1 family ----
def attention_sd(q: torch.Tensor, kv: torch.Tensor, b: int, heads: int, d_head: int, d_pad: int, tq: int, tk_rows: int,
                 tk_valid: int, q_col0: int, k_col0: int, v_col0: int, out: Optional[torch.Tensor] = None,
                 plan: Optional[Plan] = None, v_ones: bool = False) -> torch.Tensor:
    """softmax(q k^T d_head^-0.5) v per (batch, head) with heads padded to d_pad columns (include/adb200.h).
    q: bf16 [b*tq, q_width]; kv: bf16 [b*tk_rows, kv_width] (may be `q` itself) -> bf16 [b*tq, heads*d_pad]."""
    q2, kv2 = q.reshape(b * tq, -1), kv.reshape(b * tk_rows, -1)
    if out is None:
        out = torch.empty((b * tq, heads * d_pad), dtype=torch.bfloat16, device=q.device)
    d = _lib.AttnSdDesc()
    d.q, d.q_width, d.q_col0 = _dev(q2, "q", torch.bfloat16), q2.shape[1], q_col0
    d.kv, d.kv_width, d.k_col0, d.v_col0 = _dev(kv2, "kv", torch.bfloat16), kv2.shape[1], k_col0, v_col0
    d.out = _dev(out, "out", torch.bfloat16)
    d.b, d.heads, d.d_head, d.d_pad = b, heads, d_head, d_pad
    d.tq, d.tk_rows, d.tk_valid = tq, tk_rows, tk_valid
    d.v_ones = int(bool(v_ones))
    _lib.check(_lib.lib().adb_attention_sd(_ph(plan), C.byref(d), _stream()), "adb_attention_sd")
    if plan is not None:
        plan.keep(q, kv, out)
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """nn.LayerNorm over the last dim of a bf16 [..., c] tensor."""
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().adb_layernorm(_ph(plan), _dev(x, "x", torch.bfloat16), _dev(gamma, "gamma", torch.float32),
                                        _dev(beta, "beta", torch.float32), _dev(out, "out", torch.bfloat16), rows, c,
                                        float(eps), _stream()), "adb_layernorm")
    if plan is not None:
        plan.keep(x, gamma, beta, out)
    return out


def geglu(x: torch.Tensor, out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """x bf16 [..., 2*inner] -> x[..., :inner] * gelu(x[..., inner:]) (exact GELU)."""
    inner = x.shape[-1] // 2
    rows = x.numel() // (2 * inner)
    if out is None:
        out = torch.empty(x.shape[:-1] + (inner,), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().adb_geglu(_ph(plan), _dev(x, "x", torch.bfloat16), _dev(out, "out", torch.bfloat16), rows, inner,
                                    _stream()), "adb_geglu")
    if plan is not None:
        plan.keep(x, out)
    return out


def cfg_ddim_step(x: torch.Tensor, eps: torch.Tensor, coef: Sequence[float], scale: float = 1.0, cfg: bool = False,
                  x_prev: Optional[torch.Tensor] = None, pred_x0: Optional[torch.Tensor] = None,
                  plan: Optional[Plan] = None) -> torch.Tensor:
    """p_sample_ddim with eta = 0 and optional classifier-free guidance (ddim.py:184-216); eps is [2n, ...] with the
    unconditional half first when cfg. coef = (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) as fp32 values."""
    n = x.shape[0]
    chw = x.numel() // n
    assert eps.numel() == (2 if cfg else 1) * x.numel()
    if x_prev is None:
        x_prev = torch.empty_like(x)
    cf = (C.c_float * 4)(*[float(v) for v in coef])
    _lib.check(_lib.lib().adb_cfg_ddim_step(_ph(plan), _dev(x, "x", torch.float32), _dev(eps, "eps", torch.float32),
                                            _dev(x_prev, "x_prev", torch.float32), _opt(pred_x0, "pred_x0", torch.float32),
                                            n, chw, int(bool(cfg)), float(scale), cf, _stream()), "adb_cfg_ddim_step")
    if plan is not None:
        plan.keep(x, eps, x_prev, pred_x0)
    return x_prev


def pad_context(ctx: torch.Tensor, t_pad: int, out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """fp32 [n, t, c] -> bf16 [n, t_pad, c], rows t.. zero."""
    n, t, c = ctx.shape
    if out is None:
        out = torch.empty((n, t_pad, c), dtype=torch.bfloat16, device=ctx.device)
    _lib.check(_lib.lib().adb_pad_context(_ph(plan), _dev(ctx, "ctx", torch.float32), _dev(out, "out", torch.bfloat16), n, t, c,
                                          t_pad, _stream()), "adb_pad_context")
    if plan is not None:
        plan.keep(ctx, out)
    return out


def cfg_combine(eps: torch.Tensor, scale: float = 1.0, cfg: bool = False, out: Optional[torch.Tensor] = None,
                plan: Optional[Plan] = None) -> torch.Tensor:
    """e_t = e_u + scale * (e_c - e_u) from eps = [uncond | cond] (cfg) or a copy of eps: PLMS keeps it as history."""
    total = eps.numel() // (2 if cfg else 1)
    if out is None:
        out = torch.empty((eps.shape[0] // (2 if cfg else 1),) + tuple(eps.shape[1:]), dtype=torch.float32, device=eps.device)
    assert out.numel() == total
    _lib.check(_lib.lib().adb_cfg_combine(_ph(plan), _dev(eps, "eps", torch.float32), _dev(out, "out", torch.float32), total,
                                          int(bool(cfg)), float(scale), _stream()), "adb_cfg_combine")
    if plan is not None:
        plan.keep(eps, out)
    return out


def plms_update(x: torch.Tensor, e_t: torch.Tensor, old: Sequence[torch.Tensor], mode: int, coef: Sequence[float],
                x_prev: Optional[torch.Tensor] = None, pred_x0: Optional[torch.Tensor] = None,
                plan: Optional[Plan] = None) -> torch.Tensor:
    """p_sample_plms after the model call (include/adb200.h: adb_plms_update). old = (newest, ..., oldest) eps tensors;
    for mode 1 old[0] is e_t_next."""
    if x_prev is None:
        x_prev = torch.empty_like(x)
    o = [(_dev(t, "old", torch.float32) if t is not None else None) for t in (list(old) + [None] * 3)[:3]]
    cf = (C.c_float * 4)(*[float(v) for v in coef])
    _lib.check(_lib.lib().adb_plms_update(_ph(plan), _dev(x, "x", torch.float32), _dev(e_t, "e_t", torch.float32), o[0], o[1], o[2],
                                          int(mode), cf, _dev(x_prev, "x_prev", torch.float32),
                                          _opt(pred_x0, "pred_x0", torch.float32), x.numel(), _stream()), "adb_plms_update")
    if plan is not None:
        plan.keep(x, e_t, *[t for t in old if t is not None], x_prev, pred_x0)
    return x_prev


def dpm_x0(x: torch.Tensor, eps: torch.Tensor, sigma: float, alpha: float, scale: float = 1.0, cfg: bool = False,
           out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """DPM-Solver++ data prediction with classifier-free guidance: x0 = (x - sigma * noise) / alpha."""
    if out is None:
        out = torch.empty_like(x)
    assert eps.numel() == (2 if cfg else 1) * x.numel()
    _lib.check(_lib.lib().adb_dpm_x0(_ph(plan), _dev(x, "x", torch.float32), _dev(eps, "eps", torch.float32),
                                     _dev(out, "out", torch.float32), x.numel(), int(bool(cfg)), float(scale), float(sigma),
                                     float(alpha), _stream()), "adb_dpm_x0")
    if plan is not None:
        plan.keep(x, eps, out)
    return out


def dpm_update(x: torch.Tensor, m0: torch.Tensor, m1: Optional[torch.Tensor], order: int, c0: float, c1: float, c2: float = 0.0,
               inv_r0: float = 0.0, out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """Multistep DPM-Solver++ update of order 1 or 2 (include/adb200.h: adb_dpm_update). m0 = newest data prediction."""
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().adb_dpm_update(_ph(plan), _dev(x, "x", torch.float32), _dev(m0, "m0", torch.float32),
                                         _opt(m1, "m1", torch.float32), _dev(out, "out", torch.float32), x.numel(), int(order),
                                         float(c0), float(c1), float(c2), float(inv_r0), _stream()), "adb_dpm_update")
    if plan is not None:
        plan.keep(x, m0, m1, out)
    return out


# ---- Inception-V3 pool_3 extractor ----
def _sources(srcs):
    """[(bf16 NHWC tensor, relu_on_load), ...] -> ctypes arrays (ptrs, chans, relu), n, h, w, total channels."""
    n, h, w = srcs[0][0].shape[:3]
    k = len(srcs)
    ptrs, chans, relu = (C.c_void_p * k)(), (C.c_int * k)(), (C.c_int * k)()
    for i, (t, r) in enumerate(srcs):
        assert tuple(t.shape[:3]) == (n, h, w), "concatenated sources share n, h, w"
        ptrs[i] = _dev(t, f"src{i}", torch.bfloat16)
        chans[i] = t.shape[3]
        relu[i] = int(bool(r))
    return ptrs, chans, relu, n, h, w, sum(t.shape[3] for t, _ in srcs)


def resize_bilinear_u8(u8: torch.Tensor, oh: int, ow: int, out: Optional[torch.Tensor] = None,
                       plan: Optional[Plan] = None) -> torch.Tensor:
    """uint8 NHWC [n, h, w, 3] -> bf16 NHWC [n, oh, ow, 8] (3 channels + zero padding), bilinear (align_corners=False),
    normalised to [-1, 1]."""
    n, h, w, c = u8.shape
    assert c == 3
    if out is None:
        out = torch.empty((n, oh, ow, 8), dtype=torch.bfloat16, device=u8.device)
    _lib.check(_lib.lib().adb_resize_bilinear_u8(_ph(plan), _dev(u8, "u8", torch.uint8), _dev(out, "out", torch.bfloat16),
                                                 n, h, w, oh, ow, _stream()), "adb_resize_bilinear_u8")
    if plan is not None:
        plan.keep(u8, out)
    return out


def gather_patches(srcs, kh: int, kw: int, stride: int = 1, ph: int = 0, pw: int = 0, out: Optional[torch.Tensor] = None,
                   plan: Optional[Plan] = None) -> torch.Tensor:
    """im2col over a channel concatenation of sources [(tensor, relu_on_load)]: -> bf16 [n*ho*wo, 1, 1, k_pad]
    (the 1-tap operand layout of `conv_igemm`), k_pad = kh*kw*channels rounded up to 8."""
    ptrs, chans, relu, n, h, w, ctot = _sources(srcs)
    ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
    k_pad = (kh * kw * ctot + 7) // 8 * 8
    if out is None:
        out = torch.empty((n * ho * wo, 1, 1, k_pad), dtype=torch.bfloat16, device=srcs[0][0].device)
    assert out.numel() == n * ho * wo * k_pad
    _lib.check(_lib.lib().adb_gather_patches(_ph(plan), ptrs, chans, relu, len(srcs), _dev(out, "out", torch.bfloat16),
                                             n, h, w, kh, kw, stride, ph, pw, k_pad, _stream()), "adb_gather_patches")
    if plan is not None:
        plan.keep(*[t for t, _ in srcs], out)
    return out


def pool3x3(srcs, stride: int, pad: int, mode: int, out: Optional[torch.Tensor] = None,
            plan: Optional[Plan] = None) -> torch.Tensor:
    """mode 0 max / 1 average over 9 / 2 average over in-image taps; -> bf16 [n, ho, wo, channels]."""
    ptrs, chans, relu, n, h, w, ctot = _sources(srcs)
    ho, wo = (h + 2 * pad - 3) // stride + 1, (w + 2 * pad - 3) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, ctot), dtype=torch.bfloat16, device=srcs[0][0].device)
    _lib.check(_lib.lib().adb_pool3x3(_ph(plan), ptrs, chans, relu, len(srcs), _dev(out, "out", torch.bfloat16), n, h, w,
                                      stride, pad, mode, _stream()), "adb_pool3x3")
    if plan is not None:
        plan.keep(*[t for t, _ in srcs], out)
    return out


def global_avgpool(srcs, out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None) -> torch.Tensor:
    """-> fp32 [n, channels]: mean over all pixels (ReLU on load per source)."""
    ptrs, chans, relu, n, h, w, ctot = _sources(srcs)
    if out is None:
        out = torch.empty((n, ctot), dtype=torch.float32, device=srcs[0][0].device)
    _lib.check(_lib.lib().adb_global_avgpool(_ph(plan), ptrs, chans, relu, len(srcs), _dev(out, "out", torch.float32), n,
                                             h * w, _stream()), "adb_global_avgpool")
    if plan is not None:
        plan.keep(*[t for t, _ in srcs], out)
    return out
