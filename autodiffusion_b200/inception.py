"""Inception-V3 pool_3 features on the device, on this library's own kernels.

The reference extracts the 2048-d pool_3 activations with a TensorFlow graph on the host side of a uint8 round trip
(`evaluations/evaluator_v1.py:252-280, 665-679`); the Stable-Diffusion search uses pytorch-fid's InceptionV3
(`"Stable Diffusion"/scripts/search_ea.py:95-127, 171-182`). Here the extractor takes the sampler's uint8 NHWC batch
where it lies in HBM and returns fp32 `[B, 2048]` rows for `CandidateEvaluator`'s moment accumulation - no host copy,
no second framework, no cuDNN:

  * resize to 299 x 299 + normalisation ............ `adb_resize_bilinear_u8`
  * each of the 94 convolutions .................... `adb_gather_patches` (window + channel-concat + ReLU-on-load gather)
                                                     -> `adb_conv_igemm` (tcgen05 GEMM over the patch rows)
    BatchNorm (eval, eps 1e-3) is folded into the packed bf16 weights and the fp32 bias; a conv stores its
    pre-activation and whoever reads it applies the ReLU
  * 3x3 max / average pools ......................... `adb_pool3x3`
  * final 8 x 8 average pool ....................... `adb_global_avgpool`
  * `th.cat` of the Inception branches ............. never materialised: the gather / pool kernels read up to four sources

The module tree only stores parameters, under torchvision's `inception_v3` names (`Conv2d_1a_3x3.conv.weight`,
`Mixed_5b.branch1x1.bn.running_mean`, ...), so pytorch-fid's `pt_inception-2015-12-05` state_dict loads unchanged; with
`fid_variant=True` (default) the graph is the FID Inception of both references: padding-excluding 3x3 average pools in
Mixed_5x / 6x / 7b and a max pool in Mixed_7c's pool branch. No Inception weights exist offline, so by default the
network is He-initialised from a seed and parity is pinned at the activation level against the same graph in fp32 torch
on identical weights (`oracle/inception_ref.py`, `tests/test_inception_gpu.py`).
`split_weights=True` (default) feeds each GEMM the weights as a bf16 hi + lo pair over a doubled K, which removes the
weight-rounding half of the bf16 error (94 layers deep) for <0.3 % of a candidate's FLOPs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch as th
import torch.nn as nn

from . import ops
from .dynamic_unet import _Ctx, _Holder, _Pool

__all__ = ["InceptionPool3"]

_CHUNK = 64  # images per recorded pass: bounds the im2col workspace (~3.2 GB at 147 x 147 x 288 for 64 images)


class BasicConv2d(_Holder):
    """Parameters of torchvision's BasicConv2d: conv (no bias) + BatchNorm2d(eps=0.001)."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0):
        super().__init__()
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        pd = (padding, padding) if isinstance(padding, int) else tuple(padding)
        self.conv = nn.Conv2d(cin, cout, ks, stride=stride, padding=pd, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=0.001)
        self.cin, self.cout, self.ks, self.stride, self.pad = cin, cout, ks, stride, pd


class _Block(_Holder):
    pass


def _inception_a(cin, pool_features):
    b = _Block()
    b.branch1x1 = BasicConv2d(cin, 64, 1)
    b.branch5x5_1 = BasicConv2d(cin, 48, 1)
    b.branch5x5_2 = BasicConv2d(48, 64, 5, padding=2)
    b.branch3x3dbl_1 = BasicConv2d(cin, 64, 1)
    b.branch3x3dbl_2 = BasicConv2d(64, 96, 3, padding=1)
    b.branch3x3dbl_3 = BasicConv2d(96, 96, 3, padding=1)
    b.branch_pool = BasicConv2d(cin, pool_features, 1)
    b.kind = "A"
    return b


def _inception_b(cin):
    b = _Block()
    b.branch3x3 = BasicConv2d(cin, 384, 3, stride=2)
    b.branch3x3dbl_1 = BasicConv2d(cin, 64, 1)
    b.branch3x3dbl_2 = BasicConv2d(64, 96, 3, padding=1)
    b.branch3x3dbl_3 = BasicConv2d(96, 96, 3, stride=2)
    b.kind = "B"
    return b


def _inception_c(cin, c7):
    b = _Block()
    b.branch1x1 = BasicConv2d(cin, 192, 1)
    b.branch7x7_1 = BasicConv2d(cin, c7, 1)
    b.branch7x7_2 = BasicConv2d(c7, c7, (1, 7), padding=(0, 3))
    b.branch7x7_3 = BasicConv2d(c7, 192, (7, 1), padding=(3, 0))
    b.branch7x7dbl_1 = BasicConv2d(cin, c7, 1)
    b.branch7x7dbl_2 = BasicConv2d(c7, c7, (7, 1), padding=(3, 0))
    b.branch7x7dbl_3 = BasicConv2d(c7, c7, (1, 7), padding=(0, 3))
    b.branch7x7dbl_4 = BasicConv2d(c7, c7, (7, 1), padding=(3, 0))
    b.branch7x7dbl_5 = BasicConv2d(c7, 192, (1, 7), padding=(0, 3))
    b.branch_pool = BasicConv2d(cin, 192, 1)
    b.kind = "C"
    return b


def _inception_d(cin):
    b = _Block()
    b.branch3x3_1 = BasicConv2d(cin, 192, 1)
    b.branch3x3_2 = BasicConv2d(192, 320, 3, stride=2)
    b.branch7x7x3_1 = BasicConv2d(cin, 192, 1)
    b.branch7x7x3_2 = BasicConv2d(192, 192, (1, 7), padding=(0, 3))
    b.branch7x7x3_3 = BasicConv2d(192, 192, (7, 1), padding=(3, 0))
    b.branch7x7x3_4 = BasicConv2d(192, 192, 3, stride=2)
    b.kind = "D"
    return b


def _inception_e(cin):
    b = _Block()
    b.branch1x1 = BasicConv2d(cin, 320, 1)
    b.branch3x3_1 = BasicConv2d(cin, 384, 1)
    b.branch3x3_2a = BasicConv2d(384, 384, (1, 3), padding=(0, 1))
    b.branch3x3_2b = BasicConv2d(384, 384, (3, 1), padding=(1, 0))
    b.branch3x3dbl_1 = BasicConv2d(cin, 448, 1)
    b.branch3x3dbl_2 = BasicConv2d(448, 384, 3, padding=1)
    b.branch3x3dbl_3a = BasicConv2d(384, 384, (1, 3), padding=(0, 1))
    b.branch3x3dbl_3b = BasicConv2d(384, 384, (3, 1), padding=(1, 0))
    b.branch_pool = BasicConv2d(cin, 192, 1)
    b.kind = "E"
    return b


Src = List[Tuple[th.Tensor, bool]]  # a (virtual) channel concatenation: [(bf16 NHWC tensor, ReLU on load)]


class _PassPlan:
    """The recorded network for one chunk size: static uint8 input, fp32 [n, 2048] output, one launch plan (+ CUDA graph)."""

    def __init__(self, net: "InceptionPool3", n: int, h: int, w: int):
        dev = net._device()
        self.u8 = th.zeros((n, h, w, 3), dtype=th.uint8, device=dev)
        self.out = th.empty((n, 2048), dtype=th.float32, device=dev)
        self.plan = ops.Plan()
        net._record(self.plan, self.u8, self.out)
        self.launches = self.plan.run()
        th.cuda.current_stream().synchronize()
        self.graph = th.cuda.CUDAGraph()
        with th.cuda.graph(self.graph):
            self.plan.run()


class InceptionPool3(nn.Module):
    def __init__(self, weights: Optional[str] = None, seed: int = 0, fid_variant: bool = True, split_weights: bool = True):
        super().__init__()
        self.fid_variant = fid_variant
        self.split_weights = split_weights
        with th.random.fork_rng(devices=[]):
            th.manual_seed(seed)
            self.Conv2d_1a_3x3 = BasicConv2d(3, 32, 3, stride=2)
            self.Conv2d_2a_3x3 = BasicConv2d(32, 32, 3)
            self.Conv2d_2b_3x3 = BasicConv2d(32, 64, 3, padding=1)
            self.Conv2d_3b_1x1 = BasicConv2d(64, 80, 1)
            self.Conv2d_4a_3x3 = BasicConv2d(80, 192, 3)
            self.Mixed_5b = _inception_a(192, 32)
            self.Mixed_5c = _inception_a(256, 64)
            self.Mixed_5d = _inception_a(288, 64)
            self.Mixed_6a = _inception_b(288)
            self.Mixed_6b = _inception_c(768, 128)
            self.Mixed_6c = _inception_c(768, 160)
            self.Mixed_6d = _inception_c(768, 160)
            self.Mixed_6e = _inception_c(768, 192)
            self.Mixed_7a = _inception_d(768)
            self.Mixed_7b = _inception_e(1280)
            self.Mixed_7c = _inception_e(2048)
            if weights is None:
                # fan-in scaled (He) draws keep the 94 conv layers' activations O(1); BatchNorm stays at identity statistics
                for m in self.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        if weights is not None:
            sd = th.load(weights, map_location="cpu")
            sd = {k: v for k, v in sd.items() if not k.startswith("fc.") and not k.startswith("AuxLogits.")}
            self.load_state_dict(sd, strict=True)
        self.eval()
        self.dim = 2048
        self._packed: Dict[int, tuple] = {}
        self._packed_for = None
        self._plans: Dict[tuple, _PassPlan] = {}
        self._pool: Optional[_Pool] = None
        self.gpu_launches = 0

    def _device(self):
        return self.Conv2d_1a_3x3.conv.weight.device

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._packed, self._plans, self._packed_for = {}, {}, None
        return r

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._packed, self._plans, self._packed_for = {}, {}, None
        return r

    # ---- operand packing: BatchNorm folded, K ordered (tap, channel) to match adb_gather_patches ----
    def _operands(self, m: BasicConv2d, cin_pad: int):
        key = id(m)
        if key not in self._packed:
            dev = self._device()
            w = m.conv.weight.detach().double()
            scale = m.bn.weight.detach().double() / th.sqrt(m.bn.running_var.detach().double() + m.bn.eps)
            bias = (m.bn.bias.detach().double() - m.bn.running_mean.detach().double() * scale).float()
            w = (w * scale[:, None, None, None]).float()
            if cin_pad > w.shape[1]:  # the resized image carries 8 channels (3 real + zero padding)
                w = th.cat([w, w.new_zeros(w.shape[0], cin_pad - w.shape[1], *w.shape[2:])], dim=1)
            if self.split_weights:
                hi = w.to(th.bfloat16).float()
                lo = w - hi
                wp = ops.pack_conv_weight([hi, lo], dev)  # K = [taps x cin | taps x cin]: the patch matrix is passed twice
            else:
                wp = ops.pack_conv_weight([w], dev)
            self._packed[key] = (wp, bias.to(dev).contiguous())
        return self._packed[key]

    # ---- recording ----
    def _record(self, plan: ops.Plan, u8: th.Tensor, out: th.Tensor):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("InceptionPool3 runs on a CUDA device only (autodiffusion_b200 has no CPU path)")
        if self._pool is None or self._pool.device != dev:
            self._pool = _Pool(dev)
        ctx = _Ctx(self._pool, plan)
        n = u8.shape[0]
        plan.keep(u8, out)

        def release(src: Src):
            for t, _ in src:
                ctx.release(t)

        def conv(m: BasicConv2d, src: Src) -> Src:
            """-> the conv's pre-activation as a one-tensor source (ReLU on load). Does not release `src`."""
            _, h, w = src[0][0].shape[:3]
            ctot = sum(t.shape[3] for t, _ in src)
            kh, kw = m.ks
            ph, pw = m.pad
            ho, wo = (h + 2 * ph - kh) // m.stride + 1, (w + 2 * pw - kw) // m.stride + 1
            wp, bias = self._operands(m, ctot)
            k = kh * kw * ctot
            patches = ctx.alloc((n * ho * wo, 1, 1, k))
            ops.gather_patches(src, kh, kw, m.stride, ph, pw, out=patches, plan=plan)
            o = ctx.alloc((n, ho, wo, m.cout))
            segs = [(patches, 1), (patches, 1)] if self.split_weights else [(patches, 1)]
            ops.conv_igemm(segs, wp, bias, m.cout, out=o.view(n * ho * wo, 1, 1, m.cout), plan=plan)
            ctx.release(patches)
            return [(o, True)]

        def pool(src: Src, stride: int, pad: int, mode: int) -> Src:
            _, h, w = src[0][0].shape[:3]
            ctot = sum(t.shape[3] for t, _ in src)
            ho, wo = (h + 2 * pad - 3) // stride + 1, (w + 2 * pad - 3) // stride + 1
            o = ctx.alloc((n, ho, wo, ctot))
            ops.pool3x3(src, stride, pad, mode, out=o, plan=plan)
            return [(o, False)]  # already activated

        def chain(src: Src, *mods) -> Src:
            cur, first = src, True
            for m in mods:
                nxt = conv(m, cur)
                if not first:
                    release(cur)
                cur, first = nxt, False
            return cur

        avg_mode = 2 if self.fid_variant else 1

        def block(b: _Block, x: Src, last: bool = False) -> Src:
            if b.kind == "A":
                bp = pool(x, 1, 1, avg_mode)
                outs = chain(x, b.branch1x1) + chain(x, b.branch5x5_1, b.branch5x5_2) + \
                    chain(x, b.branch3x3dbl_1, b.branch3x3dbl_2, b.branch3x3dbl_3) + chain(bp, b.branch_pool)
                release(bp)
            elif b.kind == "B":
                outs = chain(x, b.branch3x3) + chain(x, b.branch3x3dbl_1, b.branch3x3dbl_2, b.branch3x3dbl_3) + pool(x, 2, 0, 0)
            elif b.kind == "C":
                bp = pool(x, 1, 1, avg_mode)
                outs = chain(x, b.branch1x1) + chain(x, b.branch7x7_1, b.branch7x7_2, b.branch7x7_3) + \
                    chain(x, b.branch7x7dbl_1, b.branch7x7dbl_2, b.branch7x7dbl_3, b.branch7x7dbl_4, b.branch7x7dbl_5) + \
                    chain(bp, b.branch_pool)
                release(bp)
            elif b.kind == "D":
                outs = chain(x, b.branch3x3_1, b.branch3x3_2) + \
                    chain(x, b.branch7x7x3_1, b.branch7x7x3_2, b.branch7x7x3_3, b.branch7x7x3_4) + pool(x, 2, 0, 0)
            else:  # E; Mixed_7c of the FID graph pools with a max
                bp = pool(x, 1, 1, 0 if (last and self.fid_variant) else avg_mode)
                b3 = chain(x, b.branch3x3_1)
                b3o = chain(b3, b.branch3x3_2a) + chain(b3, b.branch3x3_2b)
                release(b3)
                bd = chain(x, b.branch3x3dbl_1, b.branch3x3dbl_2)
                bdo = chain(bd, b.branch3x3dbl_3a) + chain(bd, b.branch3x3dbl_3b)
                release(bd)
                outs = chain(x, b.branch1x1) + b3o + bdo + chain(bp, b.branch_pool)
                release(bp)
            release(x)
            return outs

        def cat(src: Src) -> Src:
            """The gather / pool kernels read at most four sources: a wider concatenation (InceptionE has six branches)
            is merged pairwise by a 1x1 gather, which also applies the pending ReLUs."""
            while len(src) > 4:
                a, b = src[0], src[1]
                c = a[0].shape[3] + b[0].shape[3]
                _, h, w = a[0].shape[:3]
                m = ctx.alloc((n * h * w, 1, 1, c))
                ops.gather_patches([a, b], 1, 1, 1, 0, 0, out=m, plan=plan)
                ctx.release(a[0])
                ctx.release(b[0])
                src = [(m.view(n, h, w, c), False)] + src[2:]
            return src

        img = ctx.alloc((n, 299, 299, 8))
        ops.resize_bilinear_u8(u8, 299, 299, out=img, plan=plan)
        x: Src = [(img, False)]
        for m in (self.Conv2d_1a_3x3, self.Conv2d_2a_3x3, self.Conv2d_2b_3x3):
            nx = conv(m, x)
            release(x)
            x = nx
        nx = pool(x, 2, 0, 0)
        release(x)
        x = nx
        for m in (self.Conv2d_3b_1x1, self.Conv2d_4a_3x3):
            nx = conv(m, x)
            release(x)
            x = nx
        nx = pool(x, 2, 0, 0)
        release(x)
        x = nx
        for name in ("Mixed_5b", "Mixed_5c", "Mixed_5d", "Mixed_6a", "Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e", "Mixed_7a",
                     "Mixed_7b", "Mixed_7c"):
            x = cat(block(getattr(self, name), x, last=(name == "Mixed_7c")))
        ops.global_avgpool(x, out=out, plan=plan)
        release(x)

    def _plan_for(self, n: int, h: int, w: int) -> _PassPlan:
        key = (n, h, w)
        if key not in self._plans:
            with th.no_grad():
                self._plans[key] = _PassPlan(self, n, h, w)
        return self._plans[key]

    @th.no_grad()
    def forward(self, u8: th.Tensor) -> th.Tensor:
        """u8: uint8 [B, H, W, 3] on the device (the sampler's packed images) -> fp32 [B, 2048]."""
        if not isinstance(u8, th.Tensor) or u8.dtype != th.uint8 or u8.dim() != 4 or u8.shape[3] != 3:
            raise ValueError("InceptionPool3 expects uint8 NHWC RGB images")
        if not u8.is_cuda:
            raise RuntimeError("InceptionPool3: input must be a CUDA tensor (autodiffusion_b200 has no CPU path)")
        B, H, W, _ = u8.shape
        out = th.empty((B, 2048), dtype=th.float32, device=u8.device)
        for s in range(0, B, _CHUNK):
            nb = min(_CHUNK, B - s)
            pp = self._plan_for(nb, H, W)
            pp.u8.copy_(u8[s:s + nb])
            pp.graph.replay()
            out[s:s + nb].copy_(pp.out)
            self.gpu_launches += pp.launches
        return out
