"""The noisy classifier and its guidance gradient as one recorded CUDA launch plan.

Drop-in for guided_diffusion/unet.py `EncoderUNetModel` (:685-896, pool="attention") as built by
`create_classifier` (script_util.py:257-295), and for the search's `cond_fn` closure
(search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-390):

    with th.enable_grad():
        x_in = x.detach().requires_grad_(True)
        logits = classifier(x_in, t)
        log_probs = F.log_softmax(logits, dim=-1)
        selected = log_probs[range(len(logits)), y.view(-1)]
        return th.autograd.grad(selected.sum(), x_in)[0] * args.classifier_scale

The reference gets that gradient from autograd over ~600 eager kernels forward plus their backward
twins. Here `EncoderUNetModel.record_guidance` walks the network once and records forward AND the
hand-derived input-gradient (backward-data) pass into one launch plan: convolutions / projections
differentiate to the same tcgen05 implicit GEMM with transposed, flipped weights; GroupNorm + FiLM +
SiLU (+pool), attention and the attention-pool head have their own backward kernels
(csrc/backward.cu, csrc/attention_bwd.cu). Weight gradients are never formed - only d/dx is needed.

`ClassifierGuidance(classifier, scale)` is the cond_fn-compatible callable;
`sampler.SchedulePlan` recognises it and records the guidance of every DDIM step into the
candidate's CUDA graph (no Python, no autograd tape, no host sync per step).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch as th
import torch.nn as nn

from . import ops
from .dynamic_unet import AttentionBlock, ResBlock, _Ctx, _Holder, _Pool, _Seq
from .nn import conv_nd, linear, normalization

__all__ = ["EncoderUNetModel", "ClassifierGuidance"]


class AttentionPool2d(_Holder):
    """Parameters of unet.py AttentionPool2d (:22-40)."""

    def __init__(self, spacial_dim: int, embed_dim: int, num_heads_channels: int, output_dim: Optional[int] = None):
        super().__init__()
        self.positional_embedding = nn.Parameter(th.randn(embed_dim, spacial_dim ** 2 + 1) / embed_dim ** 0.5)
        self.qkv_proj = conv_nd(1, embed_dim, 3 * embed_dim, 1)
        self.c_proj = conv_nd(1, embed_dim, output_dim or embed_dim, 1)
        self.num_heads = embed_dim // num_heads_channels
        self.spacial_dim = spacial_dim
        self.embed_dim = embed_dim


@dataclass
class _PRes:
    w1: th.Tensor
    b1: th.Tensor
    w2: th.Tensor        # conv2 (+ 1x1 skip as a second K-segment)
    b2: th.Tensor
    w1t: th.Tensor       # data-gradient operands
    w2t: th.Tensor
    wst: Optional[th.Tensor]
    g1: th.Tensor
    be1: th.Tensor
    g2: th.Tensor
    be2: th.Tensor
    ss_off: int


@dataclass
class _PAttn:
    g: th.Tensor
    be: th.Tensor
    wqkv: th.Tensor
    bqkv: th.Tensor
    wproj: th.Tensor
    bproj: th.Tensor
    wqkvt: th.Tensor
    wprojt: th.Tensor


class _GuidancePlan:
    """One recorded (+ graph-captured) classifier pass for a fixed batch / resolution / mode."""

    def __init__(self, model: "EncoderUNetModel", B: int, H: int, W: int, scale: Optional[float]):
        dev = model._device()
        self.x_in = th.zeros((B, model.in_channels, H, W), dtype=th.float32, device=dev)
        self.t_in = th.zeros((B,), dtype=th.int64, device=dev)
        self.y_in = th.zeros((B,), dtype=th.int64, device=dev)
        self.grad = th.zeros_like(self.x_in) if scale is not None else None
        self.plan = ops.Plan()
        self.logits = model.record_guidance(self.plan, self.x_in, self.t_in, self.y_in, self.grad, scale)
        self.launches = self.plan.run()  # sets kernel attributes, validates the schedule
        self.graph: Optional[th.cuda.CUDAGraph] = None
        if os.environ.get("ADB_NO_GRAPH", "0") != "1":
            th.cuda.current_stream().synchronize()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                self.plan.run()
            self.graph = g

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.plan.run()


class _AutogradPlan:
    """Forward and input-gradient pass of the classifier as two recorded plans over one private set of activations:
    what `EncoderUNetModel.forward` runs when torch autograd is watching its input."""

    def __init__(self, model: "EncoderUNetModel", B: int, H: int, W: int):
        dev = model._device()
        self.x_in = th.zeros((B, model.in_channels, H, W), dtype=th.float32, device=dev)
        self.t_in = th.zeros((B,), dtype=th.int64, device=dev)
        self.grad = th.zeros_like(self.x_in)
        self.dlogits = th.zeros((B, model.out_channels), dtype=th.float32, device=dev)
        self.fwd, self.bwd = ops.Plan(), ops.Plan()
        self.pool = _Pool(dev)
        self.logits = model.record_guidance(self.fwd, self.x_in, self.t_in, None, self.grad, None, dlogits_in=self.dlogits,
                                            plan_bwd=self.bwd, pool=self.pool)
        self.launches_fwd = self.fwd.run()  # validation run (sets kernel attributes)
        self.launches_bwd = self.bwd.run()
        self.generation = 0  # bumped by every forward: a backward of an older forward would read overwritten activations
        self.g_fwd = self.g_bwd = None
        if os.environ.get("ADB_NO_GRAPH", "0") != "1":
            th.cuda.current_stream().synchronize()
            self.g_fwd, self.g_bwd = th.cuda.CUDAGraph(), th.cuda.CUDAGraph()
            with th.cuda.graph(self.g_fwd):
                self.fwd.run()
            with th.cuda.graph(self.g_bwd):
                self.bwd.run()

    def run_fwd(self):
        self.g_fwd.replay() if self.g_fwd is not None else self.fwd.run()

    def run_bwd(self):
        self.g_bwd.replay() if self.g_bwd is not None else self.bwd.run()


class _ClassifierFunction(th.autograd.Function):
    """logits = classifier(x, t) for torch autograd: backward returns the vector-Jacobian product with respect to x from
    the recorded input-gradient plan (the reference's cond_fn, …progressive.py:383-390, runs `th.autograd.grad(
    selected.sum(), x_in)` through `self.classifier`). Parameters receive no gradient: the evaluator never trains."""

    @staticmethod
    def forward(ctx, x, timesteps, model):
        B, _, H, W = x.shape
        ap = model._autograd_plan(B, H, W)
        ap.x_in.copy_(x)
        ap.t_in.copy_(timesteps)
        ap.run_fwd()
        ap.generation += 1
        ctx.ap, ctx.generation, ctx.model = ap, ap.generation, model
        model.gpu_launches += ap.launches_fwd
        return ap.logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        ap = ctx.ap
        if ap.generation != ctx.generation:
            raise RuntimeError("EncoderUNetModel: backward through a forward whose saved activations were overwritten by "
                               "a later forward of the same batch shape (call backward before the next forward)")
        ap.dlogits.copy_(dlogits)
        ap.run_bwd()
        ctx.model.gpu_launches += ap.launches_bwd
        return ap.grad.clone(), None, None


class EncoderUNetModel(nn.Module):
    """unet.py:685-896. Same constructor arguments and state_dict keys; `forward(x, timesteps)` returns the
    [N, out_channels] logits, `input_gradient(x, timesteps, y, scale)` the guidance gradient."""

    def __init__(
        self,
        image_size,
        in_channels,
        model_channels,
        out_channels,
        num_res_blocks,
        attention_resolutions,
        dropout=0,
        channel_mult=(1, 2, 4, 8),
        conv_resample=True,
        dims=2,
        use_checkpoint=False,
        use_fp16=False,
        num_heads=1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
        pool="adaptive",
    ):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only 2-D classifiers are on the evaluator path")
        if not resblock_updown or not use_scale_shift_norm:
            raise NotImplementedError("create_classifier always sets resblock_updown / use_scale_shift_norm "
                                      "(script_util.py:37-38); other variants are not built")
        if pool != "attention":
            raise NotImplementedError("classifier_pool='attention' is the only pooling any reference script uses "
                                      "(script_util.py:39)")
        if num_head_channels == -1:
            raise NotImplementedError("create_classifier fixes num_head_channels=64 (script_util.py:290)")
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.use_checkpoint = use_checkpoint
        self.dtype = th.float16 if use_fp16 else th.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.use_new_attention_order = use_new_attention_order
        self.pool = pool

        ted = model_channels * 4
        self.time_embed = _Seq({0: linear(model_channels, ted), 2: linear(ted, ted)})
        ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([_Seq({0: conv_nd(2, in_channels, ch, 3, padding=1)})])
        ds = 1
        lid = 0
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, int(mult * model_channels), use_scale_shift_norm, layer_id=lid)]
                lid += 1
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, num_heads=num_heads, num_head_channels=num_head_channels,
                                                 use_new_attention_order=use_new_attention_order, layer_id=lid))
                    lid += 1
                self.input_blocks.append(_Seq(dict(enumerate(layers))))
            if level != len(channel_mult) - 1:
                self.input_blocks.append(_Seq({0: ResBlock(ch, ted, ch, use_scale_shift_norm, down=True, layer_id=lid)}))
                lid += 1
                ds *= 2
        self.middle_block = _Seq({
            0: ResBlock(ch, ted, ch, use_scale_shift_norm, layer_id=lid),
            1: AttentionBlock(ch, num_heads=num_heads, num_head_channels=num_head_channels,
                              use_new_attention_order=use_new_attention_order, layer_id=lid + 1),
            2: ResBlock(ch, ted, ch, use_scale_shift_norm, layer_id=lid + 2),
        })
        self.out = _Seq({0: normalization(ch), 2: AttentionPool2d(image_size // ds, ch, num_head_channels, out_channels)})
        self._final_ch = ch

        self._generation = 0
        self._packed_generation = -1
        self._packed: Dict[object, object] = {}
        self._plans: Dict[tuple, _GuidancePlan] = {}
        self._pool: Optional[_Pool] = None
        self.gpu_launches = 0

    # ---- API parity helpers (fp16_util.convert_to_fp16 is a torso-only cast in the reference) ----
    def convert_to_fp16(self):
        self.dtype = th.float16

    def convert_to_fp32(self):
        self.dtype = th.float32

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self._invalidate()
        return r

    def load_state_dict(self, *args, **kwargs):
        r = super().load_state_dict(*args, **kwargs)
        self._invalidate()
        return r

    def refresh_weights(self):
        self._invalidate()

    def _invalidate(self):
        if hasattr(self, "_generation"):
            self._generation += 1

    def _device(self):
        return self.out[0].weight.device

    def _iter_layers(self):
        for blk in list(self.input_blocks)[1:]:
            yield from blk.children()
        yield from self.middle_block.children()

    # ---- operand packing: forward matrices and their data-gradient twins ----
    def _pack(self):
        dev = self._device()
        f32 = lambda p: p.detach().to(device=dev, dtype=th.float32).contiguous()
        P: Dict[object, object] = {}
        emb_w, emb_b, off = [], [], 0
        for layer in self._iter_layers():
            if isinstance(layer, ResBlock):
                c1, c2 = getattr(layer.in_layers, "2"), getattr(layer.out_layers, "3")
                n1, n2 = getattr(layer.in_layers, "0"), getattr(layer.out_layers, "0")
                el = getattr(layer.emb_layers, "1")
                has_skip = isinstance(layer.skip_connection, nn.Conv2d)
                mats, b2 = [c2.weight], c2.bias.detach().float()
                wst = None
                if has_skip:
                    mats.append(layer.skip_connection.weight)
                    b2 = b2 + layer.skip_connection.bias.detach().float()
                    wst = ops.pack_conv_weight_dgrad(layer.skip_connection.weight, dev)
                P[id(layer)] = _PRes(
                    w1=ops.pack_conv_weight([c1.weight], dev), b1=f32(c1.bias),
                    w2=ops.pack_conv_weight(mats, dev), b2=b2.to(dev).contiguous(),
                    w1t=ops.pack_conv_weight_dgrad(c1.weight, dev), w2t=ops.pack_conv_weight_dgrad(c2.weight, dev), wst=wst,
                    g1=f32(n1.weight), be1=f32(n1.bias), g2=f32(n2.weight), be2=f32(n2.bias), ss_off=off)
                emb_w.append(el.weight.detach().float())
                emb_b.append(el.bias.detach().float())
                off += el.weight.shape[0]
            else:
                P[id(layer)] = _PAttn(
                    g=f32(layer.norm.weight), be=f32(layer.norm.bias),
                    wqkv=ops.pack_conv_weight([layer.qkv.weight], dev), bqkv=f32(layer.qkv.bias),
                    wproj=ops.pack_conv_weight([layer.proj_out.weight], dev), bproj=f32(layer.proj_out.bias),
                    wqkvt=ops.pack_conv_weight_dgrad(layer.qkv.weight, dev),
                    wprojt=ops.pack_conv_weight_dgrad(layer.proj_out.weight, dev))
        P["emb_w"] = ops.pack_linear_weight_split(th.cat(emb_w, 0), dev)
        P["emb_b"] = th.cat(emb_b, 0).to(dev).contiguous()
        P["emb_total"] = off
        te0, te2 = getattr(self.time_embed, "0"), getattr(self.time_embed, "2")
        P["te0_w"], P["te0_b"], P["te2_w"], P["te2_b"] = f32(te0.weight), f32(te0.bias), f32(te2.weight), f32(te2.bias)
        stem = getattr(self.input_blocks[0], "0")
        P["stem_w"], P["stem_b"] = f32(stem.weight), f32(stem.bias)
        P["stem_wt"] = ops.pack_conv_weight_dgrad(stem.weight, dev)
        P["stem_wp"] = ops.pack_stem_weight(stem.weight, dev) if self.in_channels == 3 else None
        on, pool = getattr(self.out, "0"), getattr(self.out, "2")
        C = self._final_ch
        P["out_g"], P["out_be"] = f32(on.weight), f32(on.bias)
        P["pos"] = f32(pool.positional_embedding)
        wq = pool.qkv_proj.weight.detach()  # [3C, C, 1]: rows (q | k | v), new attention order
        P["pool_wqkv"] = ops.pack_linear_weight_split(wq[:, :, 0], dev)        # split-bf16 operands of ops.linear_tc
        P["pool_bqkv"] = f32(pool.qkv_proj.bias)
        P["pool_wqkv_t"] = ops.pack_linear_weight_split(wq[:, :, 0].t(), dev)
        P["pool_wkv"] = ops.pack_conv_weight([wq[C:]], dev)
        P["pool_bkv"] = f32(pool.qkv_proj.bias[C:])
        P["pool_wkv_t"] = ops.pack_conv_weight_dgrad(wq[C:], dev)
        P["pool_wc"] = ops.pack_linear_weight_split(pool.c_proj.weight[:, :, 0], dev)
        P["pool_bc"] = f32(pool.c_proj.bias)
        P["pool_wc_t"] = ops.pack_linear_weight_split(pool.c_proj.weight[:, :, 0].t(), dev)
        P["n_out"] = pool.c_proj.weight.shape[0]
        self._packed = P
        self._packed_generation = self._generation
        self._plans.clear()

    # ---- recording ----
    def record_guidance(self, plan: ops.Plan, x_in: th.Tensor, t_in: th.Tensor, y_in: Optional[th.Tensor],
                        grad_out: Optional[th.Tensor], scale: Optional[float], dlogits_in: Optional[th.Tensor] = None,
                        plan_bwd: Optional[ops.Plan] = None, pool: Optional[_Pool] = None) -> th.Tensor:
        """Record logits = classifier(x_in, t_in) and, when `grad_out` is given,
        grad_out = d(log_softmax(logits)[range(B), y_in].sum() * scale) / d x_in into `plan`.
        x_in fp32 NCHW, t_in / y_in int64 [B], grad_out fp32 NCHW. Returns the (static) logits tensor.

        `dlogits_in` (fp32 [B, out_channels]): instead of the log-softmax selection, back-propagate this caller-filled
        d(loss)/d(logits) - the vector-Jacobian product torch autograd asks for (`forward` under `requires_grad`).
        `plan_bwd`: record the input-gradient pass into a second plan (it then runs when autograd calls backward);
        `pool`: a private activation pool, so that nothing recorded elsewhere can reuse the saved activations between the
        two runs."""
        if self._device().type != "cuda":
            raise RuntimeError("EncoderUNetModel runs on a CUDA device only (autodiffusion_b200 has no CPU path)")
        if self._packed_generation != self._generation:
            self._pack()
        want_grad = grad_out is not None
        if want_grad:
            assert dlogits_in is not None or (y_in is not None and scale is not None)
        B, _, H, W = x_in.shape
        dev = self._device()
        P = self._packed
        if pool is None:
            if self._pool is None or self._pool.device != dev:
                self._pool = _Pool(dev)
            pool = self._pool
        ctx = _Ctx(pool, plan)
        mc = self.model_channels
        plan.keep(x_in, t_in, y_in, grad_out)

        n_layers = sum(1 for _ in self._iter_layers())
        arena = th.empty((2 * n_layers + 4, B, 32, 2), dtype=th.float64, device=dev)   # forward GroupNorm sums
        bscratch = th.empty((B, 32, 2), dtype=th.float64, device=dev)                   # backward sums (reused)
        plan.keep(arena, bscratch)
        ops.memset0(arena, plan=plan)
        slot = [0]
        stats_of: Dict[int, th.Tensor] = {}

        def new_slot(t: th.Tensor) -> th.Tensor:
            st = arena[slot[0]]
            slot[0] += 1
            stats_of[t.data_ptr()] = st
            return st

        ctx.on_alloc = lambda t: stats_of.pop(t.data_ptr(), None)  # a recycled buffer has no sums yet

        def conv_stats(t: th.Tensor) -> dict:
            """kwargs making the producing conv accumulate `t`'s GroupNorm sums in its epilogue."""
            if t.shape[3] % 32 != 0 or (t.shape[1] * t.shape[2]) % 32 != 0:
                return {}
            return {"stats_out": new_slot(t)}

        fuse_gnb = os.environ.get("ADB_NO_GNB_FUSE", "0") != "1"

        def gnb_of(x: th.Tensor, st: th.Tensor, gamma, beta, silu: bool, **kw) -> Optional[dict]:
            """`conv_igemm(gnb=...)` descriptor: the data-gradient conv producing d(GroupNorm(x) output) also reduces the
            two sums of that GroupNorm's backward (into `bscratch`), which then reads x and the gradient once."""
            if not fuse_gnb or not ops.conv_gnb_supported(*x.shape):
                return None
            return dict(x=x, stats=st, gamma=gamma, beta=beta, silu=silu, bstats=bscratch, **kw)

        def gn(x, gamma, beta, out_t, **kw) -> th.Tensor:
            """GroupNorm of x; returns the [B,32,2] sums it normalised with (kept for the backward pass)."""
            st = stats_of.get(x.data_ptr())
            ready = st is not None
            if not ready:
                st = new_slot(x)  # the groupnorm op zeroes and fills it
            ops.groupnorm(x, gamma, beta, out=out_t, stats=st, stats_ready=ready, plan=plan, **kw)
            return st

        te = ops.timestep_embedding(t_in, mc, plan=plan)
        e1 = ops.linear(te, P["te0_w"], P["te0_b"], plan=plan)
        emb = ops.linear(e1, P["te2_w"], P["te2_b"], silu_in=True, plan=plan)  # no label embedding (unet.py:871)
        ss_all = ops.linear_tc(emb, P["emb_w"], P["emb_b"], P["emb_total"], silu_in=True, plan=plan)
        ss_total = P["emb_total"]

        tape: List[Callable[[th.Tensor], th.Tensor]] = []  # backward closures, run in reverse

        def res_fwd(layer: ResBlock, x: th.Tensor) -> th.Tensor:
            pk: _PRes = P[id(layer)]
            n, h, w, cin = x.shape
            cout = layer.out_channels
            down = layer.down
            ho, wo = (h // 2, w // 2) if down else (h, w)
            mode = ops.RESAMPLE_AVGPOOL2 if down else ops.RESAMPLE_NONE
            g1 = ctx.alloc((n, ho, wo, cin))
            st_x = gn(x, pk.g1, pk.be1, g1, silu=True, resample=mode)
            c1 = ctx.alloc((n, ho, wo, cout))
            ops.conv_igemm([(g1, 9)], pk.w1, pk.b1, cout, out=c1, plan=plan, **conv_stats(c1))
            ctx.release(g1)
            g2 = ctx.alloc((n, ho, wo, cout))
            st_c1 = gn(c1, pk.g2, pk.be2, g2, scale_shift=(ss_all, pk.ss_off), ss_stride=ss_total, silu=True)
            if not want_grad:
                ctx.release(c1)
            out = ctx.alloc((n, ho, wo, cout))
            if pk.wst is not None:
                ops.conv_igemm([(g2, 9), (x, 1)], pk.w2, pk.b2, cout, out=out, plan=plan, **conv_stats(out))
            else:
                ops.conv_igemm([(g2, 9)], pk.w2, pk.b2, cout, out=out, residual=x,
                               res_mode=ops.RES_AVGPOOL2 if down else ops.RES_SAME, plan=plan, **conv_stats(out))
            ctx.release(g2)

            def bwd(dout: th.Tensor) -> th.Tensor:
                dg2 = ctx.alloc((n, ho, wo, cout))
                fz = gnb_of(c1, st_c1, pk.g2, pk.be2, True, scale_shift=(ss_all, pk.ss_off), ss_stride=ss_total)
                ops.conv_igemm([(dout, 9)], pk.w2t, None, cout, out=dg2, plan=plan, gnb=fz)
                dc1 = ctx.alloc((n, ho, wo, cout))
                ops.gn_backward(c1, st_c1, pk.g2, pk.be2, dg2, scale_shift=(ss_all, pk.ss_off),
                                ss_stride=ss_total, silu=True, dx=dc1, bstats=bscratch, plan=plan,
                                bstats_ready=fz is not None)
                ctx.release(dg2)
                ctx.release(c1)
                dg1 = ctx.alloc((n, ho, wo, cin))
                fz = None if down else gnb_of(x, st_x, pk.g1, pk.be1, True)
                ops.conv_igemm([(dc1, 9)], pk.w1t, None, cin, out=dg1, plan=plan, gnb=fz)
                ctx.release(dc1)
                if pk.wst is not None:  # skip path: x -> 1x1 conv
                    add = ctx.alloc((n, h, w, cin))
                    ops.conv_igemm([(dout, 1)], pk.wst, None, cin, out=add, plan=plan)
                    add_mode = ops.RES_SAME
                else:                   # identity (after the average pool in a down block)
                    add = dout
                    ctx.retain(dout)
                    add_mode = ops.RES_AVGPOOL2 if down else ops.RES_SAME
                dx = ctx.alloc((n, h, w, cin))
                ops.gn_backward(x, st_x, pk.g1, pk.be1, dg1, silu=True, resample=mode, add=add,
                                add_mode=add_mode, dx=dx, bstats=bscratch, plan=plan, bstats_ready=fz is not None)
                ctx.release(dg1)
                ctx.release(add)
                ctx.release(x)
                return dx

            if want_grad:
                tape.append(bwd)
            return out

        def attn_fwd(layer: AttentionBlock, x: th.Tensor) -> th.Tensor:
            pk: _PAttn = P[id(layer)]
            n, h, w, c = x.shape
            t = h * w
            heads = layer.num_heads
            legacy = not layer.use_new_attention_order
            g = ctx.alloc((n, h, w, c))
            st_x = gn(x, pk.g, pk.be, g, silu=False)
            qkv = ctx.alloc((n, h, w, 3 * c))
            ops.conv_igemm([(g, 1)], pk.wqkv, pk.bqkv, 3 * c, out=qkv, plan=plan)
            ctx.release(g)
            a = ctx.alloc((n, h, w, c))
            lse = ctx.alloc((n * heads, t), dtype=th.float32) if want_grad else None
            ops.attention(qkv.view(n * t, 3 * c), n, t, heads, legacy, out=a.view(n * t, c), lse=lse, plan=plan)
            out = ctx.alloc((n, h, w, c))
            ops.conv_igemm([(a, 1)], pk.wproj, pk.bproj, c, out=out, residual=x, res_mode=ops.RES_SAME, plan=plan,
                           **conv_stats(out))

            def bwd(dout: th.Tensor) -> th.Tensor:
                da = ctx.alloc((n, h, w, c))
                ops.conv_igemm([(dout, 1)], pk.wprojt, None, c, out=da, plan=plan)
                dqkv = ctx.alloc((n, h, w, 3 * c))
                dsum = ctx.alloc((n * heads, t), dtype=th.float32)
                ws = ctx.alloc((n * t, c), dtype=th.float32) if t % 128 == 0 else None  # dQ partials of the single-pass kernel
                ops.attention_backward(qkv.view(n * t, 3 * c), a.view(n * t, c), da.view(n * t, c), lse, n, t, heads, legacy,
                                       dqkv=dqkv.view(n * t, 3 * c), dsum=dsum, plan=plan, dq_ws=ws)
                for tns in (da, dsum, a, qkv, lse) + ((ws,) if ws is not None else ()):
                    ctx.release(tns)
                dg = ctx.alloc((n, h, w, c))
                fz = gnb_of(x, st_x, pk.g, pk.be, False)
                ops.conv_igemm([(dqkv, 1)], pk.wqkvt, None, c, out=dg, plan=plan, gnb=fz)
                ctx.release(dqkv)
                dx = ctx.alloc((n, h, w, c))
                ops.gn_backward(x, st_x, pk.g, pk.be, dg, silu=False, add=dout, add_mode=ops.RES_SAME,
                                dx=dx, bstats=bscratch, plan=plan, bstats_ready=fz is not None)
                ctx.release(dg)
                ctx.release(x)
                return dx

            if want_grad:
                tape.append(bwd)
            else:
                ctx.release(a)
                ctx.release(qkv)
            return out

        # ---------------- forward ----------------
        ch0 = int(self.channel_mult[0] * mc)
        h = ctx.alloc((B, H, W, ch0))
        if P["stem_wp"] is not None:
            ops.stem_conv_tc(x_in, P["stem_wp"], P["stem_b"], ch0, out=h, plan=plan, **conv_stats(h))
        else:
            ops.stem_conv(x_in, P["stem_w"], P["stem_b"], out=h, plan=plan)
        for layer in self._iter_layers():
            nxt = res_fwd(layer, h) if isinstance(layer, ResBlock) else attn_fwd(layer, h)
            if not want_grad:
                ctx.release(h)
            h = nxt
        n, hh, ww, C = h.shape
        g = ctx.alloc((n, hh, ww, C))
        st_h = gn(h, P["out_g"], P["out_be"], g, silu=True)
        xp, mean = ops.pool_prepare(g, P["pos"], plan=plan)
        ctx.release(g)
        kv = ops.conv_igemm([(xp, 1)], P["pool_wkv"], P["pool_bkv"], 2 * C, plan=plan)
        qkv0 = ops.linear_tc(mean, P["pool_wqkv"], P["pool_bqkv"], 3 * C, plan=plan)
        out0, probs = ops.pool_attention(qkv0, kv, plan=plan)
        logits = ops.linear_tc(out0, P["pool_wc"], P["pool_bc"], P["n_out"], plan=plan)
        if not want_grad:
            return logits

        # ---------------- backward (data gradients only) ----------------
        if plan_bwd is not None:  # the closures above read `plan` when they are called: from here on, the second plan
            plan = plan_bwd
            ctx.plan = plan_bwd
            plan.keep(x_in, t_in, grad_out, arena, bscratch, dlogits_in)
        dlog = dlogits_in if dlogits_in is not None else ops.logsoftmax_grad(logits, y_in, scale, plan=plan)
        dout0 = ops.linear_tc(dlog, P["pool_wc_t"], None, C, plan=plan)
        dqkv0, dkv = ops.pool_attention_backward(dout0, probs, qkv0, kv, plan=plan)
        dmean = ops.linear_tc(dqkv0, P["pool_wqkv_t"], None, C, plan=plan)
        dxp = ops.conv_igemm([(dkv, 1)], P["pool_wkv_t"], None, C, plan=plan)
        dg = ops.pool_merge(dxp, dmean, plan=plan)
        d = ctx.alloc((n, hh, ww, C))
        ops.gn_backward(h, st_h, P["out_g"], P["out_be"], dg, silu=True, dx=d, bstats=bscratch, plan=plan)
        ctx.release(h)
        for bwd in reversed(tape):
            nd = bwd(d)
            ctx.release(d)
            d = nd
        # d = gradient w.r.t. the stem's bf16 NHWC output; the stem's data gradient lands as fp32 NCHW
        ops.conv_igemm([(d, 9)], P["stem_wt"], None, self.in_channels, out=grad_out, out_mode=ops.OUT_F32_NCHW, plan=plan)
        ctx.release(d)
        return logits

    # ---- public calls ----
    def _plan_for(self, B: int, H: int, W: int, scale: Optional[float]) -> _GuidancePlan:
        if self._device().type != "cuda":
            raise RuntimeError("EncoderUNetModel runs on a CUDA device only (autodiffusion_b200 has no CPU path)")
        if self._packed_generation != self._generation:
            self._pack()
        key = (B, H, W, None if scale is None else float(scale))
        gp = self._plans.get(key)
        if gp is None:
            with th.no_grad():
                gp = _GuidancePlan(self, B, H, W, scale)
            self._plans[key] = gp
        return gp

    def _autograd_plan(self, B: int, H: int, W: int) -> _AutogradPlan:
        if self._packed_generation != self._generation:
            self._pack()
        key = (B, H, W, "autograd")
        ap = self._plans.get(key)
        if ap is None:
            with th.no_grad():
                ap = _AutogradPlan(self, B, H, W)
            self._plans[key] = ap
        return ap

    def forward(self, x, timesteps):
        """unet.py:861-896: [N, C, H, W] fp32, timesteps [N] -> logits [N, out_channels]. When autograd is recording and
        `x` requires grad, the result carries a graph whose backward is the recorded input-gradient pass."""
        if not x.is_cuda:
            raise RuntimeError("EncoderUNetModel.forward: input must be a CUDA tensor (no CPU path)")
        B, _, H, W = x.shape
        assert timesteps.shape == (B,)
        tr = self.__dict__.get("_trace")
        if tr is not None:  # fastpath.py: record the call, run nothing, return logits that report what flows back
            from .fastpath import _TracedLogits

            rec = {}
            tr.append((x, timesteps, rec))
            return _TracedLogits.apply(x, self.out_channels, rec)
        if th.is_grad_enabled() and x.requires_grad:
            return _ClassifierFunction.apply(x.float().contiguous(), timesteps, self)
        gp = self._plan_for(B, H, W, None)
        gp.x_in.copy_(x)
        gp.t_in.copy_(timesteps)
        gp.replay()
        self.gpu_launches += gp.launches
        return gp.logits.clone()

    def input_gradient(self, x, timesteps, y, scale: float = 1.0):
        """grad_x [log_softmax(self(x, t))[range(N), y].sum()] * scale — the search's cond_fn
        (…progressive.py:383-390) without autograd."""
        if not x.is_cuda:
            raise RuntimeError("EncoderUNetModel.input_gradient: input must be a CUDA tensor (no CPU path)")
        B, _, H, W = x.shape
        assert timesteps.shape == (B,) and y.shape == (B,)
        gp = self._plan_for(B, H, W, scale)
        gp.x_in.copy_(x)
        gp.t_in.copy_(timesteps)
        gp.y_in.copy_(y)
        gp.replay()
        self.gpu_launches += gp.launches
        return gp.grad.clone()


class ClassifierGuidance:
    """cond_fn(x, t, y=None, **kwargs) -> grad log p(y | x_t) * classifier_scale, as the search scripts define
    it (…progressive.py:383-390; scripts/classifier_sample_prunedUNET.py:139-149). Callable anywhere a
    reference cond_fn is; `sampler.SchedulePlan` additionally fuses it into the candidate's CUDA graph."""

    def __init__(self, classifier: EncoderUNetModel, classifier_scale: float = 1.0):
        self.classifier = classifier
        self.classifier_scale = float(classifier_scale)

    def __call__(self, x, t, y=None, **_):
        assert y is not None
        return self.classifier.input_gradient(x, t, y, self.classifier_scale)

    def record(self, plan: ops.Plan, x_in, t_in, y_in, grad_out):
        return self.classifier.record_guidance(plan, x_in, t_in, y_in, grad_out, self.classifier_scale)

    def shared_plan(self, x_in, t_in, y_in, grad_out) -> ops.Plan:
        """The recorded forward + input-gradient pass over these exact buffers, recorded once and reused by every
        candidate's schedule (all SchedulePlans of one model geometry share x / t / y / grad buffers)."""
        clf = self.classifier
        if clf._packed_generation != clf._generation:
            clf._pack()  # clears clf._plans; the shared plans below are dropped with them
            clf._shared = {}
        key = (x_in.data_ptr(), t_in.data_ptr(), y_in.data_ptr(), grad_out.data_ptr(), tuple(x_in.shape), self.classifier_scale)
        shared = clf.__dict__.setdefault("_shared", {})
        plan = shared.get(key)
        if plan is None:
            with th.no_grad():
                plan = ops.Plan()
                self.record(plan, x_in, t_in, y_in, grad_out)
                plan.run()  # first run outside any capture: sets kernel attributes, validates the schedule
            shared[key] = plan
        return plan
