"""ADM UNet with a per-call block-skip list, executed as a recorded CUDA launch plan.

Drop-in for guided_diffusion/dynamic_unet.py `Dynamic_UNetModel` (constructor :447-655,
`forward(x, timesteps, y=None, skip_layer=[])` :673-702) and guided_diffusion/unet.py
`UNetModel` (:396-665): identical constructor arguments, identical `state_dict()` keys and
shapes (checkpoints such as 64x64_diffusion.pt load unchanged), `layer_num`,
`convert_to_fp16()`.

The nn.Module tree only *stores* parameters. `forward` never calls those modules: for each
(batch, resolution, skip set) it walks the block list once — ResBlock / AttentionBlock
semantics of :245-271 and :316-325, skipped blocks elided at this point — records the
resulting kernel sequence into an `adb_plan` (C-ABI), captures it in a CUDA graph and
replays it. Activations are bf16 NHWC; GEMM-shaped work (3x3/1x1 convs, qkv/proj) runs on
tcgen05 tensor cores with fp32 accumulation; GroupNorm statistics, the timestep-embedding
MLP and softmax stay fp32 (as the reference keeps them, fp16_util.py:15-22, nn.py:17-19).

Fusions relative to the reference's op list:
  * GroupNorm + (1+scale)*x+shift + SiLU + avg-pool/nearest-upsample  -> one kernel pair
  * th.cat([h, hs.pop()], 1) is never materialised: GroupNorm and the 1x1 skip conv read two sources
  * out_layers conv + skip_connection 1x1 conv + residual add        -> one implicit GEMM
  * all 36 emb_layers Linear(768 -> 2C) products                       -> one GEMM per forward
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch as th
import torch.nn as nn

from . import ops
from .nn import conv_nd, linear, normalization, zero_module

__all__ = ["Dynamic_UNetModel", "UNetModel"]


# ------------------------------------------------------------------------------------------
# parameter containers: same attribute names as the reference modules so state_dict matches
# ------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    """A module that only owns parameters; it is never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; compute runs through the CUDA launch plan")


class _Seq(_Holder):
    """Sparse nn.Sequential look-alike: children registered under their reference indices."""

    def __init__(self, children: Dict[int, nn.Module]):
        super().__init__()
        for i, m in children.items():
            self.add_module(str(i), m)

    def __getitem__(self, i: int) -> nn.Module:
        return getattr(self, str(i))


class ResBlock(_Holder):
    """Parameters of dynamic_unet.py ResBlock (:150-231), created in the reference's order."""

    def __init__(self, channels, emb_channels, out_channels, use_scale_shift_norm, up=False, down=False,
                 layer_id=-1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.up, self.down = up, down
        self.layer_id = layer_id
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = _Seq({0: normalization(channels), 2: conv_nd(2, channels, self.out_channels, 3, padding=1)})
        self.emb_layers = _Seq({1: linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels)})
        self.out_layers = _Seq({
            0: normalization(self.out_channels),
            3: zero_module(conv_nd(2, self.out_channels, self.out_channels, 3, padding=1)),
        })
        if self.out_channels != channels:
            self.skip_connection = conv_nd(2, channels, self.out_channels, 1)
        else:
            self.skip_connection = nn.Identity()


class AttentionBlock(_Holder):
    """Parameters of dynamic_unet.py AttentionBlock (:274-311)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_new_attention_order=False, layer_id=-1):
        super().__init__()
        self.channels = channels
        self.layer_id = layer_id
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, (
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}")
            self.num_heads = channels // num_head_channels
        self.use_new_attention_order = use_new_attention_order
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))


# ------------------------------------------------------------------------------------------
# activation pool + recording context
# ------------------------------------------------------------------------------------------
class _Pool:
    """Reuses activation buffers by byte size; safe because ops run in recorded stream order."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[int, List[th.Tensor]] = {}
        self.total_bytes = 0

    def take(self, nbytes: int) -> th.Tensor:
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        self.total_bytes += nbytes
        return th.empty(nbytes, dtype=th.uint8, device=self.device)

    def give(self, raw: th.Tensor):
        self.free.setdefault(raw.numel(), []).append(raw)


class _Ctx:
    def __init__(self, pool: _Pool, plan: ops.Plan):
        self.pool = pool
        self.plan = plan
        self.raw: Dict[int, th.Tensor] = {}   # data_ptr -> raw buffer
        self.refs: Dict[int, int] = {}
        self.on_alloc = None  # hook(t): called for every fresh (possibly recycled) buffer

    def alloc(self, shape, dtype=th.bfloat16) -> th.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * th.empty((), dtype=dtype).element_size()
        nbytes = (nbytes + 255) // 256 * 256
        raw = self.pool.take(nbytes)
        t = raw[: n * th.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        self.raw[t.data_ptr()] = raw
        self.refs[t.data_ptr()] = 1
        self.plan.keep(raw)
        if self.on_alloc is not None:
            self.on_alloc(t)
        return t

    def retain(self, t: th.Tensor):
        if t.data_ptr() in self.refs:
            self.refs[t.data_ptr()] += 1

    def release(self, t: th.Tensor):
        p = t.data_ptr()
        if p not in self.refs:
            return
        self.refs[p] -= 1
        if self.refs[p] == 0:
            del self.refs[p]
            self.pool.give(self.raw.pop(p))


@dataclass
class _PackedRes:
    w1: th.Tensor
    b1: th.Tensor
    w2_raw: th.Tensor       # conv2 weight [cout,cout,3,3]; packed per concat split in _w2_for
    b2: th.Tensor           # b_conv2 (+ b_skip when the 1x1 skip is folded into the same GEMM)
    ws_raw: Optional[th.Tensor]  # 1x1 skip weight [cout,cin,1,1] or None (identity skip)
    bskip: Optional[th.Tensor]
    g1: th.Tensor
    be1: th.Tensor
    g2: th.Tensor
    be2: th.Tensor
    ss_off: int


@dataclass
class _PackedAttn:
    g: th.Tensor
    be: th.Tensor
    wqkv: th.Tensor
    bqkv: th.Tensor
    wproj: th.Tensor
    bproj: th.Tensor


@dataclass
class _IO:
    x_in: th.Tensor
    t_in: th.Tensor
    y_in: Optional[th.Tensor]
    out: th.Tensor


class _UNetPlan:
    """One recorded + graph-captured forward for a fixed (batch, H, W, skip set)."""

    def __init__(self, model: "Dynamic_UNetModel", B: int, H: int, W: int, skip: Tuple[int, ...]):
        # every plan of one (B, H, W) shares the same input/output buffers, so a searched schedule can
        # chain the cached per-mask graphs without copies (sampler.SchedulePlan)
        io = model.io_buffers(B, H, W)
        self.model = model
        self.x_in, self.t_in, self.y_in, self.out = io.x_in, io.t_in, io.y_in, io.out
        self.plan = ops.Plan()
        self.graph: Optional[th.cuda.CUDAGraph] = None
        self.launches = 0
        model.record_forward(self.plan, self.x_in, self.t_in, self.y_in, self.out, skip)

    def finalize(self, use_graph: bool, validate: bool = True):
        # first run outside capture: sets kernel attributes, validates the schedule. Later plans of the same model and
        # geometry launch the same kernels at the same shapes, so they may skip it (`validate=False`): a population of
        # fresh skip sets then costs no device time to plan (launch count = recorded ops until the first real run).
        self.launches = self.plan.run() if validate else self.plan.num_ops()
        self._want_graph = use_graph

    def capture(self):
        """Capture this forward in its own CUDA graph. Deferred to the first stand-alone `replay()`: a plan that is
        only ever consumed by `sampler.SchedulePlan` (which re-issues `plan.run()` into the schedule's graph) never
        pays for a graph of its own."""
        th.cuda.current_stream().synchronize()
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            self.plan.run()
        self.graph = g

    def replay(self):
        if self.graph is None and getattr(self, "_want_graph", False) and not th.cuda.is_current_stream_capturing():
            self.capture()
        if self.graph is not None:
            self.graph.replay()
        else:
            self.plan.run()


class Dynamic_UNetModel(nn.Module):
    """See module docstring. Constructor signature = dynamic_unet.py:447-468."""

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
        num_classes=None,
        use_checkpoint=False,
        use_fp16=False,
        num_heads=1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
    ):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only 2-D models are on the evaluator path")
        if not resblock_updown:
            raise NotImplementedError("resblock_updown=False (strided-conv resampling) is not used by any reference config")
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is not used by any reference config")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout  # identity at inference (dynamic_unet.py:218 under .eval())
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = th.float16 if use_fp16 else th.float32  # API compatibility; compute is bf16/fp32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.use_new_attention_order = use_new_attention_order

        time_embed_dim = model_channels * 4
        self.time_embed = _Seq({0: linear(model_channels, time_embed_dim), 2: linear(time_embed_dim, time_embed_dim)})
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)

        def res(cin, cout, lid, up=False, down=False):
            return ResBlock(cin, time_embed_dim, cout, use_scale_shift_norm, up=up, down=down, layer_id=lid)

        def attn(ch, heads, lid):
            return AttentionBlock(ch, num_heads=heads, num_head_channels=num_head_channels,
                                  use_new_attention_order=use_new_attention_order, layer_id=lid)

        # ---- same construction loops and layer-id numbering as dynamic_unet.py:500-655 ----
        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([_Seq({0: conv_nd(2, in_channels, ch, 3, padding=1)})])
        input_block_chans = [ch]
        ds = 1
        layer_id = 0
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, int(mult * model_channels), layer_id)]
                layer_id += 1
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads, layer_id))
                    layer_id += 1
                self.input_blocks.append(_Seq(dict(enumerate(layers))))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(_Seq({0: res(ch, ch, layer_id, down=True)}))
                layer_id += 1
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = _Seq({
            0: res(ch, ch, layer_id),
            1: attn(ch, num_heads, layer_id + 1),
            2: res(ch, ch, layer_id + 2),
        })
        layer_id += 3
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = input_block_chans.pop()
                layers = [res(ch + ich, int(model_channels * mult), layer_id)]
                layer_id += 1
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample, layer_id))
                    layer_id += 1
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, layer_id, up=True))
                    layer_id += 1
                    ds //= 2
                self.output_blocks.append(_Seq(dict(enumerate(layers))))
        self.out = _Seq({0: normalization(ch), 2: zero_module(conv_nd(2, input_ch, out_channels, 3, padding=1))})
        self.layer_num = layer_id

        # runtime caches (not part of state_dict)
        self._generation = 0
        self._packed_generation = -1
        self._packed: Dict[str, object] = {}
        # recorded forwards, least recently used first. Bounded: once the progressive search widens its prune range
        # nearly every candidate brings a new skip set, and each plan pins a launch list (+ a CUDA graph if it was
        # ever replayed stand-alone). The empty-mask plan - most steps of most candidates - is never evicted.
        self._plans: "OrderedDict[tuple, _UNetPlan]" = OrderedDict()
        self.max_cached_plans = int(os.environ.get("ADB_MAX_UNET_PLANS", "24"))
        self._io: Dict[tuple, _IO] = {}
        self._pool: Optional[_Pool] = None
        self.gpu_launches = 0  # kernels launched through this model (bench.py reports it)

    # ---- API parity helpers ----
    def convert_to_fp16(self):
        """dynamic_unet.py:657-663 casts the torso's conv weights to fp16. Here the master
        parameters stay fp32 and the tensor-core operands are packed to bf16 at plan build."""
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
        """Call after modifying parameters in place; re-packs operands and rebuilds plans."""
        self._invalidate()

    def _invalidate(self):
        if hasattr(self, "_generation"):
            self._generation += 1

    def io_buffers(self, B: int, H: int, W: int) -> _IO:
        """Static input/output buffers shared by all plans of this geometry."""
        key = (B, H, W, str(self._device()))
        io = self._io.get(key)
        if io is None:
            dev = self._device()
            io = _IO(th.zeros((B, self.in_channels, H, W), dtype=th.float32, device=dev),
                     th.zeros((B,), dtype=th.int64, device=dev),
                     th.zeros((B,), dtype=th.int64, device=dev) if self.num_classes is not None else None,
                     th.empty((B, self.out_channels, H, W), dtype=th.float32, device=dev))
            self._io[key] = io
        return io

    def _device(self):
        return self.out[0].weight.device

    # ---- weight packing ----
    def _iter_layers(self):
        for blk in list(self.input_blocks)[1:]:
            yield from blk.children()
        yield from self.middle_block.children()
        for blk in self.output_blocks:
            yield from blk.children()

    def _pack(self):
        dev = self._device()
        f32 = lambda p: p.detach().to(device=dev, dtype=th.float32).contiguous()
        P: Dict[str, object] = {}
        emb_w, emb_b, off = [], [], 0
        for layer in self._iter_layers():
            if isinstance(layer, ResBlock):
                c1 = getattr(layer.in_layers, "2")
                c2 = getattr(layer.out_layers, "3")
                n1 = getattr(layer.in_layers, "0")
                n2 = getattr(layer.out_layers, "0")
                el = getattr(layer.emb_layers, "1")
                has_skip = isinstance(layer.skip_connection, nn.Conv2d)
                if has_skip:
                    ws = layer.skip_connection.weight.detach()
                    bskip = f32(layer.skip_connection.bias)
                    b2 = (c2.bias.detach().float() + layer.skip_connection.bias.detach().float()).to(dev).contiguous()
                else:
                    ws = bskip = None
                    b2 = f32(c2.bias)
                P[id(layer)] = _PackedRes(
                    w1=ops.pack_conv_weight([c1.weight], dev), b1=f32(c1.bias),
                    w2_raw=c2.weight.detach(), b2=b2, ws_raw=ws, bskip=bskip,
                    g1=f32(n1.weight), be1=f32(n1.bias), g2=f32(n2.weight), be2=f32(n2.bias), ss_off=off,
                )
                emb_w.append(el.weight.detach().float())
                emb_b.append(el.bias.detach().float())
                off += el.weight.shape[0]
            elif isinstance(layer, AttentionBlock):
                P[id(layer)] = _PackedAttn(
                    g=f32(layer.norm.weight), be=f32(layer.norm.bias),
                    wqkv=ops.pack_conv_weight([layer.qkv.weight], dev), bqkv=f32(layer.qkv.bias),
                    wproj=ops.pack_conv_weight([layer.proj_out.weight], dev), bproj=f32(layer.proj_out.bias),
                )
        P["emb_w"] = ops.pack_linear_weight_split(th.cat(emb_w, 0), dev)
        P["emb_b"] = th.cat(emb_b, 0).to(dev).contiguous()
        P["emb_total"] = off
        te0, te2 = getattr(self.time_embed, "0"), getattr(self.time_embed, "2")
        P["te0_w"], P["te0_b"], P["te2_w"], P["te2_b"] = f32(te0.weight), f32(te0.bias), f32(te2.weight), f32(te2.bias)
        if self.num_classes is not None:
            P["label"] = f32(self.label_emb.weight)
        stem = getattr(self.input_blocks[0], "0")
        P["stem_w"], P["stem_b"] = f32(stem.weight), f32(stem.bias)
        P["stem_wp"] = ops.pack_stem_weight(stem.weight, dev) if self.in_channels == 3 else None
        on, oc = getattr(self.out, "0"), getattr(self.out, "2")
        P["out_g"], P["out_be"] = f32(on.weight), f32(on.bias)
        P["out_w"], P["out_b"] = ops.pack_conv_weight([oc.weight], dev), f32(oc.bias)
        P["w2_cache"] = {}
        self._packed = P
        self._packed_generation = self._generation
        self._plans.clear()

    def _w2_for(self, layer: ResBlock, split: Tuple[int, ...]) -> th.Tensor:
        """conv2 weights, with the 1x1 skip appended along K split at the concat boundary."""
        pk: _PackedRes = self._packed[id(layer)]
        key = (id(layer), split)
        cache = self._packed["w2_cache"]
        if key not in cache:
            ws = pk.ws_raw
            mats = [pk.w2_raw]
            if ws is not None:
                o = 0
                for c in split:
                    mats.append(ws[:, o:o + c])
                    o += c
                assert o == ws.shape[1]
            cache[key] = ops.pack_conv_weight(mats, self._device())
        return cache[key]

    # ---- recording: the reference's forward walked once, skipped blocks elided ----
    def record_forward(self, plan: ops.Plan, x_in: th.Tensor, t_in: th.Tensor, y_in: Optional[th.Tensor],
                       out: th.Tensor, skip_layer: Sequence[int] = ()):
        """Record one forward (x_in fp32 NCHW, t_in int64 [B], y_in int64 [B] or None -> out fp32 NCHW)
        into `plan`. Several forwards — e.g. every step of a searched schedule, each with its own
        skip set — may be recorded into the same plan; they share this model's activation pool."""
        if self._device().type != "cuda":
            raise RuntimeError("Dynamic_UNetModel runs on a CUDA device only: move it with .to('cuda') "
                               "(autodiffusion_b200 has no CPU path)")
        if self._packed_generation != self._generation:
            self._pack()
        B, _, H, W = x_in.shape
        skip = set(int(s) for s in skip_layer)
        if self._pool is None or self._pool.device != self._device():
            self._pool = _Pool(self._device())
        P = self._packed
        up = _IO(x_in, t_in, y_in, out)
        ctx = _Ctx(self._pool, plan)
        dev = self._device()
        mc, ted = self.model_channels, self.model_channels * 4
        # Per-forward scratch (statistics arena, embedding vectors) comes from the model's shared pool like the
        # activations and goes back to it at the end of the forward: a recorded plan pins no device memory of its own,
        # so the number of cached / live plans does not grow the footprint (plans run one after another on a stream).
        scratch: List[th.Tensor] = []

        def salloc(shape, dtype):
            t = ctx.alloc(shape, dtype)
            scratch.append(t)
            return t

        plan.keep(up.x_in, up.t_in, up.y_in, up.out)
        # GroupNorm sums accumulated by the PRODUCING conv's epilogue (one slot per conv output that a
        # single-source GroupNorm will read); the whole arena is zeroed by one memset per forward.
        fuse_stats = os.environ.get("ADB_NO_FUSED_STATS", "0") != "1" and (H * W) % 32 == 0
        # which conv outputs are later read as one half of a channel concat by a GroupNorm: decided by a
        # symbolic pre-walk, so that producer can also accumulate sums in the concat's group layout
        cat_uses, n_cat = self._plan_concat_stats(H, W, skip, fuse_stats)
        n_slots = 2 * self.layer_num + 4
        stats = salloc((B, 32, 2), th.float64)  # scratch of the stand-alone stats pass
        arena = salloc((n_slots + n_cat, B, 32, 2), th.float64)
        slot = [0]
        prod_idx = [0]
        produced: Dict[int, th.Tensor] = {}  # data_ptr of an activation -> its producer-filled stats
        produced_tag: Dict[int, int] = {}    # data_ptr -> production index (for concat lookups)

        def _forget(t):
            produced.pop(t.data_ptr(), None)
            produced_tag.pop(t.data_ptr(), None)

        ctx.on_alloc = _forget  # a recycled buffer has no stats yet
        if fuse_stats:
            ops.memset0(arena, plan=plan)

        def new_stats(t: th.Tensor):
            """-> kwargs for conv_igemm: stats slot(s) for conv output `t` ({} when fusion is off or the
            geometry does not allow it). Must be called in the same order as _plan_concat_stats counts."""
            if not fuse_stats or t.shape[3] % 32 != 0 or (t.shape[1] * t.shape[2]) % 32 != 0:
                _forget(t)
                return {}
            st = arena[slot[0]]
            slot[0] += 1
            p = prod_idx[0]
            prod_idx[0] += 1
            produced[t.data_ptr()] = st
            produced_tag[t.data_ptr()] = p
            kw = {"stats_out": st}
            use = cat_uses.get(p)
            if use is not None:
                cslot, choff, cpg = use
                kw["stats2"] = (arena[n_slots + cslot], cpg, choff)
            return kw

        def gn(srcs, gamma, beta, out_t, **kw):
            """GroupNorm reading producer-filled sums when its source(s) have them."""
            st = None
            if len(srcs) == 1:
                st = produced.get(srcs[0].data_ptr())
            else:
                p0, p1 = produced_tag.get(srcs[0].data_ptr()), produced_tag.get(srcs[1].data_ptr())
                u0, u1 = cat_uses.get(p0), cat_uses.get(p1)
                if u0 is not None and u1 is not None and u0[0] == u1[0] and u0[1] == 0 and u1[1] == srcs[0].shape[3]:
                    st = arena[n_slots + u0[0]]
            ops.groupnorm(srcs[0], gamma, beta, src1=srcs[1] if len(srcs) > 1 else None, out=out_t,
                          stats=st if st is not None else stats, stats_ready=st is not None, plan=plan, **kw)

        # timestep / label embedding (dynamic_unet.py:687-691) and every emb_layers product (:259)
        te = ops.timestep_embedding(up.t_in, mc, out=salloc((B, mc), th.float32), plan=plan)
        e1 = ops.linear(te, P["te0_w"], P["te0_b"], out=salloc((B, ted), th.float32), plan=plan)
        emb = ops.linear(e1, P["te2_w"], P["te2_b"], silu_in=True, table=P.get("label"), idx=up.y_in,
                         out=salloc((B, ted), th.float32), plan=plan)
        # every emb_layers Linear of the forward as ONE tensor-core product (fp32-grade: split-bf16 operands)
        ss_total = P["emb_total"]
        ss_all = ops.linear_tc(emb, P["emb_w"], P["emb_b"], ss_total, silu_in=True, plan=plan,
                               out=salloc((B, ss_total), th.float32),
                               hi_lo=(salloc((B, 1, 1, ted), th.bfloat16), salloc((B, 1, 1, ted), th.bfloat16)))

        def run_res(layer: ResBlock, srcs: List[th.Tensor]) -> th.Tensor:
            pk: _PackedRes = P[id(layer)]
            n, h, w = srcs[0].shape[:3]
            cout = layer.out_channels
            updown = layer.up or layer.down
            if layer.layer_id in skip:  # dynamic_unet.py:246-249
                if updown:
                    o = ctx.alloc((n, h * 2, w * 2, cout) if layer.up else (n, h // 2, w // 2, cout))
                    _forget(o)
                    return ops.resample2x(srcs[0], ops.RESAMPLE_NEAREST2 if layer.up else ops.RESAMPLE_AVGPOOL2,
                                          out=o, plan=plan)
                if pk.ws_raw is None:
                    ctx.retain(srcs[0])
                    return srcs[0]
                wsk = self._w2_skip_only(layer, tuple(s.shape[3] for s in srcs))
                o = ctx.alloc((n, h, w, cout))
                return ops.conv_igemm([(s, 1) for s in srcs], wsk, pk.bskip, cout, out=o, plan=plan, **new_stats(o))
            ho, wo = (h * 2, w * 2) if layer.up else ((h // 2, w // 2) if layer.down else (h, w))
            mode = ops.RESAMPLE_NEAREST2 if layer.up else (ops.RESAMPLE_AVGPOOL2 if layer.down else ops.RESAMPLE_NONE)
            cin = sum(s.shape[3] for s in srcs)
            g1 = ctx.alloc((n, ho, wo, cin))
            gn(srcs, pk.g1, pk.be1, g1, silu=True, resample=mode)
            c1 = ctx.alloc((n, ho, wo, cout))
            ops.conv_igemm([(g1, 9)], pk.w1, pk.b1, cout, out=c1, plan=plan, **new_stats(c1))
            ctx.release(g1)
            g2 = ctx.alloc((n, ho, wo, cout))
            gn([c1], pk.g2, pk.be2, g2, scale_shift=(ss_all, pk.ss_off), ss_stride=ss_total, silu=True)
            ctx.release(c1)
            out = ctx.alloc((n, ho, wo, cout))
            if pk.ws_raw is not None:
                w2 = self._w2_for(layer, tuple(s.shape[3] for s in srcs))
                ops.conv_igemm([(g2, 9)] + [(s, 1) for s in srcs], w2, pk.b2, cout, out=out, plan=plan,
                               **new_stats(out))
            else:
                w2 = self._w2_for(layer, ())
                rm = ops.RES_NEAREST2 if layer.up else (ops.RES_AVGPOOL2 if layer.down else ops.RES_SAME)
                ops.conv_igemm([(g2, 9)], w2, pk.b2, cout, out=out, residual=srcs[0], res_mode=rm, plan=plan,
                               **new_stats(out))
            ctx.release(g2)
            return out

        def run_attn(layer: AttentionBlock, x: th.Tensor) -> th.Tensor:
            if layer.layer_id in skip:  # dynamic_unet.py:317-318
                ctx.retain(x)
                return x
            pk: _PackedAttn = P[id(layer)]
            n, h, w, c = x.shape
            t = h * w
            g = ctx.alloc((n, h, w, c))
            gn([x], pk.g, pk.be, g, silu=False)
            qkv = ctx.alloc((n, h, w, 3 * c))
            ops.conv_igemm([(g, 1)], pk.wqkv, pk.bqkv, 3 * c, out=qkv, plan=plan)
            ctx.release(g)
            a = ctx.alloc((n, h, w, c))
            ops.attention(qkv.view(n * t, 3 * c), n, t, layer.num_heads, not layer.use_new_attention_order,
                          out=a.view(n * t, c), plan=plan)
            ctx.release(qkv)
            out = ctx.alloc((n, h, w, c))
            ops.conv_igemm([(a, 1)], pk.wproj, pk.bproj, c, out=out, residual=x, res_mode=ops.RES_SAME, plan=plan,
                           **new_stats(out))
            ctx.release(a)
            return out

        def run_block(blk: _Seq, srcs: List[th.Tensor]) -> th.Tensor:
            """Consumes one reference to each tensor in srcs; returns a tensor the caller owns."""
            for layer in blk.children():
                out = run_res(layer, srcs) if isinstance(layer, ResBlock) else run_attn(layer, srcs[0])
                for s in srcs:
                    ctx.release(s)
                srcs = [out]
            return srcs[0]

        ch0 = int(self.channel_mult[0] * mc)
        h = ctx.alloc((B, H, W, ch0))
        if P["stem_wp"] is not None:  # tensor-core stem; its epilogue also fills the first GroupNorm's sums
            ops.stem_conv_tc(up.x_in, P["stem_wp"], P["stem_b"], ch0, out=h, plan=plan, **new_stats(h))
        else:
            _forget(h)
            ops.stem_conv(up.x_in, P["stem_w"], P["stem_b"], out=h, plan=plan)
        hs = [h]
        ctx.retain(h)  # one reference for hs, one for the running h
        for blk in list(self.input_blocks)[1:]:
            h = run_block(blk, [h])
            hs.append(h)
            ctx.retain(h)
        h = run_block(self.middle_block, [h])
        for blk in self.output_blocks:
            h = run_block(blk, [h, hs.pop()])  # th.cat([h, hs.pop()], dim=1), dynamic_unet.py:699
        g = ctx.alloc(tuple(h.shape))
        gn([h], P["out_g"], P["out_be"], g, silu=True)
        ctx.release(h)
        ops.conv_igemm([(g, 9)], P["out_w"], P["out_b"], self.out_channels, out=up.out,
                       out_mode=ops.OUT_F32_NCHW, plan=plan)
        ctx.release(g)
        for t in scratch:
            ctx.release(t)

    def _plan_concat_stats(self, H: int, W: int, skip: set, fuse_stats: bool):
        """Symbolic twin of the walk in record_forward: numbers every stats-producing conv output in
        production order and returns {production index: (concat slot, channel offset, channels per group)}
        for those later consumed as one half of `th.cat([h, hs.pop()], 1)` by a (non-skipped) ResBlock's
        first GroupNorm, plus the number of concat slots."""
        uses: Dict[int, Tuple[int, int, int]] = {}
        if not fuse_stats:
            return uses, 0
        counter = [0]
        n_cat = [0]

        class T:  # symbolic activation
            __slots__ = ("tag", "c", "h", "w")

            def __init__(self, tag, c, h, w):
                self.tag, self.c, self.h, self.w = tag, c, h, w

        def produce(c, h, w):
            if c % 32 != 0 or (h * w) % 32 != 0:
                return T(None, c, h, w)
            t = T(counter[0], c, h, w)
            counter[0] += 1
            return t

        def res(layer, srcs):
            s0 = srcs[0]
            cout = layer.out_channels
            if layer.layer_id in skip:
                if layer.up or layer.down:
                    return T(None, cout, s0.h * 2 if layer.up else s0.h // 2, s0.w * 2 if layer.up else s0.w // 2)
                if not isinstance(layer.skip_connection, nn.Conv2d):
                    return s0
                return produce(cout, s0.h, s0.w)
            ho, wo = (s0.h * 2, s0.w * 2) if layer.up else ((s0.h // 2, s0.w // 2) if layer.down else (s0.h, s0.w))
            if len(srcs) == 2 and srcs[0].tag is not None and srcs[1].tag is not None:
                cpg = (srcs[0].c + srcs[1].c) // 32
                if 96 // cpg + 2 <= 40:
                    k = n_cat[0]
                    n_cat[0] += 1
                    uses[srcs[0].tag] = (k, 0, cpg)
                    uses[srcs[1].tag] = (k, srcs[0].c, cpg)
            produce(cout, ho, wo)          # conv1 output
            return produce(cout, ho, wo)   # block output

        def attn(layer, x):
            return x if layer.layer_id in skip else produce(x.c, x.h, x.w)

        def block(blk, srcs):
            for layer in blk.children():
                srcs = [res(layer, srcs) if isinstance(layer, ResBlock) else attn(layer, srcs[0])]
            return srcs[0]

        ch0 = int(self.channel_mult[0] * self.model_channels)
        h = produce(ch0, H, W) if self.in_channels == 3 else T(None, ch0, H, W)  # stem: same rule as record_forward
        hs = [h]
        for blk in list(self.input_blocks)[1:]:
            h = block(blk, [h])
            hs.append(h)
        h = block(self.middle_block, [h])
        for blk in self.output_blocks:
            h = block(blk, [h, hs.pop()])
        return uses, n_cat[0]

    # ---- analytic cost model (scheduling of population evaluation: longest candidates first) ----
    def layer_flops(self, H: Optional[int] = None, W: Optional[int] = None):
        """-> (base, {layer_id: (run, skipped)}) in FLOP (2 x MAC) per image per forward: what block `layer_id` costs when
        it runs and what is left of it when it is skipped (its 1x1 skip_connection, dynamic_unet.py:246-249). `base` is the
        unskippable remainder (stem, out conv, time embedding). Reproduces SURVEY.md table A3' (hook-measured on the
        reference module) to the digits printed there."""
        H = H or self.image_size
        W = W or self.image_size
        mc, ted = self.model_channels, self.model_channels * 4
        costs: Dict[int, Tuple[float, float]] = {}

        def res(layer: ResBlock, cin, h, w):
            cout = layer.out_channels
            ho, wo = (h * 2, w * 2) if layer.up else ((h // 2, w // 2) if layer.down else (h, w))
            px = ho * wo
            skipc = 2.0 * cin * cout * px if isinstance(layer.skip_connection, nn.Conv2d) else 0.0
            run = 2.0 * 9 * cin * cout * px + 2.0 * 9 * cout * cout * px + skipc + 2.0 * ted * 2 * cout
            costs[layer.layer_id] = (run, skipc)
            return cout, ho, wo

        def attn(layer: AttentionBlock, c, h, w):
            t = h * w
            costs[layer.layer_id] = (2.0 * c * 3 * c * t + 2.0 * c * c * t + 4.0 * t * t * c, 0.0)

        def block(blk, c, h, w):
            for layer in blk.children():
                if isinstance(layer, ResBlock):
                    c, h, w = res(layer, layer.channels, h, w)
                else:
                    attn(layer, c, h, w)
            return c, h, w

        ch0 = int(self.channel_mult[0] * mc)
        c, h, w = ch0, H, W
        for blk in list(self.input_blocks)[1:]:
            c, h, w = block(blk, c, h, w)
        c, h, w = block(self.middle_block, c, h, w)
        for blk in self.output_blocks:
            c, h, w = block(blk, c, h, w)
        base = 2.0 * 9 * self.in_channels * ch0 * H * W + 2.0 * 9 * ch0 * self.out_channels * H * W \
            + 2.0 * mc * ted + 2.0 * ted * ted
        return base, costs

    def forward_flops(self, skip_layer: Sequence[int] = (), H: Optional[int] = None, W: Optional[int] = None) -> float:
        """Algorithmic FLOP per image of one forward with `skip_layer` elided (F(cand) is the sum of
        this over the candidate's steps)."""
        key = (H or self.image_size, W or self.image_size)
        cache = self.__dict__.setdefault("_flops_cache", {})
        if key not in cache:
            cache[key] = self.layer_flops(*key)
        base, costs = cache[key]
        skip = set(int(s) for s in skip_layer)
        return base + sum(sk if lid in skip else run for lid, (run, sk) in costs.items())

    def _w2_skip_only(self, layer: ResBlock, split: Tuple[int, ...]) -> th.Tensor:
        key = (id(layer), "skip", split)
        cache = self._packed["w2_cache"]
        if key not in cache:
            ws = self._packed[id(layer)].ws_raw
            mats, o = [], 0
            for c in split:
                mats.append(ws[:, o:o + c])
                o += c
            cache[key] = ops.pack_conv_weight(mats, self._device())
        return cache[key]

    # ---- public forward ----
    def get_plan(self, B: int, H: int, W: int, skip_layer: Sequence[int] = (), validate: Optional[bool] = None) -> _UNetPlan:
        if self._device().type != "cuda":
            raise RuntimeError("Dynamic_UNetModel runs on a CUDA device only: move it with .to('cuda') "
                               "(autodiffusion_b200 has no CPU path)")
        if self._packed_generation != self._generation:
            self._pack()
        key = (B, H, W, tuple(sorted(set(int(s) for s in skip_layer))))
        up = self._plans.get(key)
        if up is None:
            with th.no_grad():
                up = _UNetPlan(self, B, H, W, key[3])
                if validate is None or validate:
                    validate = True
                else:  # the caller allows skipping the validation run: only once this geometry ran for real
                    validate = not any(k[:3] == key[:3] and getattr(v, "_validated", False) for k, v in self._plans.items())
                up.finalize(use_graph=os.environ.get("ADB_NO_GRAPH", "0") != "1", validate=validate)
                up._validated = validate
            self._plans[key] = up
            while len(self._plans) > max(2, self.max_cached_plans):
                victim = next((k for k in self._plans if k[3] != () and k != key), None)  # oldest first
                if victim is None:
                    break
                del self._plans[victim]  # a SchedulePlan still holding it keeps it alive; the cache lets go
        else:
            self._plans.move_to_end(key)
        return up

    def forward(self, x, timesteps, y=None, skip_layer=[]):
        """dynamic_unet.py:673-702. x fp32 [N,C,H,W]; timesteps [N]; y int64 [N] iff class-conditional."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("Dynamic_UNetModel.forward: input must be a CUDA tensor (no CPU path)")
        B, _, H, W = x.shape
        assert timesteps.shape == (B,)
        if y is not None:
            assert y.shape == (B,)
        tr = self.__dict__.get("_trace")
        if tr is not None:  # fastpath.py: record the call, run nothing, hand back the sentinel
            tr.append((x, timesteps, y, list(skip_layer)))
            return self.io_buffers(B, H, W).out
        up = self.get_plan(B, H, W, skip_layer)
        up.x_in.copy_(x)
        up.t_in.copy_(timesteps)  # integer timesteps (rescale_timesteps=False in every reference config)
        if y is not None:
            up.y_in.copy_(y)
        up.replay()
        self.gpu_launches += up.launches
        return up.out.clone()


class UNetModel(Dynamic_UNetModel):
    """guided_diffusion/unet.py UNetModel (:396-665): same network, no skip argument."""

    def forward(self, x, timesteps, y=None):
        return super().forward(x, timesteps, y, skip_layer=[])
