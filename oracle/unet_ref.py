"""Functional fp32 restatement of the reference UNets (test oracle; see oracle/__init__.py).

Follows, module by module:
  * Dynamic_UNetModel.__init__/forward   guided_diffusion/dynamic_unet.py:447-655, 673-702
  * ResBlock._forward (skip branch)      guided_diffusion/dynamic_unet.py:245-271
  * AttentionBlock._forward              guided_diffusion/dynamic_unet.py:316-325
  * QKVAttentionLegacy / QKVAttention    guided_diffusion/dynamic_unet.py:357-374 / 390-409
  * Upsample / Downsample (no conv)      guided_diffusion/dynamic_unet.py:107-117, 145-147
  * GroupNorm32, timestep_embedding      guided_diffusion/nn.py:17-19, 103-121
  * EncoderUNetModel (+AttentionPool2d)  guided_diffusion/unet.py:685-896, 22-51
  * create_model / create_classifier     guided_diffusion/script_util.py:133-211, 257-295

Weights are a plain dict keyed exactly like the reference modules' `state_dict()`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

NUM_CLASSES = 1000  # script_util.py:9


@dataclass
class UNetConfig:
    image_size: int = 64
    in_channels: int = 3
    model_channels: int = 192
    out_channels: int = 6
    num_res_blocks: int = 3
    attention_resolutions: Tuple[int, ...] = (2, 4, 8)  # downsample rates (image_size // res)
    channel_mult: Tuple[int, ...] = (1, 2, 3, 4)
    num_classes: Optional[int] = NUM_CLASSES
    num_heads: int = 4
    num_head_channels: int = 64
    num_heads_upsample: int = -1
    use_scale_shift_norm: bool = True
    resblock_updown: bool = True
    use_new_attention_order: bool = True
    # encoder (classifier) only
    pool: str = "attention"


def adm_g64_config() -> UNetConfig:
    """Flags of search_dynamic_unet_imagenet64_classifier_guidance_progressive.sh:1."""
    return UNetConfig()


def classifier64_config(depth: int = 4, width: int = 128) -> UNetConfig:
    """create_classifier (script_util.py:257-295) with classifier_defaults (:27-40)."""
    return UNetConfig(
        image_size=64, in_channels=3, model_channels=width, out_channels=1000, num_res_blocks=depth,
        attention_resolutions=(2, 4, 8), channel_mult=(1, 2, 3, 4), num_classes=None, num_heads=1,
        num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True,
        use_new_attention_order=False, pool="attention",
    )


# --------------------------------------------------------------------------------------
# architecture walk: the same loops as the reference constructors, yielding block records
# --------------------------------------------------------------------------------------
@dataclass
class ResSpec:
    name: str
    cin: int
    cout: int
    up: bool = False
    down: bool = False
    layer_id: int = -1


@dataclass
class AttnSpec:
    name: str
    channels: int
    heads: int
    layer_id: int = -1


@dataclass
class Arch:
    input_blocks: List[list] = field(default_factory=list)   # each a list of ResSpec/AttnSpec ('stem' first)
    middle: list = field(default_factory=list)
    output_blocks: List[list] = field(default_factory=list)
    layer_num: int = 0
    final_ch: int = 0
    time_embed_dim: int = 0


def _heads(cfg: UNetConfig, ch: int, upsample: bool) -> int:
    if cfg.num_head_channels != -1:
        return ch // cfg.num_head_channels
    nhu = cfg.num_heads if cfg.num_heads_upsample == -1 else cfg.num_heads_upsample
    return nhu if upsample else cfg.num_heads


def build_arch(cfg: UNetConfig, encoder_only: bool = False) -> Arch:
    """dynamic_unet.py:500-655 (layer ids :507-655); encoder: unet.py EncoderUNetModel.__init__."""
    assert cfg.resblock_updown, "oracle covers resblock_updown=True (all reference configs in scope)"
    a = Arch(time_embed_dim=cfg.model_channels * 4)
    mc = cfg.model_channels
    ch = int(cfg.channel_mult[0] * mc)
    a.input_blocks.append(["stem"])
    chans = [ch]
    ds = 1
    lid = 0
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            idx = len(a.input_blocks)
            layers = [ResSpec(f"input_blocks.{idx}.0", ch, int(mult * mc), layer_id=lid)]
            lid += 1
            ch = int(mult * mc)
            if ds in cfg.attention_resolutions:
                layers.append(AttnSpec(f"input_blocks.{idx}.1", ch, _heads(cfg, ch, False), layer_id=lid))
                lid += 1
            a.input_blocks.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            idx = len(a.input_blocks)
            a.input_blocks.append([ResSpec(f"input_blocks.{idx}.0", ch, ch, down=True, layer_id=lid)])
            lid += 1
            chans.append(ch)
            ds *= 2
    a.middle = [
        ResSpec("middle_block.0", ch, ch, layer_id=lid),
        AttnSpec("middle_block.1", ch, _heads(cfg, ch, False), layer_id=lid + 1),
        ResSpec("middle_block.2", ch, ch, layer_id=lid + 2),
    ]
    lid += 3
    if encoder_only:
        a.layer_num = lid
        a.final_ch = ch
        return a
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            idx = len(a.output_blocks)
            layers = [ResSpec(f"output_blocks.{idx}.0", ch + ich, int(mc * mult), layer_id=lid)]
            lid += 1
            ch = int(mc * mult)
            if ds in cfg.attention_resolutions:
                layers.append(AttnSpec(f"output_blocks.{idx}.{len(layers)}", ch, _heads(cfg, ch, True), layer_id=lid))
                lid += 1
            if level and i == cfg.num_res_blocks:
                layers.append(ResSpec(f"output_blocks.{idx}.{len(layers)}", ch, ch, up=True, layer_id=lid))
                lid += 1
                ds //= 2
            a.output_blocks.append(layers)
    a.layer_num = lid
    a.final_ch = ch
    return a


def param_shapes(cfg: UNetConfig, encoder_only: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Every state_dict key and shape of the reference module, in construction order."""
    a = build_arch(cfg, encoder_only)
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, i, o):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    def conv(name, i, o, k, dims=2):
        s[name + ".weight"] = (o, i) + (k,) * dims
        s[name + ".bias"] = (o,)

    def gn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    def res(r: ResSpec):
        gn(r.name + ".in_layers.0", r.cin)
        conv(r.name + ".in_layers.2", r.cin, r.cout, 3)
        lin(r.name + ".emb_layers.1", ted, 2 * r.cout if cfg.use_scale_shift_norm else r.cout)
        gn(r.name + ".out_layers.0", r.cout)
        conv(r.name + ".out_layers.3", r.cout, r.cout, 3)
        if r.cin != r.cout:
            conv(r.name + ".skip_connection", r.cin, r.cout, 1)

    def attn(t: AttnSpec):
        gn(t.name + ".norm", t.channels)
        conv(t.name + ".qkv", t.channels, 3 * t.channels, 1, dims=1)
        conv(t.name + ".proj_out", t.channels, t.channels, 1, dims=1)

    def block(layers):
        for l in layers:
            if isinstance(l, ResSpec):
                res(l)
            elif isinstance(l, AttnSpec):
                attn(l)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)
    if cfg.num_classes is not None and not encoder_only:
        s["label_emb.weight"] = (cfg.num_classes, ted)
    conv("input_blocks.0.0", cfg.in_channels, int(cfg.channel_mult[0] * mc), 3)
    for layers in a.input_blocks[1:]:
        block(layers)
    block(a.middle)
    if encoder_only:
        assert cfg.pool == "attention"
        gn("out.0", a.final_ch)
        spatial = cfg.image_size // (2 ** (len(cfg.channel_mult) - 1))
        s["out.2.positional_embedding"] = (a.final_ch, spatial ** 2 + 1)
        conv("out.2.qkv_proj", a.final_ch, 3 * a.final_ch, 1, dims=1)
        conv("out.2.c_proj", a.final_ch, cfg.out_channels, 1, dims=1)
        return s
    for layers in a.output_blocks:
        block(layers)
    gn("out.0", a.final_ch)
    conv("out.2", int(cfg.channel_mult[0] * mc), cfg.out_channels, 3)
    return s


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm32(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.py:17-19 (GroupNorm(32, C), eps 1e-5, computed in fp32)."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5).type(x.dtype)


def qkv_attention(qkv: torch.Tensor, n_heads: int, new_order: bool) -> torch.Tensor:
    """dynamic_unet.py:390-409 (new order) / 357-374 (legacy). qkv: [N, 3*H*C, T]."""
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    scale = 1 / math.sqrt(math.sqrt(ch))
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q = (q * scale).reshape(bs * n_heads, ch, length)
        k = (k * scale).reshape(bs * n_heads, ch, length)
        v = v.reshape(bs * n_heads, ch, length)
    else:
        q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
        q, k = q * scale, k * scale
    weight = torch.einsum("bct,bcs->bts", q, k)
    weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
    a = torch.einsum("bts,bcs->bct", weight, v)
    return a.reshape(bs, -1, length)


def _x_upd(x: torch.Tensor, r: ResSpec) -> torch.Tensor:
    if r.up:
        return F.interpolate(x, scale_factor=2, mode="nearest")  # dynamic_unet.py:114
    if r.down:
        return F.avg_pool2d(x, 2, 2)  # dynamic_unet.py:143
    return x


def res_block(sd, r: ResSpec, x: torch.Tensor, emb: torch.Tensor, skip_layer: Sequence[int],
              scale_shift: bool = True) -> torch.Tensor:
    """ResBlock._forward, dynamic_unet.py:245-271 (eval mode: dropout is identity)."""
    p = r.name
    has_skip_conv = (p + ".skip_connection.weight") in sd

    def skip_connection(t):
        if has_skip_conv:
            return F.conv2d(t, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
        return t

    if r.layer_id in skip_layer:  # :246-249
        return skip_connection(_x_upd(x, r))
    h = F.silu(group_norm32(x, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"]))
    if r.up or r.down:  # :251-256
        h = _x_upd(h, r)
        x = _x_upd(x, r)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"]).type(h.dtype)
    emb_out = emb_out[..., None, None]
    hn = group_norm32(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"])
    if scale_shift:  # :262-266
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = hn * (1 + scale) + shift
        h = F.silu(h)
    else:  # :268-269
        h = F.silu(group_norm32(h + emb_out, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"]))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    return skip_connection(x) + h


def attention_block(sd, t: AttnSpec, x: torch.Tensor, skip_layer: Sequence[int], new_order: bool) -> torch.Tensor:
    """AttentionBlock._forward, dynamic_unet.py:316-325."""
    if t.layer_id in skip_layer:
        return x
    p = t.name
    b, c, *spatial = x.shape
    xr = x.reshape(b, c, -1)
    qkv = F.conv1d(group_norm32(xr, sd[p + ".norm.weight"], sd[p + ".norm.bias"]), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    h = qkv_attention(qkv, t.heads, new_order)
    h = F.conv1d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xr + h).reshape(b, c, *spatial)


def _run_block(sd, cfg, layers, h, emb, skip_layer):
    for l in layers:
        if isinstance(l, ResSpec):
            h = res_block(sd, l, h, emb, skip_layer, cfg.use_scale_shift_norm)
        else:
            h = attention_block(sd, l, h, skip_layer, cfg.use_new_attention_order)
    return h


def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, x: torch.Tensor, timesteps: torch.Tensor,
                 y: Optional[torch.Tensor] = None, skip_layer: Sequence[int] = ()) -> torch.Tensor:
    """Dynamic_UNetModel.forward, dynamic_unet.py:673-702."""
    assert (y is not None) == (cfg.num_classes is not None), \
        "must specify y if and only if the model is class-conditional"
    a = build_arch(cfg)
    emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg.num_classes is not None:
        assert y.shape == (x.shape[0],)
        emb = emb + sd["label_emb.weight"][y]
    hs = []
    h = F.conv2d(x, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
    hs.append(h)
    for layers in a.input_blocks[1:]:
        h = _run_block(sd, cfg, layers, h, emb, skip_layer)
        hs.append(h)
    h = _run_block(sd, cfg, a.middle, h, emb, skip_layer)
    for layers in a.output_blocks:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, cfg, layers, h, emb, skip_layer)
    h = F.silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def encoder_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, x: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """EncoderUNetModel.forward (unet.py, pool='attention') + AttentionPool2d.forward (unet.py:43-51)."""
    a = build_arch(cfg, encoder_only=True)
    emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    h = F.conv2d(x, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
    for layers in a.input_blocks[1:]:
        h = _run_block(sd, cfg, layers, h, emb, ())
    h = _run_block(sd, cfg, a.middle, h, emb, ())
    h = F.silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"]))
    b, c = h.shape[:2]
    h = h.reshape(b, c, -1)
    h = torch.cat([h.mean(dim=-1, keepdim=True), h], dim=-1)
    h = h + sd["out.2.positional_embedding"][None, :, :].to(h.dtype)
    h = F.conv1d(h, sd["out.2.qkv_proj.weight"], sd["out.2.qkv_proj.bias"])
    heads = c // cfg.num_head_channels
    h = qkv_attention(h, heads, new_order=True)  # AttentionPool2d always uses QKVAttention (unet.py:41)
    h = F.conv1d(h, sd["out.2.c_proj.weight"], sd["out.2.c_proj.bias"])
    return h[:, :, 0]


def classifier_cond_fn(sd, cfg: UNetConfig, classifier_scale: float):
    """The search script's cond_fn closure (…progressive.py:383-390)."""

    def cond_fn(x, t, y=None, **_):
        assert y is not None
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = encoder_forward(sd, cfg, x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return torch.autograd.grad(selected.sum(), x_in)[0] * classifier_scale

    return cond_fn
