"""Functional fp32 restatement of the Stable-Diffusion-v1 UNet and its searched-timestep DDIM sampler
(test oracle; see oracle/__init__.py — never imported by the product path).

Paths below are relative to /root/reference/examples/Stable Diffusion/ (`SD/`). Follows:
  * UNetModel.__init__ / forward         ldm/modules/diffusionmodules/openaimodel.py:413-708, 710-742
  * ResBlock._forward (no scale-shift)   ldm/modules/diffusionmodules/openaimodel.py:255-275
  * Upsample / Downsample (with conv)    ldm/modules/diffusionmodules/openaimodel.py:109-118, 158-160
  * SpatialTransformer.forward           ldm/modules/attention.py:245-260 (Normalize = GroupNorm eps 1e-6, :75-76)
  * BasicTransformerBlock._forward       ldm/modules/attention.py:212-216
  * CrossAttention.forward               ldm/modules/attention.py:170-194
  * GEGLU / FeedForward                  ldm/modules/attention.py:37-64
  * timestep_embedding, GroupNorm32      ldm/modules/diffusionmodules/util.py:152-172, 215-217
  * make_beta_schedule("linear")         ldm/modules/diffusionmodules/util.py:21-43
  * make_ddim_sampling_parameters        ldm/modules/diffusionmodules/util.py:63-75
  * DDIMSampler.sample / ddim_sampling / p_sample_ddim with `sampled_timestep`
                                         ldm/models/diffusion/ddim.py:59-119, 121-175, 177-217
  * PLMSSampler.sample / plms_sampling / p_sample_plms with `sampled_timestep`
                                         ldm/models/diffusion/plms.py:62-122, 124-188, 190-257
  * DPMSolverSampler.sample -> DPM_Solver(predict_x0=True).sample(method="multistep", order=2, lower_order_final=True,
    ea_timesteps=cand) with NoiseScheduleVP("discrete") and the classifier-free model_wrapper
                                         ldm/models/diffusion/dpm_solver/sampler.py:20-83,
                                         ldm/models/diffusion/dpm_solver/dpm_solver.py:97-156, 278-343, 386-399,
                                         504-533, 755-790, 1072-1121, 1149-1188
  * the candidate call (CFG 7.5)         scripts/search_ea.py:737-739

Weights are a plain dict keyed exactly like the reference module's `state_dict()`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class SDConfig:
    """unet_config.params of configs/stable-diffusion/v1-inference_coco.yaml:29-44."""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_heads: int = 8
    transformer_depth: int = 1
    context_dim: int = 768


def sd_v1_config() -> SDConfig:
    return SDConfig()


@dataclass
class Block:
    kind: str          # "conv_in" | "res" | "st" | "down" | "up"
    name: str          # state_dict prefix
    cin: int = 0
    cout: int = 0
    heads: int = 0
    d_head: int = 0


@dataclass
class SDArch:
    input_blocks: List[List[Block]]
    middle: List[Block]
    output_blocks: List[List[Block]]
    skip_chans: List[int]   # channels of hs[i] (input block outputs), in push order
    final_ch: int


def build_arch(cfg: SDConfig) -> SDArch:
    """The constructor loops of openaimodel.py:485-690 with use_spatial_transformer=True, legacy=False,
    num_head_channels=-1 (dim_head = ch // num_heads), resblock_updown=False, conv_resample=True."""
    mc = cfg.model_channels
    inp: List[List[Block]] = [[Block("conv_in", "input_blocks.0.0", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            n = len(inp)
            layers = [Block("res", f"input_blocks.{n}.0", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(Block("st", f"input_blocks.{n}.1", ch, ch, cfg.num_heads, ch // cfg.num_heads))
            inp.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            n = len(inp)
            inp.append([Block("down", f"input_blocks.{n}.0", ch, ch)])
            chans.append(ch)
            ds *= 2
    middle = [Block("res", "middle_block.0", ch, ch),
              Block("st", "middle_block.1", ch, ch, cfg.num_heads, ch // cfg.num_heads),
              Block("res", "middle_block.2", ch, ch)]
    out: List[List[Block]] = []
    stack = list(chans)
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = stack.pop()
            n = len(out)
            layers = [Block("res", f"output_blocks.{n}.0", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(Block("st", f"output_blocks.{n}.{len(layers)}", ch, ch, cfg.num_heads, ch // cfg.num_heads))
            if level and i == cfg.num_res_blocks:
                layers.append(Block("up", f"output_blocks.{n}.{len(layers)}", ch, ch))
                ds //= 2
            out.append(layers)
    return SDArch(inp, middle, out, chans, ch)


def param_shapes(cfg: SDConfig) -> Dict[str, Tuple[int, ...]]:
    """name -> shape, identical to `UNetModel(**cfg).state_dict()` of the reference."""
    arch = build_arch(cfg)
    mc, ted, cd = cfg.model_channels, cfg.model_channels * 4, cfg.context_dim
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, i, o, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    def conv(name, i, o, k):
        s[name + ".weight"] = (o, i, k, k)
        s[name + ".bias"] = (o,)

    def norm(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)

    def add(b: Block):
        if b.kind == "conv_in":
            conv(b.name, b.cin, b.cout, 3)
        elif b.kind == "res":
            norm(b.name + ".in_layers.0", b.cin)
            conv(b.name + ".in_layers.2", b.cin, b.cout, 3)
            lin(b.name + ".emb_layers.1", ted, b.cout)
            norm(b.name + ".out_layers.0", b.cout)
            conv(b.name + ".out_layers.3", b.cout, b.cout, 3)
            if b.cin != b.cout:
                conv(b.name + ".skip_connection", b.cin, b.cout, 1)
        elif b.kind == "st":
            inner = b.heads * b.d_head
            norm(b.name + ".norm", b.cin)
            conv(b.name + ".proj_in", b.cin, inner, 1)
            for d in range(cfg.transformer_depth):
                t = f"{b.name}.transformer_blocks.{d}"
                for a, kdim in (("attn1", inner), ("attn2", cd)):
                    lin(f"{t}.{a}.to_q", inner, inner, bias=False)
                    lin(f"{t}.{a}.to_k", kdim, inner, bias=False)
                    lin(f"{t}.{a}.to_v", kdim, inner, bias=False)
                    lin(f"{t}.{a}.to_out.0", inner, inner)
                lin(f"{t}.ff.net.0.proj", inner, inner * 4 * 2)
                lin(f"{t}.ff.net.2", inner * 4, inner)
                for k in ("norm1", "norm2", "norm3"):
                    norm(f"{t}.{k}", inner)
            conv(b.name + ".proj_out", inner, b.cin, 1)
        elif b.kind == "down":
            conv(b.name + ".op", b.cin, b.cout, 3)
        elif b.kind == "up":
            conv(b.name + ".conv", b.cin, b.cout, 3)

    for layers in arch.input_blocks:
        for b in layers:
            add(b)
    for b in arch.middle:
        add(b)
    for layers in arch.output_blocks:
        for b in layers:
            add(b)
    norm("out.0", arch.final_ch)
    conv("out.2", mc, cfg.out_channels, 3)
    return s


def make_weights(cfg: SDConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """oracle.weights recipe (every zero_module parameter re-drawn) with the transformer's norms recognised."""
    import zlib

    sd = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        leaf = name.rsplit(".", 2)[-2]
        is_norm = (".in_layers.0." in name or ".out_layers.0." in name or leaf in ("norm", "norm1", "norm2", "norm3")
                   or name.startswith("out.0."))
        if name.endswith(".bias"):
            t = (0.05 if is_norm else 0.02) * torch.randn(shape, generator=g)
        elif is_norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / fan_in ** 0.5
        sd[name] = t.float()
    return sd


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """util.py:152-172 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, name, eps=1e-5):
    return F.group_norm(x.float(), 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def res_block(sd, b: Block, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """openaimodel.py:255-275, use_scale_shift_norm=False, no up/down."""
    h = F.conv2d(F.silu(_gn(x, sd, b.name + ".in_layers.0")), sd[b.name + ".in_layers.2.weight"],
                 sd[b.name + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[b.name + ".emb_layers.1.weight"], sd[b.name + ".emb_layers.1.bias"])
    h = h + emb_out[..., None, None]
    h = F.conv2d(F.silu(_gn(h, sd, b.name + ".out_layers.0")), sd[b.name + ".out_layers.3.weight"],
                 sd[b.name + ".out_layers.3.bias"], padding=1)
    if b.cin != b.cout:
        x = F.conv2d(x, sd[b.name + ".skip_connection.weight"], sd[b.name + ".skip_connection.bias"])
    return x + h


def cross_attention(sd, name: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """attention.py:170-194: softmax(q k^T * dim_head^-0.5) v, heads split as (h d)."""
    ctx = x if context is None else context
    q = F.linear(x, sd[name + ".to_q.weight"])
    k = F.linear(ctx, sd[name + ".to_k.weight"])
    v = F.linear(ctx, sd[name + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return F.linear(out, sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def transformer_block(sd, name: str, x: torch.Tensor, context: torch.Tensor, heads: int) -> torch.Tensor:
    """attention.py:212-216 + FeedForward(glu=True) :46-64."""
    x = cross_attention(sd, name + ".attn1", _ln(x, sd, name + ".norm1"), None, heads) + x
    x = cross_attention(sd, name + ".attn2", _ln(x, sd, name + ".norm2"), context, heads) + x
    h = F.linear(_ln(x, sd, name + ".norm3"), sd[name + ".ff.net.0.proj.weight"], sd[name + ".ff.net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    return F.linear(h, sd[name + ".ff.net.2.weight"], sd[name + ".ff.net.2.bias"]) + x


def spatial_transformer(sd, cfg: SDConfig, b: Block, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """attention.py:245-260."""
    n, c, hh, ww = x.shape
    x_in = x
    x = _gn(x, sd, b.name + ".norm", eps=1e-6)
    x = F.conv2d(x, sd[b.name + ".proj_in.weight"], sd[b.name + ".proj_in.bias"])
    x = x.reshape(n, x.shape[1], hh * ww).permute(0, 2, 1)
    for d in range(cfg.transformer_depth):
        x = transformer_block(sd, f"{b.name}.transformer_blocks.{d}", x, context, b.heads)
    x = x.permute(0, 2, 1).reshape(n, -1, hh, ww)
    x = F.conv2d(x, sd[b.name + ".proj_out.weight"], sd[b.name + ".proj_out.bias"])
    return x + x_in


def _run(sd, cfg, layers: Sequence[Block], h, emb, context):
    for b in layers:
        if b.kind == "conv_in":
            h = F.conv2d(h, sd[b.name + ".weight"], sd[b.name + ".bias"], padding=1)
        elif b.kind == "res":
            h = res_block(sd, b, h, emb)
        elif b.kind == "st":
            h = spatial_transformer(sd, cfg, b, h, context)
        elif b.kind == "down":
            h = F.conv2d(h, sd[b.name + ".op.weight"], sd[b.name + ".op.bias"], stride=2, padding=1)
        elif b.kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[b.name + ".conv.weight"], sd[b.name + ".conv.bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: SDConfig, x: torch.Tensor, timesteps: torch.Tensor,
                 context: torch.Tensor) -> torch.Tensor:
    """openaimodel.py:710-742 (num_classes=None)."""
    arch = build_arch(cfg)
    emb = F.linear(timestep_embedding(timesteps, cfg.model_channels), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x.float()
    for layers in arch.input_blocks:
        h = _run(sd, cfg, layers, h, emb, context)
        hs.append(h)
    h = _run(sd, cfg, arch.middle, h, emb, context)
    for layers in arch.output_blocks:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run(sd, cfg, layers, h, emb, context)
    h = F.silu(_gn(h, sd, "out.0"))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# --------------------------------------------------------------------------------------
# schedule + searched-timestep DDIM with classifier-free guidance
# --------------------------------------------------------------------------------------
def sd_alphas_cumprod(n_timestep: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.0120) -> torch.Tensor:
    """make_beta_schedule("linear") (util.py:21-26) -> DDPM.register_schedule (ldm/models/diffusion/ddpm.py:117-134):
    float64 cumprod, stored as a float32 buffer. linear_start/end from v1-inference_coco.yaml:6-7."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    betas = betas.numpy()
    return torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)


def ddim_tables(alphas_cumprod: torch.Tensor, sampled_timestep: Sequence[int]):
    """DDIMSampler.sample sorts the searched steps (ddim.py:93-94); make_schedule uses them verbatim as
    ddim_timesteps (:30); make_ddim_sampling_parameters (util.py:63-75) with eta = 0:
    alphas = acp[steps] (fp32), alphas_prev = [acp[0]] + acp[steps[:-1]], sqrt(1 - alphas) in fp32."""
    steps = sorted(int(t) for t in sampled_timestep)
    acp = alphas_cumprod.float()
    alphas = acp[steps]
    alphas_prev = torch.tensor([acp[0].item()] + acp[steps[:-1]].tolist(), dtype=torch.float32)
    return steps, alphas, alphas_prev, torch.sqrt(1.0 - alphas)


def ddim_step(x, e_t, a_t, a_prev, sqrt_one_minus_at):
    """p_sample_ddim (ddim.py:200-216) with sigma_t = 0 (eta = 0): every tensor op in fp32, in the reference's order.
    The noise term `sigma_t * randn * temperature` is exactly 0 and x_prev + 0 == x_prev."""
    shape = (x.shape[0], 1, 1, 1)
    a_t = torch.full(shape, float(a_t))
    a_prev = torch.full(shape, float(a_prev))
    sigma_t = torch.full(shape, 0.0)
    s1m = torch.full(shape, float(sqrt_one_minus_at))
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


@torch.no_grad()
def ddim_sample(apply_model, x_T: torch.Tensor, cond: torch.Tensor, uncond: Optional[torch.Tensor], scale: float,
                sampled_timestep: Sequence[int], alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """DDIMSampler.sample(..., sampled_timestep=cand, eta=0, unconditional_guidance_scale=scale,
    unconditional_conditioning=uncond, x_T=x_T) (ddim.py:59-175): descending over the sorted searched steps."""
    steps, alphas, alphas_prev, s1m = ddim_tables(alphas_cumprod, sampled_timestep)
    img = x_T
    b = x_T.shape[0]
    for i, step in enumerate(reversed(steps)):
        index = len(steps) - i - 1
        ts = torch.full((b,), step, dtype=torch.long)
        if uncond is None or scale == 1.0:
            e_t = apply_model(img, ts, cond)
        else:
            e_u, e_c = apply_model(torch.cat([img] * 2), torch.cat([ts] * 2), torch.cat([uncond, cond])).chunk(2)
            e_t = e_u + scale * (e_c - e_u)
        img, _ = ddim_step(img, e_t, alphas[index], alphas_prev[index], s1m[index])
    return img


@torch.no_grad()
def plms_sample(apply_model, x_T: torch.Tensor, cond: torch.Tensor, uncond: Optional[torch.Tensor], scale: float,
                sampled_timestep: Sequence[int], alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """PLMSSampler.sample(..., sampled_timestep=cand) (plms.py:62-257): same tables as DDIM (make_schedule :24-60);
    the first step is a pseudo improved Euler step (one extra model call at t_next), later steps use 2nd/3rd/4th order
    Adams-Bashforth combinations of the last eps values; every combination feeds the eta = 0 update of ddim_step."""
    steps, alphas, alphas_prev, s1m = ddim_tables(alphas_cumprod, sampled_timestep)
    time_range = list(reversed(steps))
    img = x_T
    b = x_T.shape[0]
    old_eps: List[torch.Tensor] = []

    def model_out(x, t):
        if uncond is None or scale == 1.0:
            return apply_model(x, t, cond)
        e_u, e_c = apply_model(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uncond, cond])).chunk(2)
        return e_u + scale * (e_c - e_u)

    for i, step in enumerate(time_range):
        index = len(steps) - i - 1
        ts = torch.full((b,), step, dtype=torch.long)
        ts_next = torch.full((b,), time_range[min(i + 1, len(time_range) - 1)], dtype=torch.long)
        e_t = model_out(img, ts)
        if len(old_eps) == 0:
            x_prev, _ = ddim_step(img, e_t, alphas[index], alphas_prev[index], s1m[index])
            e_t_next = model_out(x_prev, ts_next)
            e_t_prime = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:
            e_t_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_t_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_t_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, _ = ddim_step(img, e_t_prime, alphas[index], alphas_prev[index], s1m[index])
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


# --------------------------------------------------------------------------------------
# DPM-Solver++(2M) with searched time steps (the sampler search_dpm_solver.sh uses)
# --------------------------------------------------------------------------------------
def _interp(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """interpolate_fn (dpm_solver.py:1149-1188) for one channel: piecewise-linear through (xp, yp) with xp ascending,
    the outermost segments extended beyond the ends; same arithmetic `y0 + (x - x0) * (y1 - y0) / (x1 - x0)`."""
    K = xp.shape[0]
    idx = torch.searchsorted(xp, x.contiguous(), right=False)  # number of knots < x
    lo = torch.clamp(idx - 1, 0, K - 2)
    x0, x1, y0, y1 = xp[lo], xp[lo + 1], yp[lo], yp[lo + 1]
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0)


class DiscreteVP:
    """NoiseScheduleVP('discrete', alphas_cumprod=acp) (dpm_solver.py:97-156): log alpha_t by interpolation over
    t_n = n / N, n = 1..N."""

    def __init__(self, alphas_cumprod: torch.Tensor):
        self.log_alpha = 0.5 * torch.log(alphas_cumprod.float())
        self.N = self.log_alpha.shape[0]
        self.t_array = torch.linspace(0.0, 1.0, self.N + 1)[1:]

    def log_mean(self, t):
        return _interp(t, self.t_array, self.log_alpha)

    def alpha(self, t):
        return torch.exp(self.log_mean(t))

    def std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_mean(t)))

    def lam(self, t):
        lm = self.log_mean(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))


def dpm_timesteps(ea_timesteps: Sequence[float], N: int = 1000) -> torch.Tensor:
    """dpm_solver.py:1079-1091: integer candidates index the reversed 1001-point uniform grid between t_T = 1 and
    t_0 = 1/N IN THE ORDER GIVEN; candidates already in (0, 1] are sorted descending."""
    if max(ea_timesteps) > 1:
        full = list(torch.linspace(1.0, 1.0 / N, 1000 + 1))
        full.reverse()
        return torch.Tensor([full[int(ea)].item() for ea in ea_timesteps])
    return torch.Tensor(sorted(ea_timesteps, reverse=True))


@torch.no_grad()
def dpm_solver_sample(apply_model, x_T: torch.Tensor, cond: torch.Tensor, uncond: Optional[torch.Tensor], scale: float,
                      ea_timesteps: Sequence[float], alphas_cumprod: torch.Tensor, record=None) -> torch.Tensor:
    """DPMSolverSampler.sample(S=len(cand) - 1, ..., sampled_timestep=cand): data-prediction multistep solver of order 2
    with a first-order start and (fewer than 15 steps) a first-order final step."""
    ns = DiscreteVP(alphas_cumprod)
    ts = dpm_timesteps(ea_timesteps, ns.N)
    steps = ts.shape[0] - 1
    b = x_T.shape[0]

    def e4(v):
        return v[:, None, None, None]

    def model_fn(x, t):  # model_wrapper "classifier-free" (:336-343) + data_prediction_fn (:386-391)
        t_in = (t - 1.0 / ns.N) * 1000.0  # get_model_input_time (:278-286): fractional model timesteps
        if scale == 1.0 or uncond is None:
            noise = apply_model(x, t_in, cond)
        else:
            n_u, n_c = apply_model(torch.cat([x] * 2), torch.cat([t_in] * 2), torch.cat([uncond, cond])).chunk(2)
            noise = n_u + scale * (n_c - n_u)
        if record is not None:
            record.append(float(t_in[0]))
        return (x - e4(ns.std(t)) * noise) / e4(ns.alpha(t))

    def first_update(x, s, t, m_s):  # dpm_solver_first_update, predict_x0 branch (:519-533)
        h = ns.lam(t) - ns.lam(s)
        alpha_t = torch.exp(ns.log_mean(t))
        return e4(ns.std(t) / ns.std(s)) * x - e4(alpha_t * torch.expm1(-h)) * m_s

    def second_update(x, m1, m0, t1, t0, t):  # multistep_dpm_solver_second_update (:770-790), 'dpm_solver' type
        l1, l0, lt = ns.lam(t1), ns.lam(t0), ns.lam(t)
        alpha_t = torch.exp(ns.log_mean(t))
        h_0, h = l0 - l1, lt - l0
        r0 = h_0 / h
        D1_0 = e4(1.0 / r0) * (m0 - m1)
        return (e4(ns.std(t) / ns.std(t0)) * x - e4(alpha_t * (torch.exp(-h) - 1.0)) * m0
                - 0.5 * e4(alpha_t * (torch.exp(-h) - 1.0)) * D1_0)

    x = x_T
    vec = lambda k: ts[k].expand(b)
    models, times = [model_fn(x, vec(0))], [vec(0)]
    x = first_update(x, times[-1], vec(1), models[-1])
    models.append(model_fn(x, vec(1)))
    times.append(vec(1))
    for step in range(2, steps + 1):
        order = min(2, steps + 1 - step) if steps < 15 else 2
        if order == 1:
            x = first_update(x, times[-1], vec(step), models[-1])
        else:
            x = second_update(x, models[0], models[1], times[0], times[1], vec(step))
        models[0], times[0] = models[1], times[1]
        times[1] = vec(step)
        if step < steps:
            models[1] = model_fn(x, vec(step))
    return x
