"""GPU parity of the ops added for the Stable-Diffusion-v1 family against fp32 torch restatements on the CPU
(oracle/sd_unet_ref.py for the composite pieces). Inputs are rounded to bf16 first, as in tests/test_ops_gpu.py.
Tolerances: tensor-core GEMM outputs 2^-7 of the output range, attention 2^-6, memory-bound bf16 ops 2^-7,
the fp32 sampler step bit-exact against the reference's own recorded outputs."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sd_unet_ref as R
from tests.util import golden

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from autodiffusion_b200 import ops

    return ops


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def _check(out, ref, rel, what):
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{what}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err / max(scale, 1e-30):.4g}")
    assert math.isfinite(err), what
    assert err <= rel * scale + 1e-6, f"{what}: err {err} > {rel} * {scale}"


@pytest.mark.parametrize("n,res,cin,cout", [(2, 64, 320, 320), (2, 32, 640, 640), (3, 16, 1280, 1280), (2, 16, 64, 64),
                                            (1, 64, 128, 128)])
def test_conv3x3_stride2(n, res, cin, cout):
    """Downsample.op (openaimodel.py:147-149): 3x3, stride 2, padding 1 - the 5-D TMA view of the input."""
    ops = _ops()
    x = _bf(_rand((n, cin, res, res), 1))
    w = _bf(_rand((cout, cin, 3, 3), 2, 1.0 / math.sqrt(9 * cin)))
    b = _rand((cout,), 3, 0.1)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    out = ops.conv_igemm([(_nhwc(x), 9, 2)], ops.pack_conv_weight([w]).to(DEV), b.to(DEV), cout)
    torch.cuda.synchronize()
    assert out.shape == (n, res // 2, res // 2, cout)
    _check(_nchw(out), ref, 2 ** -7, f"conv3x3 stride 2 n{n} r{res} {cin}->{cout}")


@pytest.mark.parametrize("n,res,cin,cout", [(3, 32, 320, 320), (4, 8, 1280, 1280), (5, 8, 64, 64), (2, 16, 640, 1280)])
def test_conv_per_image_bias_and_stats(n, res, cin, cout):
    """ResBlock without scale-shift (openaimodel.py:255-275): h = conv(x) + emb_out[n, :, None, None] in the conv
    epilogue, with the next GroupNorm's sums of the stored result."""
    ops = _ops()
    x = _bf(_rand((n, cin, res, res), 1))
    w = _bf(_rand((cout, cin, 3, 3), 2, 1.0 / math.sqrt(9 * cin)))
    emb = _rand((n, 2 * cout + 8), 3, 0.5)  # a column slice of a wider fp32 matrix: row stride != cout
    ref = F.conv2d(x, w, None, padding=1) + emb[:, 8:8 + cout, None, None]
    stats = torch.zeros((n, 32, 2), dtype=torch.float64, device=DEV)
    out = ops.conv_igemm([(_nhwc(x), 9)], ops.pack_conv_weight([w]).to(DEV), emb.to(DEV)[:, 8:8 + cout], cout,
                         stats_out=stats)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -7, f"conv + per-image bias n{n} r{res} {cin}->{cout}")
    o = _nchw(out).double().reshape(n, 32, -1)
    want = torch.stack([o.sum(-1), (o * o).sum(-1)], dim=-1)
    assert torch.allclose(stats.cpu(), want, rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("rows,c", [(4096, 320), (1000, 640), (257, 1280), (64, 64), (33, 2048)])
def test_layernorm(rows, c):
    ops = _ops()
    x = _bf(_rand((rows, c), 1, 2.0) + 0.3)
    g, b = 1.0 + _rand((c,), 2, 0.1), _rand((c,), 3, 0.1)
    ref = F.layer_norm(x, (c,), g, b, 1e-5)
    out = ops.layernorm(x.to(torch.bfloat16).to(DEV), g.to(DEV), b.to(DEV))
    torch.cuda.synchronize()
    _check(out.float().cpu(), ref, 2 ** -7, f"layernorm {rows}x{c}")


@pytest.mark.parametrize("rows,inner", [(512, 1280), (100, 2560), (7, 256)])
def test_geglu(rows, inner):
    ops = _ops()
    x = _bf(_rand((rows, 2 * inner), 1, 1.5))
    a, gate = x.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    out = ops.geglu(x.to(torch.bfloat16).to(DEV))
    torch.cuda.synchronize()
    _check(out.float().cpu(), ref, 2 ** -7, f"geglu {rows}x{inner}")


def _attn_ref(q, k, v, heads, d):
    b, tq, _ = q.shape
    tk = k.shape[1]

    def split(t, n):
        return t.reshape(b, n, heads, d).permute(0, 2, 1, 3)

    qh, kh, vh = split(q, tq), split(k, tk), split(v, tk)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(b, tq, heads * d)


def _pad_heads(t, heads, d, d_pad):  # [b, n, heads*d] -> [b, n, heads*d_pad] with zero columns
    b, n, _ = t.shape
    o = torch.zeros(b, n, heads, d_pad)
    o[..., :d] = t.reshape(b, n, heads, d)
    return o.reshape(b, n, heads * d_pad)


@pytest.mark.parametrize("b,heads,d,t", [(2, 8, 40, 4096), (2, 8, 80, 1024), (3, 8, 160, 256), (5, 8, 160, 64),
                                         (2, 4, 8, 256), (2, 8, 16, 1024), (1, 2, 32, 64), (2, 8, 64, 256)])
def test_attention_sd_self(b, heads, d, t):
    """attn1 of BasicTransformerBlock: q, k, v side by side in one projection output, heads padded to 64-column chunks."""
    ops = _ops()
    dp = ((d + 63) // 64) * 64
    q, k, v = (_bf(_rand((b, t, heads * d), s, 1.0)) for s in (1, 2, 3))
    ref = _attn_ref(q, k, v, heads, d)
    qkv = torch.cat([_pad_heads(x, heads, d, dp) for x in (q, k, v)], dim=-1).to(torch.bfloat16).to(DEV)
    out = ops.attention_sd(qkv, qkv, b, heads, d, dp, t, t, t, 0, heads * dp, 2 * heads * dp)
    torch.cuda.synchronize()
    o = out.float().cpu().reshape(b, t, heads, dp)
    if dp > d:
        assert o[..., d:].abs().max().item() == 0.0  # the padding columns come out exactly zero
    _check(o[..., :d].reshape(b, t, heads * d), ref, 2 ** -6, f"self-attention b{b} h{heads} d{d} t{t}")
    if dp > d:  # denominators from the tensor core: a column of ones in V's padding (v_ones)
        qkv4 = qkv.view(b, t, 3, heads, dp).clone()
        qkv4[:, :, 2, :, d] = 1.0
        out1 = ops.attention_sd(qkv4.view(b, t, -1), qkv4.view(b, t, -1), b, heads, d, dp, t, t, t, 0, heads * dp, 2 * heads * dp,
                                v_ones=True)
        torch.cuda.synchronize()
        o1 = out1.float().cpu().reshape(b, t, heads, dp)
        assert o1[..., d:].abs().max().item() == 0.0
        _check(o1[..., :d].reshape(b, t, heads * d), ref, 2 ** -6, f"self-attention (v_ones) b{b} h{heads} d{d} t{t}")


@pytest.mark.parametrize("b,heads,d,tq,tk", [(2, 8, 40, 4096, 77), (2, 8, 80, 1024, 77), (3, 8, 160, 64, 77),
                                             (2, 8, 8, 256, 77), (2, 8, 160, 256, 128), (2, 8, 40, 256, 5)])
def test_attention_sd_cross(b, heads, d, tq, tk):
    """attn2: keys / values from the 77-token context stored in a 128-row zero-padded buffer; rows >= tk are masked."""
    ops = _ops()
    dp = ((d + 63) // 64) * 64
    q = _bf(_rand((b, tq, heads * d), 1))
    k, v = _bf(_rand((b, tk, heads * d), 2)), _bf(_rand((b, tk, heads * d), 3))
    ref = _attn_ref(q, k, v, heads, d)
    qd = _pad_heads(q, heads, d, dp).to(torch.bfloat16).to(DEV)
    kv = torch.zeros(b, 128, 2 * heads * dp)
    kv[:, :tk] = torch.cat([_pad_heads(k, heads, d, dp), _pad_heads(v, heads, d, dp)], dim=-1)
    kv[:, tk:] = 3.0  # whatever sits in the masked rows must not matter
    out = ops.attention_sd(qd, kv.to(torch.bfloat16).to(DEV), b, heads, d, dp, tq, 128, tk, 0, 0, heads * dp)
    torch.cuda.synchronize()
    o = out.float().cpu().reshape(b, tq, heads, dp)
    _check(o[..., :d].reshape(b, tq, heads * d), ref, 2 ** -6, f"cross-attention b{b} d{d} tq{tq} tk{tk}")
    kv4 = kv.view(b, 128, 2, heads, dp).clone()
    kv4[:, :, 1, :, d] = 1.0
    out1 = ops.attention_sd(qd, kv4.view(b, 128, -1).to(torch.bfloat16).to(DEV), b, heads, d, dp, tq, 128, tk, 0, 0, heads * dp,
                            v_ones=True)
    torch.cuda.synchronize()
    o1 = out1.float().cpu().reshape(b, tq, heads, dp)
    _check(o1[..., :d].reshape(b, tq, heads * d), ref, 2 ** -6, f"cross-attention (v_ones) b{b} d{d} tq{tq} tk{tk}")


def test_cfg_ddim_step_bit_exact_vs_reference():
    """The fused CFG + DDIM update against x_prev recorded from the reference's p_sample_ddim (sd_small.npz)."""
    ops = _ops()
    g = golden("sd_small.npz")
    steps, a, ap, s1m = R.ddim_tables(R.sd_alphas_cumprod(), g["cand"].tolist())
    x = torch.tensor(g["x_T"]).to(DEV)
    eps = torch.cat([torch.tensor(g["e_u"]), torch.tensor(g["e_c"])]).to(DEV)
    from autodiffusion_b200.sd_ddim import ddim_coefficients

    for index in range(len(steps)):
        out = ops.cfg_ddim_step(x, eps, ddim_coefficients(a, ap, s1m, index), scale=7.5, cfg=True)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g[f"step_x_prev_{index}"])
    # no guidance: eps used as is
    e = torch.tensor(g["e_c"])
    want, _ = R.ddim_step(torch.tensor(g["x_T"]), e, a[1], ap[1], s1m[1])
    out = ops.cfg_ddim_step(x, e.to(DEV), ddim_coefficients(a, ap, s1m, 1))
    assert np.array_equal(out.cpu().numpy(), want.numpy())


def test_pad_context():
    ops = _ops()
    c = _rand((3, 77, 768), 1)
    out = ops.pad_context(c.to(DEV), 128)
    torch.cuda.synchronize()
    assert out.shape == (3, 128, 768)
    assert torch.equal(out[:, :77].cpu(), c.to(torch.bfloat16))
    assert out[:, 77:].abs().max().item() == 0.0


def test_plms_update_bit_exact_vs_reference():
    """The fused PLMS eps combination + update against x_prev recorded from the reference's p_sample_plms with 0..3
    older eps tensors (sd_small_plms.npz): pseudo improved Euler and 2nd / 3rd / 4th order Adams-Bashforth."""
    ops = _ops()
    from autodiffusion_b200.sd_ddim import ddim_coefficients

    g = golden("sd_small_plms.npz")
    steps, a, ap, s1m = R.ddim_tables(R.sd_alphas_cumprod(), g["cand"].tolist())
    coef = ddim_coefficients(a, ap, s1m, 2)
    x = torch.tensor(g["x_T"]).to(DEV)
    es = [torch.tensor(e).to(DEV) for e in g["es"]]
    e_t = ops.cfg_combine(es[0])  # no guidance: a copy
    assert torch.equal(e_t, es[0])
    cases = {0: (1, [es[1]]), 1: (2, [es[3]]), 2: (3, [es[3], es[2]]), 3: (4, [es[3], es[2], es[1]])}
    for n_old, (mode, olds) in cases.items():
        out = ops.plms_update(x, e_t, olds, mode, coef)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g[f"plms_x_prev_{n_old}"]), n_old
    # CFG combine: the same expression as inside the DDIM step kernel
    eu, ec = es[1], es[2]
    comb = ops.cfg_combine(torch.cat([eu, ec]), scale=7.5, cfg=True)
    assert torch.equal(comb.cpu(), (eu.cpu() + 7.5 * (ec.cpu() - eu.cpu())))
