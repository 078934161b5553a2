"""GPU parity of every C-ABI op against the CPU oracle (plain fp32 torch restatements).

Inputs are rounded to bf16 first (the kernels' storage type), the oracle then computes in
fp32 on the CPU, so the comparison isolates kernel correctness from input quantisation.
Tolerances are stated per test: bf16 outputs carry 2^-9 relative rounding error; integer /
index / fp32-elementwise paths are bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import diffusion_ref, unet_ref

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from autodiffusion_b200 import ops

    return ops


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _nhwc(x):  # fp32 NCHW (cpu) -> bf16 NHWC (cuda)
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def _nchw(x):  # bf16 NHWC (cuda) -> fp32 NCHW (cpu)
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def _check(out, ref, rel, what):
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{what}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err / max(scale, 1e-30):.4g}")
    assert math.isfinite(err), what
    assert err <= rel * scale + 1e-6, f"{what}: err {err} > {rel} * {scale}"


# ---------------------------------------------------------------- conv / GEMM (tcgen05)
@pytest.mark.parametrize(
    "n,res,cin,cout",
    [
        (2, 64, 192, 192),   # bw=64,bh=2
        (2, 32, 384, 384),   # 2 N tiles
        (2, 16, 576, 576),
        (4, 8, 768, 768),    # two images per 128-row tile
        (3, 8, 768, 768),    # ragged last tile (odd n at 8x8)
        (1, 32, 64, 64),     # N tile 64
        (1, 32, 128, 128),   # N tile 128
        (1, 16, 96, 160),    # cin not a multiple of 64 (zero-filled K tail), padded cout
        (2, 64, 384, 192),
    ],
)
def test_conv3x3(n, res, cin, cout):
    ops = _ops()
    x = _bf(_rand((n, cin, res, res), 1))
    w = _bf(_rand((cout, cin, 3, 3), 2, 1.0 / math.sqrt(9 * cin)))
    b = _rand((cout,), 3, 0.1)
    ref = F.conv2d(x, w, b, padding=1)
    wp = ops.pack_conv_weight([w]).to(DEV)
    out = ops.conv_igemm([(_nhwc(x), 9)], wp, b.to(DEV), cout)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -7, f"conv3x3 n{n} r{res} {cin}->{cout}")


@pytest.mark.parametrize("b,t,c,cout", [(2, 1024, 384, 1152), (3, 64, 768, 768), (2, 256, 576, 1728)])
def test_gemm_k1(b, t, c, cout):
    """nn.Conv1d k=1 (qkv / proj_out) as a 1-tap GEMM over [b*t, c]."""
    ops = _ops()
    x = _bf(_rand((b, c, t), 4))
    w = _bf(_rand((cout, c, 1), 5, 1.0 / math.sqrt(c)))
    bias = _rand((cout,), 6, 0.1)
    ref = F.conv1d(x, w, bias)  # [b, cout, t]
    side = int(math.isqrt(t))
    act = x.permute(0, 2, 1).contiguous().to(torch.bfloat16).to(DEV).view(b, side, side, c)
    out = ops.conv_igemm([(act, 1)], ops.pack_conv_weight([w]).to(DEV), bias.to(DEV), cout)
    torch.cuda.synchronize()
    got = out.float().view(b, t, cout).permute(0, 2, 1).cpu()
    _check(got, ref, 2 ** -7, f"gemm_k1 b{b} t{t} {c}->{cout}")


@pytest.mark.parametrize("res_mode", ["same", "avgpool", "nearest"])
def test_conv_fused_skip_and_residual(res_mode):
    """Second ResBlock conv with the 1x1 skip over a two-source concat folded in as extra
    K-segments (dynamic_unet.py:271 + :699), and the identity-skip residual variants."""
    ops = _ops()
    n, r, c = 2, 16, 192
    h = _bf(_rand((n, c, r, r), 7))
    w2 = _bf(_rand((c, c, 3, 3), 8, 1.0 / math.sqrt(9 * c)))
    b2 = _rand((c,), 9, 0.1)
    if res_mode == "same":
        xa, xb = _bf(_rand((n, 256, r, r), 10)), _bf(_rand((n, 128, r, r), 11))
        ws = _bf(_rand((c, 384, 1, 1), 12, 1.0 / math.sqrt(384)))
        bs = _rand((c,), 13, 0.1)
        ref = F.conv2d(h, w2, b2, padding=1) + F.conv2d(torch.cat([xa, xb], 1), ws, bs)
        wp = ops.pack_conv_weight([w2, ws[:, :256], ws[:, 256:]]).to(DEV)
        out = ops.conv_igemm([(_nhwc(h), 9), (_nhwc(xa), 1), (_nhwc(xb), 1)], wp, (b2 + bs).to(DEV), c)
        torch.cuda.synchronize()
        _check(_nchw(out), ref, 2 ** -7, "conv fused 1x1 skip over concat")
        # identity residual
        xr = _bf(_rand((n, c, r, r), 14))
        out = ops.conv_igemm([(_nhwc(h), 9)], ops.pack_conv_weight([w2]).to(DEV), b2.to(DEV), c,
                             residual=_nhwc(xr), res_mode=ops.RES_SAME)
        torch.cuda.synchronize()
        _check(_nchw(out), F.conv2d(h, w2, b2, padding=1) + xr, 2 ** -7, "conv + identity residual")
    elif res_mode == "avgpool":
        xr = _bf(_rand((n, c, 2 * r, 2 * r), 15))
        out = ops.conv_igemm([(_nhwc(h), 9)], ops.pack_conv_weight([w2]).to(DEV), b2.to(DEV), c,
                             residual=_nhwc(xr), res_mode=ops.RES_AVGPOOL2)
        torch.cuda.synchronize()
        _check(_nchw(out), F.conv2d(h, w2, b2, padding=1) + F.avg_pool2d(xr, 2, 2), 2 ** -7, "conv + avgpool residual")
    else:
        xr = _bf(_rand((n, c, r // 2, r // 2), 16))
        out = ops.conv_igemm([(_nhwc(h), 9)], ops.pack_conv_weight([w2]).to(DEV), b2.to(DEV), c,
                             residual=_nhwc(xr), res_mode=ops.RES_NEAREST2)
        torch.cuda.synchronize()
        _check(_nchw(out), F.conv2d(h, w2, b2, padding=1) + F.interpolate(xr, scale_factor=2, mode="nearest"),
               2 ** -7, "conv + nearest-upsample residual")


def test_conv_out_f32_nchw():
    """Final `out` conv 192 -> 6, fp32 NCHW result (dynamic_unet.py:650-654,702)."""
    ops = _ops()
    n, r, cin, cout = 2, 64, 192, 6
    x = _bf(_rand((n, cin, r, r), 17))
    w = _bf(_rand((cout, cin, 3, 3), 18, 1.0 / math.sqrt(9 * cin)))
    b = _rand((cout,), 19, 0.1)
    out = ops.conv_igemm([(_nhwc(x), 9)], ops.pack_conv_weight([w]).to(DEV), b.to(DEV), cout,
                         out_mode=ops.OUT_F32_NCHW)
    torch.cuda.synchronize()
    assert out.shape == (n, cout, r, r) and out.dtype == torch.float32
    _check(out.cpu(), F.conv2d(x, w, b, padding=1), 1e-4, "out conv fp32 NCHW")


# ---------------------------------------------------------------- attention
@pytest.mark.parametrize("b,t,heads", [(2, 64, 12), (2, 256, 9), (1, 1024, 6), (3, 64, 2)])
@pytest.mark.parametrize("legacy", [False, True])
def test_attention(b, t, heads, legacy):
    ops = _ops()
    c = heads * 64
    qkv = _bf(_rand((b, 3 * c, t), 20 + t + heads, 1.5))
    ref = unet_ref.qkv_attention(qkv, heads, new_order=not legacy)  # [b, c, t]
    rows = qkv.permute(0, 2, 1).contiguous().to(torch.bfloat16).to(DEV).view(b * t, 3 * c)
    out = ops.attention(rows, b, t, heads, legacy)
    torch.cuda.synchronize()
    got = out.float().view(b, t, c).permute(0, 2, 1).cpu()
    _check(got, ref, 2 ** -6, f"attention b{b} t{t} h{heads} legacy={legacy}")


# ---------------------------------------------------------------- GroupNorm family
@pytest.mark.parametrize("n,res,c0,c1", [(2, 64, 192, 0), (2, 16, 768, 576), (3, 8, 768, 768), (2, 32, 64, 32), (1, 16, 1536, 0)])
@pytest.mark.parametrize("variant", ["silu", "plain", "film", "down", "up"])
def test_groupnorm(n, res, c0, c1, variant):
    ops = _ops()
    c = c0 + c1
    x0 = _bf(_rand((n, c0, res, res), 30, 2.0) + 0.5)
    x1 = _bf(_rand((n, c1, res, res), 31, 0.5) - 1.0) if c1 else None
    x = torch.cat([x0, x1], 1) if c1 else x0
    g = 1 + 0.1 * _rand((c,), 32)
    bt = 0.1 * _rand((c,), 33)
    hn = unet_ref.group_norm32(x, g, bt)
    kw = dict(silu=True)
    if variant == "plain":
        ref, kw = hn, dict(silu=False)
    elif variant == "silu":
        ref = F.silu(hn)
    elif variant == "film":
        ss = 0.3 * _rand((n, 2 * c + 5), 34)  # row stride larger than 2c on purpose
        scale, shift = ss[:, :c, None, None], ss[:, c:2 * c, None, None]
        ref = F.silu(hn * (1 + scale) + shift)
        kw = dict(silu=True, scale_shift=ss.to(DEV), ss_stride=2 * c + 5)
    elif variant == "down":
        ref = F.avg_pool2d(F.silu(hn), 2, 2)
        kw = dict(silu=True, resample=ops.RESAMPLE_AVGPOOL2)
    else:
        ref = F.interpolate(F.silu(hn), scale_factor=2, mode="nearest")
        kw = dict(silu=True, resample=ops.RESAMPLE_NEAREST2)
    out = ops.groupnorm(_nhwc(x0), g.to(DEV), bt.to(DEV), src1=_nhwc(x1) if c1 else None, **kw)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -7, f"groupnorm {variant} n{n} r{res} c{c0}+{c1}")


@pytest.mark.parametrize("mode", ["down", "up"])
def test_resample2x(mode):
    ops = _ops()
    x = _bf(_rand((2, 192, 16, 16), 35))
    if mode == "down":
        ref, m = F.avg_pool2d(x, 2, 2), ops.RESAMPLE_AVGPOOL2
    else:
        ref, m = F.interpolate(x, scale_factor=2, mode="nearest"), ops.RESAMPLE_NEAREST2
    out = ops.resample2x(_nhwc(x), m)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -8 if mode == "down" else 0.0, f"resample2x {mode}")


# ---------------------------------------------------------------- small fp32 kernels
def test_stem_conv():
    ops = _ops()
    x = _rand((3, 3, 64, 64), 40)
    w = _rand((192, 3, 3, 3), 41, 0.2)
    b = _rand((192,), 42, 0.1)
    out = ops.stem_conv(x.to(DEV), w.to(DEV), b.to(DEV))
    torch.cuda.synchronize()
    _check(_nchw(out), F.conv2d(x, w, b, padding=1), 2 ** -8, "stem conv")


def test_timestep_embedding_and_linear():
    ops = _ops()
    t = torch.tensor([0, 1, 85, 137, 676, 971, 999], dtype=torch.int64)
    ref = unet_ref.timestep_embedding(t, 192)
    out = ops.timestep_embedding(t.to(DEV), 192)
    torch.cuda.synchronize()
    _check(out.cpu(), ref, 2e-6, "timestep_embedding")  # same fp32 args; sin/cos differ by ulps

    b, k, nout = 7, 768, 1000
    x, w, bias = _rand((b, k), 43), _rand((nout, k), 44, 1 / math.sqrt(k)), _rand((nout,), 45, 0.1)
    table, idx = _rand((50, nout), 46), torch.randint(0, 50, (b,), generator=torch.Generator().manual_seed(47))
    out = ops.linear(x.to(DEV), w.to(DEV), bias.to(DEV), silu_in=True, table=table.to(DEV), idx=idx.to(DEV))
    torch.cuda.synchronize()
    _check(out.cpu(), F.linear(F.silu(x), w, bias) + table[idx], 1e-4, "linear(silu)+table")  # fp32, K=768 summation order
    out = ops.linear(x.to(DEV), w.to(DEV), None)
    torch.cuda.synchronize()
    _check(out.cpu(), F.linear(x, w), 1e-4, "linear plain")


@pytest.mark.parametrize("with_grad", [False, True])
@pytest.mark.parametrize("clip", [False, True])
def test_ddim_step_bit_exact(with_grad, clip):
    """Given identical eps / grad, x_{t-1} must equal the reference's op chain bit for bit."""
    ops = _ops()
    from autodiffusion_b200.gaussian_diffusion import ddim_coefficients

    base = diffusion_ref.base_tables("cosine", 1000)
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], [153, 424, 926, 690])
    tables = diffusion_ref.diffusion_tables(nb)
    n = 4
    x = _rand((n, 3, 64, 64), 50)
    mo = _rand((n, 6, 64, 64), 51)
    g = _rand((n, 3, 64, 64), 52, 3.0)
    for i in range(len(tmap)):
        calls = {}

        def model(xx, ts, **kw):
            calls["t"] = ts
            return mo

        cond = (lambda xx, ts, **kw: g) if with_grad else None
        one = {k: v[i:i + 1] for k, v in tables.items()}  # a 1-step "schedule" = step i in isolation
        ref = diffusion_ref.ddim_sample_loop(model, x.shape, one, [tmap[i]], x, clip_denoised=clip, cond_fn=cond)
        out = ops.ddim_step(x.to(DEV), mo.to(DEV), g.to(DEV) if with_grad else None,
                            ddim_coefficients(tables, i), clip_denoised=clip)
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), ref), f"step {i}: max diff {(out.cpu() - ref).abs().max().item()}"


def test_pack_uint8_bit_exact():
    ops = _ops()
    s = _rand((5, 3, 64, 64), 53, 0.8)
    s.view(-1)[:6] = torch.tensor([-1.0, 1.0, -1.5, 1.5, 0.0, 0.999])
    out = ops.pack_uint8(s.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), diffusion_ref.pack_uint8(s))


def test_moments_accumulate():
    ops = _ops()
    d = 2048
    sx = torch.zeros(d, dtype=torch.float64, device=DEV)
    sxx = torch.zeros(d, d, dtype=torch.float64, device=DEV)
    chunks = [_rand((300, d), 60) + 0.3, _rand((211, d), 61) * 2 - 0.1]
    for ch in chunks:
        ops.moments_accumulate(ch.to(DEV), sx, sxx)
    torch.cuda.synchronize()
    f = torch.cat(chunks).double().numpy()
    np.testing.assert_allclose(sx.cpu().numpy(), f.sum(0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(sxx.cpu().numpy(), f.T @ f, rtol=1e-12, atol=1e-9)


def test_plan_records_and_replays():
    ops = _ops()
    x = _bf(_rand((2, 192, 16, 16), 70))
    w = _bf(_rand((192, 192, 3, 3), 71, 1.0 / math.sqrt(9 * 192)))
    plan = ops.Plan()
    act = _nhwc(x)
    out = ops.conv_igemm([(act, 9)], ops.pack_conv_weight([w]).to(DEV), None, 192, plan=plan)
    assert plan.num_ops() == 1
    out.zero_()
    launches = plan.run()
    torch.cuda.synchronize()
    assert launches == 1
    _check(_nchw(out), F.conv2d(x, w, None, padding=1), 2 ** -7, "plan replay")


def test_errors_are_reported_not_swallowed():
    ops = _ops()
    from autodiffusion_b200._lib import AdbError

    with pytest.raises(RuntimeError):
        ops.pack_uint8(torch.zeros(1, 3, 8, 8))  # CPU tensor: no CPU path
    with pytest.raises(AdbError):
        ops.attention(torch.zeros(2 * 100, 3 * 64, dtype=torch.bfloat16, device=DEV), 2, 100, 1, False)


def test_linear_tc_matches_fp32_linear():
    """The emb_layers product on tensor cores with split-bf16 operands vs the fp32 Linear the reference keeps
    (fp16_util.py:15-22): error <= 2^-13 of the output scale (three bf16 cross products, fp32 accumulation)."""
    ops = _ops()
    b, k, n = 37, 768, 35712 // 8
    x = _rand((b, k), 90, 1.5)
    w = _rand((n, k), 91, k ** -0.5)
    bias = 0.1 * _rand((n,), 92)
    ref = F.linear(F.silu(x), w, bias)
    out = ops.linear_tc(x.to(DEV), ops.pack_linear_weight_split(w, DEV), bias.to(DEV), n, silu_in=True)
    torch.cuda.synchronize()
    assert out.shape == (b, n) and out.dtype == torch.float32
    _check(out.cpu(), ref, 2 ** -13, "linear_tc (split bf16) vs fp32 linear")


def test_stem_conv_tc_matches_fp32_conv():
    """Tensor-core stem (im2col with hi|lo split of the input + one 64-deep k-step) vs the fp32 conv of the input
    with bf16-rounded weights: only the output's bf16 rounding remains (2^-8 of the output scale)."""
    ops = _ops()
    n, r, cout = 3, 64, 192
    x = _rand((n, 3, r, r), 93, 1.3)
    w = _bf(_rand((cout, 3, 3, 3), 94, 27 ** -0.5))
    b = 0.1 * _rand((cout,), 95)
    ref = F.conv2d(x, w, b, padding=1)
    stats = torch.zeros((n, 32, 2), dtype=torch.float64, device=DEV)
    out = ops.stem_conv_tc(x.to(DEV), ops.pack_stem_weight(w, DEV), b.to(DEV), cout, stats_out=stats)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -8, "stem_conv_tc")
    got = _nchw(out).double()
    s_ref = got.reshape(n, 32, -1).sum(-1)
    assert (stats[:, :, 0].cpu() - s_ref).abs().max().item() <= 1e-3 * s_ref.abs().max().item() + 1e-2
