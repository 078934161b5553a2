"""GPU parity AT THE BENCHMARKED SIZES (BASELINE.json configs[1]: batch 256 per GPU).

The small-shape op tests never leave the first wave of a persistent grid. At batch 256 the same kernels run
1 048 576-row implicit GEMMs, >148-tile persistent loops with several accumulator hand-offs per CTA, 2^31-scale
element offsets, 9 GB of pooled activations. Here every hot op is run at those sizes against the oracle's torch
expressions evaluated in fp32 ON THE GPU (cuDNN/cuBLAS with TF32 disabled - test infrastructure only; inputs are
bf16-rounded first so the comparison isolates the kernel), with the same bars as the small tests:
max-abs <= 2^-7 x max|ref| for bf16 tensor-core outputs (attention 2^-6, gradients 2^-6 / 2^-5) plus a relative-RMS
bar of 0.4 % (measured 0.17-0.25 %); fp32 elementwise / integer outputs bit-exact.

The last tests run the benchmarked candidate itself at batch 256 (10 steps + skip mask, native classifier guidance,
one CUDA graph): the first 8 samples use the noise / labels of the reference's own CPU run
(tests/golden/config2_admg64_cand10_guided.npz) and must reproduce its images.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import diffusion_ref, unet_ref

pytestmark = pytest.mark.gpu

DEV = "cuda"
N = 256


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()


def _ops():
    from autodiffusion_b200 import ops

    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(shape, generator=g, device=DEV) * scale


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc(x):  # fp32 NCHW (cuda) -> bf16 NHWC (cuda)
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(x):  # bf16 NHWC -> fp32 NCHW
    return x.float().permute(0, 3, 1, 2)


def _check(out, ref, rel, what, rms_bar=0.004):
    assert out.shape == ref.shape, (out.shape, ref.shape)
    diff = out - ref
    err = diff.abs().max().item()
    scale = ref.abs().max().item()
    rms = (diff.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item()
    print(f"{what}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err / max(scale, 1e-30):.4g} rel_rms={rms:.4g}")
    assert math.isfinite(err), what
    assert err <= rel * scale + 1e-6, f"{what}: err {err} > {rel} * {scale}"
    assert rms <= rms_bar, f"{what}: rel rms {rms} > {rms_bar}"


# ---------------------------------------------------------------- implicit GEMM
@pytest.mark.parametrize("res,cin,cout", [(64, 192, 192), (32, 384, 384), (16, 576, 576), (8, 768, 768)])
def test_conv3x3_batch256(res, cin, cout):
    """The four resolutions' dominant 3x3 convs: 8192 / 2048 x2 / 512 x3 / 128 x4 tiles over <= 148 CTAs."""
    ops = _ops()
    x = _bf(_rand((N, cin, res, res), 1))
    w = _bf(_rand((cout, cin, 3, 3), 2, 1.0 / math.sqrt(9 * cin)))
    b = _rand((cout,), 3, 0.1)
    stats = torch.zeros((N, 32, 2), dtype=torch.float64, device=DEV)
    out = ops.conv_igemm([(_nhwc(x), 9)], ops.pack_conv_weight([w.cpu()]).to(DEV), b, cout, stats_out=stats)
    torch.cuda.synchronize()
    ref = F.conv2d(x, w, b, padding=1)
    _check(_nchw(out), ref, 2 ** -7, f"conv3x3 n{N} r{res} {cin}->{cout}")
    # the fused GroupNorm sums of the STORED (bf16) output: sum and sum of squares per (sample, group), fp64
    o64 = out.double().view(N, res * res, 32, cout // 32)
    want = torch.stack([o64.sum((1, 3)), (o64 * o64).sum((1, 3))], dim=-1)
    rel = ((stats - want).abs().max() / want.abs().max()).item()
    print(f"  fused GroupNorm sums: max rel err {rel:.3g}")
    assert rel <= 1e-6  # fp32 partial sums per 128-row tile, fp64 across tiles (measured 4e-8 .. 1.2e-7)


def test_conv_fused_skip_concat_batch256():
    """Output ResBlock at 32x32 (SURVEY A3 id 47): conv2 3x3 384->384 + 1x1 skip over the [384 | 384] concat in one GEMM."""
    ops = _ops()
    r, c = 32, 384
    h = _bf(_rand((N, c, r, r), 7))
    xa, xb = _bf(_rand((N, 384, r, r), 10)), _bf(_rand((N, 384, r, r), 11))
    w2 = _bf(_rand((c, c, 3, 3), 8, 1.0 / math.sqrt(9 * c)))
    ws = _bf(_rand((c, 768, 1, 1), 12, 1.0 / math.sqrt(768)))
    b = _rand((c,), 9, 0.1)
    wp = ops.pack_conv_weight([w2.cpu(), ws[:, :384].cpu(), ws[:, 384:].cpu()]).to(DEV)
    out = ops.conv_igemm([(_nhwc(h), 9), (_nhwc(xa), 1), (_nhwc(xb), 1)], wp, b, c)
    torch.cuda.synchronize()
    ref = F.conv2d(h, w2, b, padding=1) + F.conv2d(torch.cat([xa, xb], 1), ws)
    _check(_nchw(out), ref, 2 ** -7, "conv2 + 1x1 skip over concat, n256 r32")


@pytest.mark.parametrize("t,c", [(1024, 384), (256, 576), (64, 768)])
def test_qkv_and_proj_gemm_batch256(t, c):
    """AttentionBlock qkv (c -> 3c) and proj_out (+ residual) as 1-tap GEMMs over [256 * t, c] token rows."""
    ops = _ops()
    side = int(math.isqrt(t))
    x = _bf(_rand((N * t, c), 4))
    wq = _bf(_rand((3 * c, c), 5, 1.0 / math.sqrt(c)))
    bq = _rand((3 * c,), 6, 0.1)
    act = x.to(torch.bfloat16).view(N, side, side, c)
    out = ops.conv_igemm([(act, 1)], ops.pack_conv_weight([wq.cpu().unsqueeze(-1)]).to(DEV), bq, 3 * c)
    torch.cuda.synchronize()
    _check(out.float().view(N * t, 3 * c), x @ wq.t() + bq, 2 ** -7, f"qkv gemm t{t} c{c}")
    wp = _bf(_rand((c, c), 7, 1.0 / math.sqrt(c)))
    bp = _rand((c,), 8, 0.1)
    res = _bf(_rand((N * t, c), 9))
    out = ops.conv_igemm([(act, 1)], ops.pack_conv_weight([wp.cpu().unsqueeze(-1)]).to(DEV), bp, c,
                         residual=res.to(torch.bfloat16).view(N, side, side, c), res_mode=ops.RES_SAME)
    torch.cuda.synchronize()
    _check(out.float().view(N * t, c), x @ wp.t() + bp + res, 2 ** -7, f"proj gemm + residual t{t} c{c}")


# ---------------------------------------------------------------- attention
def _attention_ref(qkv_rows, b, t, heads, legacy, dout_rows=None, chunk=16):
    """oracle qkv_attention (and its autograd gradient) in fp32 on the GPU, `chunk` samples at a time."""
    c = heads * 64
    outs, grads = [], []
    for s in range(0, b, chunk):
        q = qkv_rows[s * t:(s + chunk) * t].float().view(-1, t, 3 * c).permute(0, 2, 1).contiguous()
        if dout_rows is not None:
            q.requires_grad_(True)
        o = unet_ref.qkv_attention(q, heads, new_order=not legacy)  # [chunk, c, t]
        if dout_rows is not None:
            do = dout_rows[s * t:(s + chunk) * t].float().view(-1, t, c).permute(0, 2, 1)
            grads.append(torch.autograd.grad((o * do).sum(), q)[0].permute(0, 2, 1).reshape(-1, 3 * c))
        outs.append(o.detach().permute(0, 2, 1).reshape(-1, c))
    return torch.cat(outs), (torch.cat(grads) if grads else None)


@pytest.mark.parametrize("t,heads,legacy", [(1024, 6, False), (256, 9, False), (64, 12, False), (1024, 4, True)])
def test_attention_forward_batch256(t, heads, legacy):
    ops = _ops()
    c = heads * 64
    rows = _rand((N * t, 3 * c), 20 + t, 1.5).to(torch.bfloat16)
    out = ops.attention(rows, N, t, heads, legacy)
    torch.cuda.synchronize()
    ref, _ = _attention_ref(rows, N, t, heads, legacy)
    _check(out.float(), ref, 2 ** -6, f"attention n{N} t{t} h{heads} legacy={legacy}")


@pytest.mark.parametrize("t,heads,fused", [(1024, 4, False), (256, 6, False), (64, 8, False), (1024, 4, True), (256, 6, True)])
def test_attention_backward_batch256(t, heads, fused):
    """The classifier's attention layers (legacy order, width 128: 256 / 384 / 512 channels) at batch 256, in the default
    two-kernel form and in the opt-in single-pass form."""
    ops = _ops()
    prev = ops.set_attention_backward_fused(None)
    ops.set_attention_backward_fused(fused)
    try:
        _attention_backward_batch256(ops, t, heads)
    finally:
        ops.set_attention_backward_fused(prev)


def _attention_backward_batch256(ops, t, heads):
    c = heads * 64
    rows = _rand((N * t, 3 * c), 50 + t, 1.2).to(torch.bfloat16)
    drows = _rand((N * t, c), 51).to(torch.bfloat16)
    lse = torch.empty((N * heads, t), dtype=torch.float32, device=DEV)
    out = ops.attention(rows, N, t, heads, True, lse=lse)
    dqkv = ops.attention_backward(rows, out, drows, lse, N, t, heads, True)
    torch.cuda.synchronize()
    ref_out, ref_grad = _attention_ref(rows, N, t, heads, True, dout_rows=drows)
    _check(out.float(), ref_out, 2 ** -6, f"attention(lse) fwd n{N} t{t} h{heads}")
    _check(dqkv.float(), ref_grad, 2 ** -5, f"attention_backward n{N} t{t} h{heads}")


# ---------------------------------------------------------------- GroupNorm forward / backward
@pytest.mark.parametrize("res,c,variant", [(64, 192, "film"), (32, 384, "silu"), (8, 768, "plain"), (64, 128, "down")])
def test_groupnorm_batch256(res, c, variant):
    ops = _ops()
    x = _bf(_rand((N, c, res, res), 30, 2.0) + 0.5)
    g = 1 + 0.1 * _rand((c,), 31)
    bt = 0.1 * _rand((c,), 32)
    hn = F.group_norm(x, 32, g, bt, eps=1e-5)
    kw = dict(silu=True)
    if variant == "plain":
        ref, kw = hn, dict(silu=False)
    elif variant == "silu":
        ref = F.silu(hn)
    elif variant == "film":
        ss = 0.3 * _rand((N, 2 * c + 5), 33)
        ref = F.silu(hn * (1 + ss[:, :c, None, None]) + ss[:, c:2 * c, None, None])
        kw = dict(silu=True, scale_shift=ss.contiguous(), ss_stride=2 * c + 5)
    else:
        ref = F.avg_pool2d(F.silu(hn), 2, 2)
        kw = dict(silu=True, resample=ops.RESAMPLE_AVGPOOL2)
    out = ops.groupnorm(_nhwc(x), g, bt, **kw)
    torch.cuda.synchronize()
    _check(_nchw(out), ref, 2 ** -7, f"groupnorm {variant} n{N} r{res} c{c}")


@pytest.mark.parametrize("res,c,variant", [(64, 128, "film"), (32, 256, "silu_add"), (16, 384, "plain"), (64, 128, "down_addpool")])
def test_gn_backward_batch256(res, c, variant):
    ops = _ops()
    x = (_bf(_rand((N, c, res, res), 40, 2.0) + 0.5)).requires_grad_(True)
    g = 1 + 0.1 * _rand((c,), 41)
    bt = 0.1 * _rand((c,), 42)
    hn = F.group_norm(x, 32, g, bt, eps=1e-5)
    kw = dict(silu=True)
    ro = res // 2 if variant.startswith("down") else res
    if variant == "plain":
        y, kw = hn, dict(silu=False)
    elif variant == "silu_add":
        y = F.silu(hn)
    elif variant == "film":
        ss = 0.3 * _rand((N, 2 * c + 3), 43)
        y = F.silu(hn * (1 + ss[:, :c, None, None]) + ss[:, c:2 * c, None, None])
        kw = dict(silu=True, scale_shift=ss.contiguous(), ss_stride=2 * c + 3)
    else:
        y = F.avg_pool2d(F.silu(hn), 2, 2)
        kw = dict(silu=True, resample=ops.RESAMPLE_AVGPOOL2)
    dy = _bf(_rand((N, c, ro, ro), 44))
    loss = (y * dy).sum()
    if variant == "silu_add":
        add_ref = _bf(_rand((N, c, res, res), 45))
        loss = loss + (x * add_ref).sum()
        kw.update(add=_nhwc(add_ref), add_mode=ops.RES_SAME)
    elif variant == "down_addpool":
        add_ref = _bf(_rand((N, c, ro, ro), 46))
        loss = loss + (F.avg_pool2d(x, 2, 2) * add_ref).sum()
        kw.update(add=_nhwc(add_ref), add_mode=ops.RES_AVGPOOL2)
    ref = torch.autograd.grad(loss, x)[0]
    xd = _nhwc(x.detach())
    stats = torch.empty((N, 32, 2), dtype=torch.float64, device=DEV)
    fkw = {k: v for k, v in kw.items() if k in ("silu", "scale_shift", "ss_stride", "resample")}
    ops.groupnorm(xd, g, bt, stats=stats, **fkw)
    dx = ops.gn_backward(xd, stats, g, bt, _nhwc(dy), **kw)
    torch.cuda.synchronize()
    _check(_nchw(dx), ref, 2 ** -6, f"gn_backward {variant} n{N} r{res} c{c}")


# ---------------------------------------------------------------- epilogue, pack, moments
def test_ddim_step_and_pack_uint8_batch256_bit_exact():
    ops = _ops()
    x = _rand((N, 3, 64, 64), 60)
    mo = _rand((N, 6, 64, 64), 61)
    grad = _rand((N, 3, 64, 64), 62, 0.01)
    base = diffusion_ref.base_tables("cosine", 1000)
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], [744, 137, 647, 856, 305, 441, 676, 572, 971, 85])
    tb = diffusion_ref.diffusion_tables(nb)
    from autodiffusion_b200.gaussian_diffusion import ddim_coefficients

    for i in (9, 4, 0):
        got = ops.ddim_step(x, mo, grad, ddim_coefficients(tb, i), True)
        want = diffusion_ref.ddim_sample_loop(
            lambda xx, t, **kw: mo.cpu(), x.shape, {k: v[i:i + 1] for k, v in tb.items()}, [tmap[i]], x.cpu(), True,
            cond_fn=lambda xx, t, **kw: grad.cpu())
        assert torch.equal(got.cpu(), want), f"ddim_step step {i} is not bit-exact"
    s = torch.tanh(_rand((N, 3, 64, 64), 63, 1.5)) * 1.05
    assert torch.equal(ops.pack_uint8(s).cpu(), diffusion_ref.pack_uint8(s.cpu()))


def test_moments_batch256_d2048():
    ops = _ops()
    f = _rand((N, 2048), 70) + 0.5
    sx = torch.zeros(2048, dtype=torch.float64, device=DEV)
    sxx = torch.zeros(2048, 2048, dtype=torch.float64, device=DEV)
    ops.moments_accumulate(f, sx, sxx)
    torch.cuda.synchronize()
    f64 = f.double()
    assert (sx - f64.sum(0)).abs().max().item() <= 1e-9
    assert (sxx - f64.t() @ f64).abs().max().item() <= 1e-8


# ---------------------------------------------------------------- the benchmarked candidate at batch 256
def _full_models():
    from tests.test_classifier_gpu import build_classifier
    from tests.util import ADM_FLAGS, build_ours, oracle_weights

    cfg, sd = oracle_weights(ADM_FLAGS)
    model, diffusion = build_ours(ADM_FLAGS, sd)
    clf, _, _ = build_classifier(4, 128)
    return model, diffusion, clf


def test_benchmarked_candidate_at_batch256_reproduces_the_reference_images():
    """bench.py's workload, bit for bit the same plan: cand10 + mask, native classifier guidance, batch 256, one CUDA
    graph. Samples are independent given (noise, label): rows 0-7 carry the reference run's inputs and must match its
    outputs; the same rows sampled at batch 8 must agree with the batch-256 run (nothing leaks across the batch)."""
    from autodiffusion_b200.classifier import ClassifierGuidance
    from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate
    from tests.util import golden, parse_skip_list, psnr

    g = golden("config2_admg64_cand10_guided.npz")
    model, diffusion, clf = _full_models()
    cand = {"timesteps": g["timesteps"].tolist(), "skip_layers": parse_skip_list(g["skip_layers"])}
    active, per_step = resolve_candidate(cand, diffusion)
    assert [active.timestep_map[i] for i in range(active.num_timesteps)][::-1] == g["seen_t"].tolist()
    assert [per_step[i] for i in range(active.num_timesteps)][::-1] == [sorted(s) for s in parse_skip_list(g["seen_skip"])]
    guide = ClassifierGuidance(clf, 1.0)
    plan = SchedulePlan(model, active, per_step, N, cond_fn=guide, pack_uint8=True)
    assert plan.graph is not None
    noise = _rand((N, 3, 64, 64), 80)
    y = torch.randint(0, 1000, (N,), generator=torch.Generator().manual_seed(81)).to(DEV)
    noise[:8] = torch.from_numpy(g["noise"]).to(DEV)
    y[:8] = torch.from_numpy(g["y"]).to(DEV)
    out = plan.run(noise, y).clone()
    u8 = plan.u8.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    ref = torch.from_numpy(g["final"])
    got = out[:8].cpu()
    err = (got - ref).abs().flatten().double().numpy()
    p = psnr(got, ref)
    d8 = np.abs(u8[:8].cpu().numpy().astype(np.int32) - g["uint8"].astype(np.int32))
    print(f"cand10 guided @ batch 256, rows 0-7 vs the reference run: psnr={p:.2f} dB max_abs={err.max():.4g} "
          f"p99={np.percentile(err, 99):.4g} p99.9={np.percentile(err, 99.9):.4g}; uint8 mean |diff| {d8.mean():.3f} LSB, "
          f"within 1 LSB {(d8 <= 1).mean() * 100:.1f}%, max {d8.max()}")
    assert p >= 41.5  # measured 44.5 dB
    assert np.percentile(err, 99) <= 0.08 and np.percentile(err, 99.9) <= 0.22  # measured 0.056 / 0.147 (t = 971: Bm = 22.9)
    assert d8.mean() <= 0.5 and (d8 <= 1).mean() >= 0.91  # measured 0.33 LSB, 94.4 %
    # batch 8 through the same code path
    plan8 = SchedulePlan(model, active, per_step, 8, cond_fn=guide, pack_uint8=True)
    out8 = plan8.run(noise[:8].contiguous(), y[:8].contiguous()).clone().cpu()
    p8 = psnr(out8, ref)
    cross = psnr(out8, got)
    print(f"  batch 8: psnr={p8:.2f} dB vs the reference; batch-256 rows vs batch-8 rows: {cross:.2f} dB")
    assert p8 >= 41.5 and cross >= 55.0  # measured 44.5 / 58.6 dB (fp64 atomics order differs between the two batches)
