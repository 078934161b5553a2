"""GPU parity of the classifier-guidance backward ops against torch autograd over the CPU oracle.

Each reference gradient is produced by `torch.autograd.grad` through the oracle's fp32 restatement
(oracle/unet_ref.py) of the reference op on bf16-rounded inputs - the same computation the
reference's cond_fn differentiates (…progressive.py:383-390). Tolerances: gradients are stored in
bf16 (2^-9 relative rounding) after bf16 tensor-core products, so max-abs error <= 2^-6 of the
reference's max magnitude (attention: 2^-5, three chained bf16 products).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import unet_ref

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
    rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-30)).item()
    print(f"{what}: max_abs_err={err:.4g} ref_max={scale:.4g} rel={err / max(scale, 1e-30):.4g} rel_rms={rms:.4g}")
    assert math.isfinite(err), what
    assert err <= rel * scale + 1e-6, f"{what}: err {err} > {rel} * {scale}"


@pytest.mark.parametrize("n,res,c", [(2, 64, 128), (2, 16, 384), (3, 8, 512), (2, 32, 256)])
@pytest.mark.parametrize("variant", ["silu", "plain", "film", "down", "silu_add", "down_addpool"])
def test_gn_backward(n, res, c, variant):
    ops = _ops()
    x = (_bf(_rand((n, c, res, res), 40, 2.0) + 0.5)).requires_grad_(True)
    g = 1 + 0.1 * _rand((c,), 41)
    bt = 0.1 * _rand((c,), 42)
    hn = unet_ref.group_norm32(x, g, bt)
    kw, add, add_ref = dict(silu=True), None, None
    ro = res // 2 if variant.startswith("down") else res
    if variant == "plain":
        y, kw = hn, dict(silu=False)
    elif variant in ("silu", "silu_add"):
        y = F.silu(hn)
    elif variant == "film":
        ss = 0.3 * _rand((n, 2 * c + 3), 43)
        y = F.silu(hn * (1 + ss[:, :c, None, None]) + ss[:, c:2 * c, None, None])
        kw = dict(silu=True, scale_shift=ss.to(DEV), ss_stride=2 * c + 3)
    else:
        y = F.avg_pool2d(F.silu(hn), 2, 2)
        kw = dict(silu=True, resample=ops.RESAMPLE_AVGPOOL2)
    dy = _bf(_rand((n, c, ro, ro), 44))
    loss = (y * dy).sum()
    if variant == "silu_add":       # identity skip: out = f(x) + x  ->  dx += dy
        add_ref = _bf(_rand((n, c, res, res), 45))
        loss = loss + (x * add_ref).sum()
        kw.update(add=_nhwc(add_ref), add_mode=ops.RES_SAME)
    elif variant == "down_addpool":  # down block: out = f(x) + avg_pool(x)  ->  dx += up(dres) / 4
        add_ref = _bf(_rand((n, c, ro, ro), 46))
        loss = loss + (F.avg_pool2d(x, 2, 2) * add_ref).sum()
        kw.update(add=_nhwc(add_ref), add_mode=ops.RES_AVGPOOL2)
    ref = torch.autograd.grad(loss, x)[0]
    xd = _nhwc(x.detach())
    stats = torch.empty((n, 32, 2), dtype=torch.float64, device=DEV)
    fkw = {k: v for k, v in kw.items() if k in ("silu", "scale_shift", "ss_stride", "resample")}
    ops.groupnorm(xd, g.to(DEV), bt.to(DEV), stats=stats, **fkw)  # fills the forward sums
    dx = ops.gn_backward(xd, stats, g.to(DEV), bt.to(DEV), _nhwc(dy), **kw)
    torch.cuda.synchronize()
    _check(_nchw(dx), ref, 2 ** -6, f"gn_backward {variant} n{n} r{res} c{c}")


@pytest.mark.parametrize("n,res,cin,cout,taps,variant", [
    (4, 16, 128, 256, 9, "film"),     # two-CTA 256-wide tile or its 128-wide fallback
    (2, 8, 512, 512, 9, "silu"),      # 64-pixel images: two images per 128-row tile
    (2, 32, 192, 192, 1, "plain"),    # attention block: 1x1 data gradient, no SiLU
    (3, 32, 256, 128, 9, "silu"),     # odd image count (n*h*w % 128 == 0 still)
    (8, 16, 384, 384, 9, "film"),
])
def test_conv_fused_gn_backward_sums(n, res, cin, cout, taps, variant):
    """conv_igemm(gnb=...): the data-gradient conv's epilogue reduces sum(dxh) and sum(dxh * xh) of the consumer
    GroupNorm's backward; they must equal the sums gn_backward's own first pass computes from the stored bf16 gradient,
    and dx from the single-pass form (bstats_ready) must match the two-pass one."""
    ops = _ops()
    assert ops.conv_gnb_supported(n, res, res, cout)
    x = _nhwc(_bf(_rand((n, cout, res, res), 60, 2.0) + 0.5))
    g = (1 + 0.1 * _rand((cout,), 61)).to(DEV)
    bt = (0.1 * _rand((cout,), 62)).to(DEV)
    kw = dict(silu=variant != "plain")
    if variant == "film":
        kw.update(scale_shift=(0.3 * _rand((n, 2 * cout + 8), 63)).to(DEV), ss_stride=2 * cout + 8)
    stats = torch.empty((n, 32, 2), dtype=torch.float64, device=DEV)
    ops.groupnorm(x, g, bt, stats=stats, **kw)
    k = 3 if taps == 9 else 1
    w = _rand((cout, cin, k, k), 64, 1.0 / math.sqrt(cin * taps))
    wp = ops.pack_conv_weight([w], DEV)
    dy_in = _nhwc(_bf(_rand((n, cin, res, res), 65)))
    # reference: plain conv, then the two-pass backward (its first pass leaves the sums in bstats)
    dg_ref = ops.conv_igemm([(dy_in, taps)], wp, None, cout)
    b_ref = torch.empty((n, 32, 2), dtype=torch.float64, device=DEV)
    dx_ref = ops.gn_backward(x, stats, g, bt, dg_ref, bstats=b_ref, **kw)
    # fused
    b_f = torch.full((n, 32, 2), 123.0, dtype=torch.float64, device=DEV)  # the conv zeroes it itself
    dg = ops.conv_igemm([(dy_in, taps)], wp, None, cout, gnb=dict(x=x, stats=stats, gamma=g, beta=bt, bstats=b_f, **kw))
    dx = ops.gn_backward(x, stats, g, bt, dg, bstats=b_f, bstats_ready=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(dg, dg_ref), "gnb must not change the conv's output"
    scale = b_ref.abs().max().item()
    err = (b_f - b_ref).abs().max().item()
    assert err <= 2e-5 * max(scale, 1.0), f"fused GroupNorm-backward sums off by {err} (scale {scale})"
    d = (dx.float() - dx_ref.float()).abs().max().item()
    assert d <= 2 ** -7 * dx_ref.float().abs().max().item(), f"dx differs by {d}"
    # every argument check the header states
    with pytest.raises(Exception):
        ops.conv_igemm([(dy_in, taps)], wp, None, cout, residual=x, res_mode=ops.RES_SAME,
                       gnb=dict(x=x, stats=stats, gamma=g, beta=bt, bstats=b_f, **kw))


def test_conv_gnb_supported_shapes():
    ops = _ops()
    assert not ops.conv_gnb_supported(1, 8, 8, 512)      # 64 rows: not a full tile
    assert not ops.conv_gnb_supported(2, 4, 4, 512)      # 16 pixels per image
    assert not ops.conv_gnb_supported(4, 16, 16, 48)     # cout % 32
    assert ops.conv_gnb_supported(256, 64, 64, 128)
    assert ops.conv_gnb_supported(256, 8, 8, 512)


@pytest.mark.parametrize("b,t,heads", [(2, 64, 8), (2, 256, 6), (1, 1024, 4), (3, 64, 2), (2, 128, 3)])
@pytest.mark.parametrize("legacy", [True, False])
@pytest.mark.parametrize("fused", [False, True])
def test_attention_backward(b, t, heads, legacy, fused):
    """fused = the opt-in single-pass kernel (attention_bwd_fused.cu; t % 128 == 0), else the deterministic two-kernel form."""
    ops = _ops()
    prev = ops.set_attention_backward_fused(None)
    ops.set_attention_backward_fused(fused)
    try:
        _attention_backward_case(ops, b, t, heads, legacy, fused)
    finally:
        ops.set_attention_backward_fused(prev)


def _attention_backward_case(ops, b, t, heads, legacy, fused):
    c = heads * 64
    qkv = _bf(_rand((b, 3 * c, t), 50 + t + heads, 1.2)).requires_grad_(True)
    out_ref = unet_ref.qkv_attention(qkv, heads, new_order=not legacy)  # [b, c, t]
    dout = _bf(_rand((b, c, t), 51))
    ref = torch.autograd.grad((out_ref * dout).sum(), qkv)[0]  # [b, 3c, t]
    rows = qkv.detach().permute(0, 2, 1).contiguous().to(torch.bfloat16).to(DEV).view(b * t, 3 * c)
    drows = dout.permute(0, 2, 1).contiguous().to(torch.bfloat16).to(DEV).view(b * t, c)
    lse = torch.empty((b * heads, t), dtype=torch.float32, device=DEV)
    out = ops.attention(rows, b, t, heads, legacy, lse=lse)
    dqkv = ops.attention_backward(rows, out, drows, lse, b, t, heads, legacy)
    torch.cuda.synchronize()
    got_out = out.float().view(b, t, c).permute(0, 2, 1).cpu()
    _check(got_out, out_ref.detach(), 2 ** -6, f"attention(lse) fwd b{b} t{t} h{heads} legacy={legacy}")
    got = dqkv.float().view(b, t, 3 * c).permute(0, 2, 1).cpu()
    _check(got, ref, 2 ** -5, f"attention_backward b{b} t{t} h{heads} legacy={legacy} fused={fused}")
    if not fused:  # the default form is bit-reproducible
        again = ops.attention_backward(rows, out, drows, lse, b, t, heads, legacy)
        assert torch.equal(again, dqkv)


def test_conv_dgrad_weights():
    """dx of a 3x3 / 1x1 conv = the same implicit GEMM over dy with transposed, flipped weights."""
    ops = _ops()
    n, r, cin, cout = 2, 16, 256, 384
    x = _bf(_rand((n, cin, r, r), 60)).requires_grad_(True)
    for k in (3, 1):
        w = _bf(_rand((cout, cin, k, k), 61, (cin * k * k) ** -0.5))
        dy = _bf(_rand((n, cout, r, r), 62))
        ref = torch.autograd.grad((F.conv2d(x, w, padding=k // 2) * dy).sum(), x)[0]
        wt = ops.pack_conv_weight_dgrad(w, DEV)
        dx = ops.conv_igemm([(_nhwc(dy), k * k)], wt, None, cin)
        torch.cuda.synchronize()
        _check(_nchw(dx), ref, 2 ** -7, f"conv dgrad k{k}")


def test_attention_pool_forward_backward():
    """AttentionPool2d (unet.py:22-51) token 0 and its input gradient, pieced together as classifier.py does."""
    ops = _ops()
    n, hw, c, k = 3, 8, 512, 1000
    P = hw * hw
    h = _bf(_rand((n, c, hw, hw), 70)).requires_grad_(True)
    pos = _rand((c, P + 1), 71, c ** -0.5)
    wq = _bf(_rand((3 * c, c, 1), 72, c ** -0.5))
    bq = 0.1 * _rand((3 * c,), 73)
    wc = _rand((k, c, 1), 74, c ** -0.5)
    bc = 0.1 * _rand((k,), 75)
    sd = {"out.2.positional_embedding": pos, "out.2.qkv_proj.weight": wq, "out.2.qkv_proj.bias": bq,
          "out.2.c_proj.weight": wc, "out.2.c_proj.bias": bc}
    hh = h.reshape(n, c, -1)
    hh = torch.cat([hh.mean(dim=-1, keepdim=True), hh], dim=-1) + pos[None]
    t = F.conv1d(hh, wq, bq)
    t = unet_ref.qkv_attention(t, c // 64, new_order=True)
    logits_ref = F.conv1d(t, wc, bc)[:, :, 0]
    y = torch.tensor([3, 999, 0])
    sel = F.log_softmax(logits_ref, dim=-1)[range(n), y]
    ref_dh = torch.autograd.grad(sel.sum() * 2.5, h)[0]

    hd = _nhwc(h.detach())
    xp, mean = ops.pool_prepare(hd, pos.to(DEV))
    w2 = wq[:, :, 0]
    wkv = ops.pack_conv_weight([wq[c:]], DEV)
    kv = ops.conv_igemm([(xp, 1)], wkv, bq[c:].to(DEV), 2 * c)
    qkv0 = ops.linear(mean, w2.to(DEV).contiguous(), bq.to(DEV))
    out0, probs = ops.pool_attention(qkv0, kv)
    logits = ops.linear(out0, wc[:, :, 0].to(DEV).contiguous(), bc.to(DEV))
    torch.cuda.synchronize()
    _check(logits.cpu(), logits_ref.detach(), 2 ** -7, "attention pool logits")
    dlog = ops.logsoftmax_grad(logits, y.to(DEV), 2.5)
    lr = logits.cpu().requires_grad_(True)
    ref_dlog = torch.autograd.grad(F.log_softmax(lr, dim=-1)[range(n), y].sum() * 2.5, lr)[0]
    _check(dlog.cpu(), ref_dlog, 1e-5, "logsoftmax grad")
    dout0 = ops.linear(dlog, wc[:, :, 0].t().contiguous().to(DEV), None)
    dqkv0, dkv = ops.pool_attention_backward(dout0, probs, qkv0, kv)
    dmean = ops.linear(dqkv0, w2.t().contiguous().to(DEV), None)
    dxp = ops.conv_igemm([(dkv, 1)], ops.pack_conv_weight_dgrad(wq[c:], DEV), None, c)
    dh = ops.pool_merge(dxp, dmean)
    torch.cuda.synchronize()
    _check(_nchw(dh), ref_dh, 2 ** -6, "attention pool input gradient")
