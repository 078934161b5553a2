"""GPU parity of the native Inception-V3 pool_3 extractor against the same graph in fp32 torch on
identical seeded weights (oracle/inception_ref.py: torchvision's inception_v3 + pytorch-fid's three pooling changes, the
graph both references evaluate - evaluator_v1.py:252-280, search_ea.py:95-127). No Inception weights exist offline, so
parity is pinned at the activation level: He-initialised convolutions, randomised BatchNorm statistics and affines.
Bars: the data-movement kernels exact up to the bf16 rounding of their output; pool_3 features relative RMS <= 0.4 %
(measured 0.12 %; VERDICT r1 asked for <= 1 %)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from autodiffusion_b200 import ops

    return ops


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def test_resize_bilinear_matches_interpolate():
    ops = _ops()
    u8 = torch.randint(0, 256, (3, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    out = ops.resize_bilinear_u8(u8.to(DEV), 299, 299).float().cpu()
    ref = F.interpolate(u8.permute(0, 3, 1, 2).float(), size=(299, 299), mode="bilinear", align_corners=False) / 127.5 - 1.0
    assert out.shape == (3, 299, 299, 8) and (out[..., 3:] == 0).all()
    err = (out[..., :3].permute(0, 3, 1, 2) - ref).abs().max().item()
    print(f"resize 64 -> 299: max abs err {err:.3g} (bf16 output)")
    assert err <= 2 ** -8


@pytest.mark.parametrize("kh,kw,stride,ph,pw,h,w", [(3, 3, 2, 0, 0, 19, 19), (3, 3, 1, 0, 0, 9, 11), (5, 5, 1, 2, 2, 7, 7), (1, 7, 1, 0, 3, 6, 9),
                                                   (7, 1, 1, 3, 0, 9, 6), (1, 1, 1, 0, 0, 5, 5), (3, 3, 2, 0, 0, 17, 17)])
def test_gather_patches_matches_unfold(kh, kw, stride, ph, pw, h, w):
    """The patch gather = F.unfold over the ReLU'd channel concatenation, K ordered (tap, channel)."""
    ops = _ops()
    a, b, c = _rand((2, 16, h, w), 1), _rand((2, 8, h, w), 2), _rand((2, 24, h, w), 3)
    srcs = [(_nhwc(a), True), (_nhwc(b), False), (_nhwc(c), True)]
    out = ops.gather_patches(srcs, kh, kw, stride, ph, pw).float().cpu()
    cat = torch.cat([F.relu(a.bfloat16().float()), b.bfloat16().float(), F.relu(c.bfloat16().float())], 1)
    ctot = cat.shape[1]
    ho, wo = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
    un = F.unfold(cat, (kh, kw), padding=(ph, pw), stride=stride)  # [n, ctot*kh*kw, L], channel-major
    un = un.view(2, ctot, kh * kw, ho * wo).permute(0, 3, 2, 1).reshape(2 * ho * wo, kh * kw * ctot)
    assert out.shape == (2 * ho * wo, 1, 1, kh * kw * ctot)
    assert torch.equal(out.view(2 * ho * wo, -1), un)


@pytest.mark.parametrize("mode,stride,pad", [(0, 2, 0), (0, 1, 1), (1, 1, 1), (2, 1, 1)])
def test_pool3x3_modes(mode, stride, pad):
    ops = _ops()
    a, b = _rand((2, 16, 9, 9), 4), _rand((2, 8, 9, 9), 5)
    out = ops.pool3x3([(_nhwc(a), True), (_nhwc(b), True)], stride, pad, mode).float().cpu().permute(0, 3, 1, 2)
    cat = F.relu(torch.cat([a.bfloat16().float(), b.bfloat16().float()], 1))
    if mode == 0:
        ref = F.max_pool2d(cat, 3, stride, pad)
    else:
        ref = F.avg_pool2d(cat, 3, stride, pad, count_include_pad=(mode == 1))
    err = (out - ref).abs().max().item()
    assert out.shape == ref.shape and err <= 2 ** -8 * max(1.0, ref.abs().max().item())


def test_global_avgpool():
    ops = _ops()
    a, b = _rand((3, 16, 8, 8), 6), _rand((3, 24, 8, 8), 7)
    out = ops.global_avgpool([(_nhwc(a), True), (_nhwc(b), False)]).cpu()
    ref = torch.cat([F.relu(a.bfloat16().float()), b.bfloat16().float()], 1).mean((2, 3))
    assert (out - ref).abs().max().item() <= 1e-5


def _randomize_bn(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(1.0 + 0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))


@pytest.mark.parametrize("fid_variant", [True, False])
def test_pool3_features_match_the_fp32_graph(fid_variant):
    from autodiffusion_b200.inception import InceptionPool3
    from oracle.inception_ref import InceptionPool3Ref

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = InceptionPool3(seed=0, fid_variant=fid_variant)
        _randomize_bn(net, 1)
        ref = InceptionPool3Ref(seed=5, half=False, fid_variant=fid_variant)
        ref.net.load_state_dict(net.state_dict(), strict=True)  # identical weights, torchvision's names
        net.to(DEV)
        ref.to(DEV)
        u8 = torch.randint(0, 256, (5, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2)).to(DEV)
        # smooth images too (the samplers' outputs are not white noise)
        yy, xx = torch.meshgrid(torch.arange(64), torch.arange(64), indexing="ij")
        u8[0] = torch.stack([(yy * 4) % 256, (xx * 4) % 256, ((xx + yy) * 2) % 256], -1).to(torch.uint8).to(DEV)
        got = net(u8)
        want = ref(u8)
        torch.cuda.synchronize()
        assert got.shape == (5, 2048) and got.dtype == torch.float32 and torch.isfinite(got).all()
        rel = ((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item()
        mx = (got - want).abs().max().item()
        print(f"pool_3 (fid_variant={fid_variant}): rel_rms={rel:.4g} max_abs={mx:.4g} ref_rms={want.pow(2).mean().sqrt().item():.4g}; "
              f"kernels launched: {net.gpu_launches}")
        assert rel <= 0.004  # measured 0.11-0.12 % (bf16 activations, hi+lo split weights)
        again = net(u8)
        assert torch.equal(again, got)  # replayed graph, deterministic
        if fid_variant:  # the two graphs really differ (pool divisors, Mixed_7c max pool)
            other = InceptionPool3(seed=0, fid_variant=False)
            _randomize_bn(other, 1)
            assert not torch.allclose(other.to(DEV)(u8), got, rtol=1e-3, atol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_extractor_feeds_the_moment_kernel_in_chunks():
    """A 70-image batch runs as a 64-image and a 6-image recorded pass; the features drive MomentAccumulator directly."""
    from autodiffusion_b200.evaluator import MomentAccumulator
    from autodiffusion_b200.inception import InceptionPool3

    net = InceptionPool3(seed=0).to(DEV)
    u8 = torch.randint(0, 256, (70, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).to(DEV)
    f = net(u8)
    assert f.shape == (70, 2048)
    assert torch.allclose(net(u8[64:].contiguous()), f[64:], rtol=1e-5, atol=1e-6)  # batch-independent
    acc = MomentAccumulator(2048, DEV)
    acc.add(f)
    mu, sigma = acc.statistics()
    assert abs(mu.mean() - f.double().mean().item()) <= 1e-9
