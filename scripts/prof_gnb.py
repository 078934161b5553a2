#!/usr/bin/env python
"""A/B of the GroupNorm-backward sums fused into the data-gradient conv epilogue (conv_igemm gnb=...), at the guidance
classifier's shapes (batch from argv, default 256): conv alone, conv + sums, two-pass gn_backward, one-pass gn_backward."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)


def timed(fn, reps=8):
    fn()
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# (resolution, K-source channels, output channels, taps, silu, film)
shapes = [(64, 128, 128, 9, True, True), (32, 256, 256, 9, True, True), (32, 256, 128, 9, True, False),
          (16, 384, 384, 9, True, True), (8, 512, 512, 9, True, True), (32, 768, 256, 1, False, False),
          (16, 1152, 384, 1, False, False), (8, 1536, 512, 1, False, False)]
tot = [0.0] * 4
for (r, cin, cout, taps, silu, film) in shapes:
    x = R(B, r, r, cout).bfloat16()
    dy = R(B, r, r, cin).bfloat16()
    k = 3 if taps == 9 else 1
    w = ops.pack_conv_weight([(R(cout, cin, k, k) / math.sqrt(taps * cin)).cpu()]).to(dev)
    gamma, beta = 1 + 0.1 * R(cout), 0.1 * R(cout)
    kw = dict(silu=silu)
    if film:
        kw.update(scale_shift=0.2 * R(B, 2 * cout), ss_stride=2 * cout)
    stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
    ops.groupnorm(x, gamma, beta, stats=stats, **kw)
    bst = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
    out = torch.empty(B, r, r, cout, dtype=torch.bfloat16, device=dev)
    dx = torch.empty_like(out)
    fz = dict(x=x, stats=stats, gamma=gamma, beta=beta, bstats=bst, **kw)
    t_conv = timed(lambda: ops.conv_igemm([(dy, taps)], w, None, cout, out=out))
    t_fused = timed(lambda: ops.conv_igemm([(dy, taps)], w, None, cout, out=out, gnb=fz))
    t_gn2 = timed(lambda: ops.gn_backward(x, stats, gamma, beta, out, dx=dx, bstats=bst, **kw))
    ops.conv_igemm([(dy, taps)], w, None, cout, out=out, gnb=fz)
    t_gn1 = timed(lambda: ops.gn_backward(x, stats, gamma, beta, out, dx=dx, bstats=bst, bstats_ready=True, **kw))
    fl = 2.0 * B * r * r * cout * cin * taps
    print(f"res {r:2d} K {taps}x{cin:4d} -> {cout:3d}: conv {t_conv:.3f} ms ({fl / t_conv / 1e9:.0f} TF/s)  conv+sums {t_fused:.3f} ms "
          f"(+{t_fused - t_conv:.3f})  gn_bwd 2-pass {t_gn2:.3f}  1-pass {t_gn1:.3f} (-{t_gn2 - t_gn1:.3f})  net {t_fused - t_conv - (t_gn2 - t_gn1):+.3f} ms")
    for i, v in enumerate((t_conv, t_fused, t_gn2, t_gn1)):
        tot[i] += v
    del x, dy, w, out, dx
print("totals: conv %.3f conv+sums %.3f gn2 %.3f gn1 %.3f" % tuple(tot))
