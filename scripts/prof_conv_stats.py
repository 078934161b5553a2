#!/usr/bin/env python
"""What the fused GroupNorm statistics cost the producing conv: each 3x3 shape with and without stats_out (batch argv[1])."""
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
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (r, cin, cout) in [(64, 192, 192), (64, 384, 192), (32, 384, 384), (32, 768, 384), (16, 576, 576), (8, 768, 768), (64, 128, 128)]:
    x = R(B, r, r, cin).bfloat16()
    w = ops.pack_conv_weight([(R(cout, cin, 3, 3) / math.sqrt(9 * cin)).cpu()]).to(dev)
    b = R(cout)
    out = torch.empty(B, r, r, cout, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
    fl = 2.0 * B * r * r * cout * cin * 9
    t0 = timed(lambda: ops.conv_igemm([(x, 9)], w, b, cout, out=out))
    t1 = timed(lambda: ops.conv_igemm([(x, 9)], w, b, cout, out=out, stats_out=st))
    print(f"res {r:2d} {cin:4d}->{cout:3d}: plain {t0:.3f} ms ({fl / t0 / 1e9:.0f} TF/s)  with stats {t1:.3f} ms ({fl / t1 / 1e9:.0f} TF/s)  +{100 * (t1 / t0 - 1):.1f} %")
    del x, w, out
