#!/usr/bin/env python
"""Two launches for an ncu capture: the 64x64 128->128 data-gradient conv of the classifier without and with the fused
GroupNorm-backward sums (argv: batch, default 256)."""
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
r, cin, cout, taps = 64, 128, 128, 9
x, dy = R(B, r, r, cout).bfloat16(), R(B, r, r, cin).bfloat16()
w = ops.pack_conv_weight([(R(cout, cin, 3, 3) / math.sqrt(taps * cin)).cpu()]).to(dev)
gamma, beta = 1 + 0.1 * R(cout), 0.1 * R(cout)
kw = dict(silu=True, scale_shift=0.2 * R(B, 2 * cout), ss_stride=2 * cout)
stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
ops.groupnorm(x, gamma, beta, stats=stats, **kw)
bst = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
out = torch.empty(B, r, r, cout, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    ops.conv_igemm([(dy, taps)], w, None, cout, out=out)
    ops.conv_igemm([(dy, taps)], w, None, cout, out=out, gnb=dict(x=x, stats=stats, gamma=gamma, beta=beta, bstats=bst, **kw))
torch.cuda.synchronize()
print("ok")
