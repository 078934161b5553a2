#!/usr/bin/env python
"""One warm + one measured pass of the native Inception-V3 pool_3 extractor on 64 uint8 images (for an ncu launch list)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200.inception import InceptionPool3  # noqa: E402

m = InceptionPool3().cuda()
u8 = torch.randint(0, 256, (64, 64, 64, 3), dtype=torch.uint8, device="cuda")
f = m(u8)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measured")
f = m(u8)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", tuple(f.shape), bool(torch.isfinite(f).all()))
