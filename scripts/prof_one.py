#!/usr/bin/env python
"""One launch of each kernel worth an `ncu --set full` capture, at the benchmarked shapes (batch from argv, default 64):
attention forward (T=1024, 6 heads), attention backward (T=1024, 4 heads), GroupNorm backward 64x64x128, GroupNorm apply
64x64x192 (FiLM + SiLU), qkv / proj GEMMs at T=1024 and a 3x3 conv at each resolution. Every op runs twice (warm + 1)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["attn", "attn_bwd", "gn_bwd", "gn", "gemm", "conv"]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)

for _ in range(2):
    if "attn" in which:
        t, heads = 1024, 6
        qkv = R(B * t, 3 * heads * 64).bfloat16()
        ops.attention(qkv, B, t, heads, False)
    if "attn_bwd" in which:
        t, heads = 1024, 4
        c = heads * 64
        qkv, dout = R(B * t, 3 * c).bfloat16(), R(B * t, c).bfloat16()
        lse = torch.empty(B * heads, t, device=dev)
        out = ops.attention(qkv, B, t, heads, True, lse=lse)
        ops.attention_backward(qkv, out, dout, lse, B, t, heads, True)
    if "gn_bwd" in which:
        for (r, c) in [(64, 128), (16, 384)]:
            x, dy = R(B, r, r, c).bfloat16(), R(B, r, r, c).bfloat16()
            gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
            ops.groupnorm(x, gamma, beta, stats=stats)
            ops.gn_backward(x, stats, gamma, beta, dy, add=dy, add_mode=ops.RES_SAME)
    if "gn" in which:
        r, c = 64, 192
        x = R(B, r, r, c).bfloat16()
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
        ops.groupnorm(x, gamma, beta, stats=stats)
        ops.groupnorm(x, gamma, beta, stats=stats, stats_ready=True, scale_shift=R(B, 2 * c) * 0.1, ss_stride=2 * c)
    if "gemm" in which:
        t, c = 1024, 384
        x = R(B, 32, 32, c).bfloat16()
        wq = ops.pack_conv_weight([(R(3 * c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
        wp = ops.pack_conv_weight([(R(c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
        ops.conv_igemm([(x, 1)], wq, R(3 * c), 3 * c)
        ops.conv_igemm([(x, 1)], wp, R(c), c, residual=x, res_mode=ops.RES_SAME)
    if "conv" in which:
        for (r, c) in [(64, 192), (32, 384), (16, 576), (8, 768)]:
            x = R(B, r, r, c).bfloat16()
            w = ops.pack_conv_weight([(R(c, c, 3, 3) / math.sqrt(9 * c)).cpu()]).to(dev)
            st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
            ops.conv_igemm([(x, 9)], w, R(c), c, stats_out=st)
    if "convset" in which and _ == 0:
        # the distinct 3x3 convolutions of ADM-G 64 (Cin -> Cout @ resolution), the classifier's 128-wide
        # ones, and the qkv / proj GEMMs at the three attention resolutions: what `roofline.traffic` is averaged over
        convs = [(64, 192, 192), (64, 384, 192), (64, 576, 192), (64, 128, 128), (32, 192, 384), (32, 384, 384), (32, 576, 384),
                 (32, 768, 384), (32, 960, 384), (32, 576, 576), (32, 256, 256), (16, 384, 576), (16, 576, 576), (16, 960, 576),
                 (16, 1152, 576), (16, 1344, 576), (16, 768, 768), (8, 576, 768), (8, 768, 768), (8, 1344, 768), (8, 1536, 768)]
        for (r, cin, cout) in convs:
            x = R(B, r, r, cin).bfloat16()
            w = ops.pack_conv_weight([(R(cout, cin, 3, 3) / math.sqrt(9 * cin)).cpu()]).to(dev)
            st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
            ops.conv_igemm([(x, 9)], w, R(cout), cout, stats_out=st)
            del x, w
        for (r, c) in [(32, 384), (16, 576), (8, 768)]:
            x = R(B, r, r, c).bfloat16()
            wq = ops.pack_conv_weight([(R(3 * c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
            wp = ops.pack_conv_weight([(R(c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
            ops.conv_igemm([(x, 1)], wq, R(3 * c), 3 * c)
            ops.conv_igemm([(x, 1)], wp, R(c), c, residual=x, res_mode=ops.RES_SAME)
torch.cuda.synchronize()
print("ok")
