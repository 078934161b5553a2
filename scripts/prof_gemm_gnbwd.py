#!/usr/bin/env python
"""qkv / proj_out GEMMs of the three attention resolutions and the one-pass GroupNorm backward (sums from the producing
conv), at batch argv[1] (default 256). argv[2] = "time" (CUDA-event averages, default) or "once" (one warm + one launch of
each, for an ncu capture)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "time"
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)
HBM = 6441.0


def run(fn, reps=10):
    fn()
    if mode == "once":
        fn()
        torch.cuda.synchronize()
        return float("nan")
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (r, c) in [(32, 384), (16, 576), (8, 768)]:
    x = R(B, r, r, c).bfloat16()
    res = R(B, r, r, c).bfloat16()
    wq = ops.pack_conv_weight([(R(3 * c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
    wp = ops.pack_conv_weight([(R(c, c, 1) / math.sqrt(c)).cpu()]).to(dev)
    bq, bp = R(3 * c), R(c)
    oq = torch.empty(B, r, r, 3 * c, dtype=torch.bfloat16, device=dev)
    op = torch.empty(B, r, r, c, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
    M = B * r * r
    t = run(lambda: ops.conv_igemm([(x, 1)], wq, bq, 3 * c, out=oq))
    by = M * c * 2 * 4
    print(f"qkv  T={r * r:4d} C={c}: {t * 1e3:7.1f} us  {2.0 * M * c * 3 * c / t / 1e9:6.0f} TFLOP/s  {by / t / 1e6:5.0f} GB/s algorithmic")
    t = run(lambda: ops.conv_igemm([(x, 1)], wp, bp, c, out=op, residual=res, res_mode=ops.RES_SAME, stats_out=st))
    by = M * c * 2 * 3
    print(f"proj T={r * r:4d} C={c}: {t * 1e3:7.1f} us  {2.0 * M * c * c / t / 1e9:6.0f} TFLOP/s  {by / t / 1e6:5.0f} GB/s algorithmic "
          f"= {by / t / 1e6 / HBM:.2f} of the {HBM:.0f} GB/s copy peak (residual + GroupNorm sums in the epilogue)")
    del x, res, oq, op

for (r, c) in [(64, 128), (32, 256), (16, 384)]:
    x, dy = R(B, r, r, c).bfloat16(), R(B, r, r, c).bfloat16()
    gamma, beta = 1 + 0.1 * R(c), 0.1 * R(c)
    stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
    ops.groupnorm(x, gamma, beta, stats=stats)
    bst = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    n = x.numel()
    t2 = run(lambda: ops.gn_backward(x, stats, gamma, beta, dy, add=dy, add_mode=ops.RES_SAME, dx=dx, bstats=bst))
    t1 = run(lambda: ops.gn_backward(x, stats, gamma, beta, dy, add=dy, add_mode=ops.RES_SAME, dx=dx, bstats=bst, bstats_ready=True))
    print(f"gn_bwd {r}x{r}x{c}: two-pass {t2 * 1e3:7.1f} us ({n * 8 / t2 / 1e6:5.0f} GB/s of its algorithmic 8 B/element)  "
          f"one-pass {t1 * 1e3:7.1f} us ({n * 8 / t1 / 1e6:5.0f} GB/s = {n * 8 / t1 / 1e6 / HBM:.2f} of the copy peak)")
    del x, dy, dx
print("ok")
