#!/usr/bin/env python
"""Stand-alone launches of the classifier-guidance kernels at their largest ADM-G 64 shapes (batch 64), for
`ncu --set full` captures and CUDA-event timing: attention forward / backward at T=1024 (4 heads), GroupNorm
backward at 64x64x128. Prints per-op milliseconds (CUDA events, 10 runs after 3 warm-ups)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (t, heads) in [(1024, 4), (1024, 6), (256, 6), (64, 8)]:
    c = heads * 64
    qkv = torch.randn(B * t, 3 * c, device=dev, generator=g).bfloat16()
    dout = torch.randn(B * t, c, device=dev, generator=g).bfloat16()
    lse = torch.empty(B * heads, t, device=dev)
    out = ops.attention(qkv, B, t, heads, True, lse=lse)
    dqkv = torch.empty_like(qkv)
    dsum = torch.empty(B * heads, t, device=dev)
    ms_f = timeit(lambda: ops.attention(qkv, B, t, heads, True, out=out, lse=lse))
    ms_b = timeit(lambda: ops.attention_backward(qkv, out, dout, lse, B, t, heads, True, dqkv=dqkv, dsum=dsum))
    fl = 4.0 * B * heads * t * t * 64
    print(f"attention b{B} t{t} h{heads}: fwd {ms_f:.3f} ms ({fl / ms_f / 1e9:.0f} TF/s)  bwd {ms_b:.3f} ms "
          f"({2 * fl / ms_b / 1e9:.0f} TF/s algorithmic, {3.5 * fl / ms_b / 1e9:.0f} executed)")

for (r, c) in [(64, 128), (32, 256), (16, 384), (8, 512)]:
    x = torch.randn(B, r, r, c, device=dev, generator=g).bfloat16()
    dy = torch.randn(B, r, r, c, device=dev, generator=g).bfloat16()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
    ops.groupnorm(x, gamma, beta, stats=stats)
    dx = torch.empty_like(x)
    bst = torch.empty_like(stats)
    ms = timeit(lambda: ops.gn_backward(x, stats, gamma, beta, dy, add=dy, add_mode=ops.RES_SAME, dx=dx, bstats=bst))
    by = x.numel() * 2 * 4
    print(f"gn_backward b{B} {r}x{r}x{c}: {ms:.3f} ms ({by / ms / 1e6:.0f} GB/s algorithmic: x, dout, add read once, dx written)")

for (r, c, film) in [(64, 192, True), (32, 384, True), (16, 576, False), (8, 768, False)]:
    x = torch.randn(B, r, r, c, device=dev, generator=g).bfloat16()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    ops.groupnorm(x, gamma, beta, stats=stats)
    ss = torch.randn(B, 2 * c, device=dev, generator=g) * 0.1 if film else None
    kw = dict(scale_shift=ss, ss_stride=2 * c) if film else {}
    ms = timeit(lambda: ops.groupnorm(x, gamma, beta, out=y, stats=stats, stats_ready=True, **kw))
    print(f"groupnorm_apply b{B} {r}x{r}x{c} film={film}: {ms:.3f} ms ({x.numel() * 4 / ms / 1e6:.0f} GB/s)")
