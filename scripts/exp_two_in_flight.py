#!/usr/bin/env python
"""Experiment: two batches of the guided 10-step candidate in flight on two streams (two model / classifier instances with
their own buffers and plans) vs one. Prints images/s for both. argv[1] = batch (default 256)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autodiffusion_b200 import classifier_defaults, create_classifier, create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402
from autodiffusion_b200.classifier import ClassifierGuidance  # noqa: E402
from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)


def instance():
    d = model_and_diffusion_defaults()
    d.update(bench.ADM_FLAGS)
    model, diffusion = create_model_and_diffusion(**d)
    model.load_state_dict(bench.bench_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model.to(dev).eval()
    model.convert_to_fp16()
    cd = classifier_defaults()
    cd.update(bench.CLASSIFIER)
    clf = create_classifier(**cd)
    clf.load_state_dict(bench.bench_weights({k: tuple(v.shape) for k, v in clf.state_dict().items()}, seed=1))
    clf.to(dev).eval()
    active, per_step = resolve_candidate(bench.CAND10, diffusion)
    return SchedulePlan(model, active, per_step, B, clip_denoised=True, cond_fn=ClassifierGuidance(clf, 1.0), pack_uint8=True)


pa, pb = instance(), instance()
g = torch.Generator(device=dev).manual_seed(1)
noise = torch.randn(pa.shape, generator=g, device=dev)
y = torch.randint(0, 1000, (B,), generator=g, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one(n):
    for _ in range(n):
        pa.run(noise, y)


def two(n):
    for _ in range(n // 2):
        with torch.cuda.stream(s1):
            pa.run(noise, y)
        with torch.cuda.stream(s2):
            pb.run(noise, y)


for name, fn in (("one in flight", one), ("two in flight", two), ("one in flight", one), ("two in flight", two)):
    fn(2)
    torch.cuda.synchronize()
    t0 = time.time()
    fn(6)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{name}: {6 * B / dt:.1f} images/s ({dt / 6 * 1e3:.1f} ms per batch)")
ra = pa.run(noise, y).clone()
rb = pb.run(noise, y).clone()
torch.cuda.synchronize()
print("instances agree:", torch.equal(ra, rb))
