#!/usr/bin/env python
"""Where does building the plan of a NEW skip mask go? (record_forward / first run / per-mask graph capture /
whole-candidate graph capture). Batch 250, full ADM-G 64 + depth-4 classifier."""
import os, sys, time, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults, classifier_defaults, create_classifier, ops
from autodiffusion_b200.classifier import ClassifierGuidance
from autodiffusion_b200.dynamic_unet import _UNetPlan
from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate
from bench import ADM_FLAGS, bench_weights, CAND10

d = model_and_diffusion_defaults(); d.update(ADM_FLAGS)
model, diffusion = create_model_and_diffusion(**d)
model.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}))
model.cuda().eval(); model.convert_to_fp16()
cd = classifier_defaults(); cd.update(classifier_depth=4)
clf = create_classifier(**cd); clf.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in clf.state_dict().items()}, seed=1)); clf.cuda().eval()
B = 250
def sync(): torch.cuda.synchronize()
model.get_plan(B, 64, 64, [])  # warm: packs weights, first plan
sync()
for mask in ([3, 9, 20], [1, 2, 30, 41, 50, 7, 12, 33, 25]):
    t0 = time.time(); up = _UNetPlan(model, B, 64, 64, tuple(sorted(mask))); sync(); t1 = time.time()
    up.launches = up.plan.run(); sync(); t2 = time.time()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        up.plan.run()
    sync(); t3 = time.time()
    print(f"mask {mask}: record {t1 - t0:.3f}s first-run {t2 - t1:.3f}s per-mask graph capture {t3 - t2:.3f}s ops {up.plan.num_ops()}")
pr = cProfile.Profile(); pr.enable()
up = _UNetPlan(model, B, 64, 64, (5, 6, 7, 8)); sync()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
guid = ClassifierGuidance(clf, 1.0)
active, per_step = resolve_candidate(CAND10, diffusion)
for i in range(3):
    t0 = time.time(); pl = SchedulePlan(model, active, per_step, B, cond_fn=guid, pack_uint8=True); sync()
    print(f"SchedulePlan build #{i}: {time.time() - t0:.3f}s launches {pl.launches}")
