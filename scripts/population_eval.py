#!/usr/bin/env python
"""Population evaluation (BASELINE.json configs[2]): N candidates x num_samples images of ADM-G 64x64,
batches sharded over the ranks of one box, per-candidate FID from device-side feature moments merged by
ONE NCCL all-reduce — the multi-GPU form of the reference's serial `is_legal` -> `get_cand_fid` loop
(GD/search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:340-367, 369-445, 447-470).

    python scripts/population_eval.py --candidates 50 --num_samples 1000 --batch_size 250
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/population_eval.py ...

Candidates are drawn the way the search draws its initial population (`sample_active_subnet` /
`get_random`, :284-338, 472-485): K distinct timesteps from [0, 1000) and, per step, each prunable
block id kept in the skip list with a probability chosen so that at most `--max_prun` of the K x L
(step, block) slots are skipped. Weights are random-init (no checkpoint offline) and the Inception
pool_3 extractor is replaced by a fixed random projection to `--feature_dim` features (SURVEY.md §8f
N2): the FID values only exercise the statistic; the throughput is what is measured.

Prints one JSON line: candidates/s, images/s, and the split between sampling (+ all-reduce) and the
host-side sqrtm.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from autodiffusion_b200 import (classifier_defaults, create_classifier, create_model_and_diffusion,  # noqa: E402
                                model_and_diffusion_defaults)
from autodiffusion_b200.classifier import ClassifierGuidance  # noqa: E402
from autodiffusion_b200.evaluator import CandidateEvaluator, FIDStatistics  # noqa: E402

ADM_FLAGS = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=192, num_head_channels=64, num_res_blocks=3,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=True, use_scale_shift_norm=True,
                 use_dynamic_unet=True)


def draw_candidate(rng: random.Random, time_step: int, layer_num: int, max_prun: float, mask_pool: int):
    """One random individual. `mask_pool` bounds the number of distinct non-empty skip sets a population
    uses, mirroring how crossover/mutation recombine a few masks (each distinct (batch, mask) pair costs
    one UNet plan recording, ~0.5 s)."""
    timesteps = rng.sample(range(1000), time_step)
    budget = int(max_prun * time_step * layer_num)
    skip_layers = [[] for _ in range(time_step)]
    pool_rng = random.Random(rng.randrange(mask_pool))  # a mask is a function of its pool index only
    n_masked_steps = min(time_step, max(1, budget // 9))
    for s in rng.sample(range(time_step), rng.randint(0, n_masked_steps)):
        skip_layers[s] = sorted(pool_rng.sample(range(layer_num), min(9, budget)))
        budget -= len(skip_layers[s])
        if budget <= 0:
            break
    return {"timesteps": timesteps, "skip_layers": skip_layers}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=8)
    ap.add_argument("--num_samples", type=int, default=1000)
    ap.add_argument("--batch_size", type=int, default=250)
    ap.add_argument("--time_step", type=int, default=10)
    ap.add_argument("--max_prun", type=float, default=0.1)
    ap.add_argument("--mask_pool", type=int, default=4)
    ap.add_argument("--feature_dim", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--small", action="store_true", help="64-channel 1-res-block UNet (functional check)")
    ap.add_argument("--features", default="projection", choices=["projection", "inception"],
                    help="projection: fixed random projection to --feature_dim; inception: Inception-V3 pool_3 on the device "
                         "(autodiffusion_b200.inception, random-init offline), 2048-d")
    ap.add_argument("--shard", default="candidates", choices=["candidates", "batches"],
                    help="candidates: each rank evaluates whole candidates (no per-candidate collective); batches: every "
                         "candidate's batches are split over ranks and its moments all-reduced")
    ap.add_argument("--fid_method", default="eigh", choices=["sqrtm", "eigh"])
    ap.add_argument("--guided", action="store_true", help="classifier guidance (native depth-4 noisy classifier, scale 1.0)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    flags = model_and_diffusion_defaults()
    flags.update(ADM_FLAGS)
    if args.small:
        flags.update(num_channels=64, num_res_blocks=1)
    model, diffusion = create_model_and_diffusion(**flags)
    g = th.Generator().manual_seed(args.seed)
    with th.no_grad():
        for name, p in model.named_parameters():  # same values on every rank
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_(th.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.1 * th.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * th.randn(p.shape, generator=g))
    model.to(dev).eval()
    model.convert_to_fp16()
    cond_fn = None
    if args.guided:
        cd = classifier_defaults()
        cd.update(classifier_depth=4 if not args.small else 1, classifier_width=128 if not args.small else 64)
        clf = create_classifier(**cd)
        with th.no_grad():
            for name, p in clf.named_parameters():
                if p.dim() > 1:
                    p.copy_(th.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
                elif name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * th.randn(p.shape, generator=g))
                else:
                    p.copy_(0.02 * th.randn(p.shape, generator=g))
        clf.to(dev).eval()
        cond_fn = ClassifierGuidance(clf, 1.0)

    d = args.feature_dim if args.features == "projection" else 2048
    proj = (th.randn(3 * 64 * 64, d, generator=th.Generator().manual_seed(7)) * (3.0 / (3 * 64 * 64) ** 0.5)).to(dev)

    def feature_fn(u8):  # stand-in for Inception pool_3: uint8 NHWC -> fp32 [n, d], O(1) entries
        return (u8.reshape(u8.shape[0], -1).float() / 255.0 - 0.5) @ proj

    if args.features == "inception":
        from autodiffusion_b200.inception import InceptionPool3

        inception = InceptionPool3().to(dev)
        feature_fn = inception  # noqa: F811  uint8 NHWC -> fp32 [n, 2048] on the device

    rs = np.random.RandomState(11)
    a = rs.randn(d, d) / d ** 0.5
    ref_stats = FIDStatistics(0.05 * rs.randn(d), a @ a.T * 0.05 + 0.02 * np.eye(d))

    ev = CandidateEvaluator(model, diffusion, feature_fn, ref_stats, batch_size=args.batch_size,
                            num_samples=args.num_samples, seed=args.seed, max_cached_plans=args.candidates + 1,
                            cond_fn=cond_fn, fid_method=args.fid_method)
    rng = random.Random(args.seed)
    population = [draw_candidate(rng, args.time_step, model.layer_num, args.max_prun, args.mask_pool)
                  for _ in range(args.candidates)]

    # warm-up: one candidate outside the timed region (kernel attributes, NCCL communicator, allocator)
    ev.get_cand_fid(population[0])
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    t0 = time.time()
    t_sample, t_fid, t_plan = 0.0, 0.0, 0.0
    pending, times = [], []
    for i, cand in enumerate(population):  # the host-side FID of candidate i overlaps the sampling of the next one
        pending.append(ev.submit_cand_fid(cand, _whole_on=(i % world) if args.shard == "candidates" else None))
        times.append(ev.last_times)
    fids = ev.resolve(pending)  # one all-reduce of the FID values
    for tm in times:
        t_plan += tm["reset_time"]
        t_sample += tm["sample_time"]
        t_fid += tm["fid_time"]
    if world > 1:
        dist.barrier()
    th.cuda.synchronize()
    wall = time.time() - t0
    if world > 1:
        t = th.tensor([wall], device=dev, dtype=th.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
        f = th.tensor(fids, device=dev, dtype=th.float64)
        f0 = f.clone()
        dist.broadcast(f0, 0)
        assert th.equal(f, f0), "ranks disagree on FID (the all-reduced moments must be identical)"
    if rank == 0:
        n = len(population)
        print(json.dumps({
            "metric": "population evaluation, candidates/s", "value": n / wall, "unit": "candidates/s",
            "images_per_s": n * args.num_samples / wall, "n_gpus": world, "candidates": n,
            "num_samples": args.num_samples, "batch_size": args.batch_size, "ddim_steps": args.time_step,
            "feature_dim": d, "wall_s": wall, "guided": bool(args.guided), "shard": args.shard, "fid_method": args.fid_method, "features": args.features,
            "split_s": {"plan_build": round(t_plan, 3), "sampling_plus_allreduce": round(t_sample, 3),
                        "host_sqrtm_fid_overlapped": round(t_fid, 3)},
            "fid_first3": [round(x, 4) for x in fids[:3]],
            "note": "random-init weights, random-projection features (no Inception graph offline): FID values "
                    "exercise the statistic only",
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
