#!/usr/bin/env python
"""Population evaluation (BASELINE.json configs[2]): N candidates x num_samples images of ADM-G 64x64,
batches sharded over the ranks of one box, per-candidate FID from device-side feature moments merged by
ONE NCCL all-reduce — the multi-GPU form of the reference's serial `is_legal` -> `get_cand_fid` loop
(GD/search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:340-367, 369-445, 447-470).

    python scripts/population_eval.py --candidates 50 --num_samples 1000 --batch_size 250
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/population_eval.py ...

Candidates are drawn by the search's own `sample_active_subnet` (:284-338, restated in
autodiffusion_b200/search.py) with the prune range fully open (`[0, --max_prun]`): K timesteps and, per
step, a random skip list of up to max_prun x 58 block ids - nearly every step of every candidate is a new
launch plan. Whole candidates are placed on ranks longest-first, the `n mod world` tail is split by batches
and merged by one NCCL all-reduce of the moment buffer each (`evaluator.schedule_population`); the launch plan
of the next candidate is recorded while the current one samples. Weights are random-init (no checkpoint offline) and the Inception
pool_3 extractor is replaced by a fixed random projection to `--feature_dim` features (SURVEY.md §8f
N2): the FID values only exercise the statistic; the throughput is what is measured.

Prints one JSON line: candidates/s, images/s, and the split between sampling (+ all-reduce) and the
host-side sqrtm.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from autodiffusion_b200 import (classifier_defaults, create_classifier, create_model_and_diffusion,  # noqa: E402
                                model_and_diffusion_defaults)
from autodiffusion_b200.classifier import ClassifierGuidance  # noqa: E402
from autodiffusion_b200.evaluator import CandidateEvaluator  # noqa: E402

ADM_FLAGS = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=192, num_head_channels=64, num_res_blocks=3,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=True, use_scale_shift_norm=True,
                 use_dynamic_unet=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=8)
    ap.add_argument("--num_samples", type=int, default=1000)
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--time_step", type=int, default=10)
    ap.add_argument("--max_prun", type=float, default=0.1)
    ap.add_argument("--feature_dim", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--small", action="store_true", help="64-channel 1-res-block UNet (functional check)")
    ap.add_argument("--features", default="projection", choices=["projection", "inception"],
                    help="projection: fixed random projection to --feature_dim; inception: Inception-V3 pool_3 on the device "
                         "(autodiffusion_b200.inception, random-init offline), 2048-d")
    ap.add_argument("--shard", default="candidates", choices=["candidates", "batches"],
                    help="candidates: whole candidates longest-first + batch-sharded tail (evaluate_population); batches: "
                         "every candidate's batches are split over ranks and its moments all-reduced")
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

    feature_fn = None
    if args.features == "inception":
        from autodiffusion_b200.inception import InceptionPool3

        feature_fn = InceptionPool3().to(dev)  # uint8 NHWC -> fp32 [n, 2048] on the device

    if args.shard == "batches":  # every candidate batch-sharded + one moment all-reduce each (the single-candidate path)
        from autodiffusion_b200.population import projection_features, synthetic_reference_statistics
        from autodiffusion_b200.search import draw_population

        ff = feature_fn or projection_features(dev, args.feature_dim)
        ev = CandidateEvaluator(model, diffusion, ff, synthetic_reference_statistics(2048 if feature_fn else args.feature_dim),
                                batch_size=args.batch_size, num_samples=args.num_samples, seed=args.seed, cond_fn=cond_fn,
                                fid_method=args.fid_method)
        population = draw_population(args.candidates, args.time_step, model.layer_num, args.max_prun, seed=args.seed)
        ev.get_cand_fid(population[0])
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()
        t0 = time.time()
        fids = ev.resolve([ev.submit_cand_fid(c) for c in population])
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()
        wall = time.time() - t0
        res = {"candidates": len(population), "candidates_per_s": len(population) / wall,
               "images_per_s": len(population) * args.num_samples / wall, "wall_s": wall, "n_gpus": world, "fids": fids,
               "shard": "batches", "fid_method": args.fid_method}
    else:
        from autodiffusion_b200.population import run_population

        res = run_population(model, diffusion, cond_fn, args.candidates, num_samples=args.num_samples,
                             batch_size=args.batch_size, time_step=args.time_step, max_prun=args.max_prun, seed=args.seed,
                             feature_fn=feature_fn, feature_dim=args.feature_dim, fid_method=args.fid_method)
        res["shard"] = "candidates longest-first + batch-sharded tail"
    if rank == 0:
        fids = res.pop("fids")
        res.update({"metric": "population evaluation, candidates/s", "value": res["candidates_per_s"], "unit": "candidates/s",
                    "features": args.features, "fid_first3": [round(x, 4) for x in fids[:3]], "fid_checksum": float(np.sum(fids)),
                    "note": "random-init weights" + ("" if args.features == "inception" else
                                                     ", random-projection features (stand-in for Inception pool_3)") +
                            ": FID values exercise the statistic only"})
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
