#!/usr/bin/env python
"""Evolutionary search of (timesteps, block-skip lists) — the B200 twin of the reference's search main
GD/search_dynamic_unet_imagenet64_classifier_guidance_progressive.py (:720-830): same flags
(`--time_step`, `--max_epochs`, `--select_num`, `--population_num`, `--m_prob`, `--crossover_num`,
`--mutation_num`, `--max_prun`, `--min_prun`, `--use_ddim_init_x`, `--classifier_scale`, `--num_samples`,
`--batch_size`, model / classifier flags of script_util), same log lines.

    torchrun --nproc-per-node 8 scripts/search_candidates.py --attention_resolutions 32,16,8 --class_cond True \
        --image_size 64 --num_channels 192 --num_head_channels 64 --num_res_blocks 3 --resblock_updown True \
        --use_new_attention_order True --use_fp16 True --use_scale_shift_norm True --learn_sigma True \
        --noise_schedule cosine --use_dynamic_unet True --classifier_depth 4 --classifier_scale 1.0 \
        --model_path 64x64_diffusion.pt --classifier_path 64x64_classifier.pt --ref_path ref_stats.npz \
        --time_step 10 --max_prun 0.1 --num_samples 1000 --batch_size 250 --save_dir out/

`--mode timesteps` runs the timestep-only drivers instead (full architecture, individuals = timestep lists):
GD/search_imagenet64_classifier_guidance.py (classifier-guided; `--search_space "[926, 153, ...]"` builds the +-N/100
window around those steps as its `__main__` does, :645-668) or, with `--mode timesteps_uncond`,
GD/search_uncondition_model.py (no classifier, `--init_x "[644, 737, ...]"`), e.g. the LSUN-bedroom search
(GD/search_lsun_bedroom.sh).

Differences, all deliberate: one process per GPU (the reference forces world size 1, :757-760) with a
candidate's batches sharded over ranks; FID statistics from `--ref_path` as an .npz with `mu`, `sigma`
(the reference unpickles an object, :201-203); the Inception extractor is supplied by
`--feature_module pkg.mod:factory` (a callable uint8 NHWC -> [n, d] features on the device) because the
reference's TensorFlow graph cannot run here — without it a fixed random projection is used and the FID
values only rank candidates among themselves; `--state_path` makes the search resumable.
Without `--model_path` / `--classifier_path` the networks are random-init (functional / throughput runs).
"""
import argparse
import importlib
import os
import random
import sys
import time

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from autodiffusion_b200 import (add_dict_to_argparser, args_to_dict, classifier_defaults, create_classifier,  # noqa: E402
                                create_model_and_diffusion, model_and_diffusion_defaults)
from autodiffusion_b200.evaluator import FIDStatistics  # noqa: E402
from autodiffusion_b200.search import EvolutionSearcher  # noqa: E402


def create_argparser():
    defaults = dict(clip_denoised=True, num_samples=10000, batch_size=16, use_ddim=True, model_path="", save_dir="",
                    time_step=100, seed=0, max_epochs=20, select_num=10, population_num=50, m_prob=0.1, crossover_num=25,
                    mutation_num=35, classifier_path="", classifier_scale=1.0, max_fid=48.0, use_ddim_init_x=False,
                    index_step=None, max_prun=0.0, min_prun=0.0, ref_path="", feature_module="", feature_dim=2048,
                    state_path="", randomize_zero_init=False, mode="time_arch", search_space="", init_x="",
                    fid_method="sqrtm")
    defaults.update(model_and_diffusion_defaults())
    defaults.update(classifier_defaults())
    parser = argparse.ArgumentParser()
    add_dict_to_argparser(parser, defaults)
    return parser


def _randomize(module, seed):
    """Random-init networks output exactly zero (zero_module convs): re-draw all-zero tensors for dry runs."""
    g = th.Generator().manual_seed(seed)
    with th.no_grad():
        for p in module.parameters():
            if p.numel() > 0 and float(p.abs().max()) == 0.0:
                p.copy_(0.02 * th.randn(p.shape, generator=g))


def main():
    args = create_argparser().parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    th.manual_seed(args.seed)
    np.random.seed(args.seed)
    random.seed(args.seed)  # identical on every rank: all ranks draw the same individuals

    log_f = None
    if args.save_dir and rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)
        log_f = open(os.path.join(args.save_dir, "log.txt"), "a")

    def log(msg):
        if rank == 0:
            print(msg, flush=True)
            if log_f:
                log_f.write(str(msg) + "\n")
                log_f.flush()

    log("creating model and diffusion...")
    model, diffusion = create_model_and_diffusion(**args_to_dict(args, model_and_diffusion_defaults().keys()))
    if args.model_path:
        model.load_state_dict(th.load(args.model_path, map_location="cpu"))
    elif args.randomize_zero_init:
        _randomize(model, 1)
    model.to(dev).eval()
    if args.use_fp16:
        model.convert_to_fp16()
    classifier = None
    if args.mode != "timesteps_uncond":
        classifier = create_classifier(**args_to_dict(args, classifier_defaults().keys()))
        if args.classifier_path:
            classifier.load_state_dict(th.load(args.classifier_path, map_location="cpu"))
        elif args.randomize_zero_init:
            _randomize(classifier, 2)
        classifier.to(dev).eval()

    if args.feature_module:
        mod, fn = args.feature_module.split(":")
        feature_fn = getattr(importlib.import_module(mod), fn)(dev)
        d = args.feature_dim
    else:
        d = args.feature_dim
        hw = args.image_size
        proj = (th.randn(3 * hw * hw, d, generator=th.Generator().manual_seed(7)) * (3.0 / (3 * hw * hw) ** 0.5)).to(dev)
        feature_fn = lambda u8: (u8.reshape(u8.shape[0], -1).float() / 255.0 - 0.5) @ proj
    if args.ref_path:
        z = np.load(args.ref_path)
        ref_stats = FIDStatistics(z["mu"], z["sigma"])
    else:
        rs = np.random.RandomState(11)
        a = rs.randn(d, d) / d ** 0.5
        ref_stats = FIDStatistics(0.05 * rs.randn(d), a @ a.T * 0.05 + 0.02 * np.eye(d))

    t0 = time.time()
    from autodiffusion_b200.classifier import ClassifierGuidance
    from autodiffusion_b200.evaluator import CandidateEvaluator

    evaluator = CandidateEvaluator(
        model, diffusion, feature_fn, ref_stats, batch_size=args.batch_size, num_samples=args.num_samples,
        image_size=args.image_size, class_cond=args.class_cond, clip_denoised=args.clip_denoised, seed=args.seed,
        cond_fn=None if classifier is None else ClassifierGuidance(classifier, args.classifier_scale),
        fid_method=args.fid_method)  # "sqrtm" = the reference's arithmetic; "eigh" = the cheaper symmetric form
    if args.mode == "time_arch":
        searcher = EvolutionSearcher(args, model=model, base_diffusion=diffusion, time_step=args.time_step,
                                     classifier=classifier, index_step=args.index_step, evaluator=evaluator, log=log)
    else:
        from autodiffusion_b200.respace import space_timesteps
        from autodiffusion_b200.timestep_search import TimestepSearcher, build_search_space

        space = None
        if args.search_space:
            init = list(space_timesteps(diffusion.original_num_steps, ("ddim" if args.use_ddim else "") + str(args.time_step))) \
                if args.use_ddim_init_x else None
            space = build_search_space(eval(args.search_space), diffusion.original_num_steps, init)
            log("search space: " + str(space))
        searcher = TimestepSearcher(args, model=model, base_diffusion=diffusion, time_step=args.time_step, classifier=classifier,
                                    search_space=space, variant="uncondition" if args.mode == "timesteps_uncond" else "imagenet64",
                                    evaluator=evaluator, log=log)
    if args.state_path and os.path.exists(args.state_path):
        searcher.load_state(args.state_path)
        log("resumed from {} at epoch {} ({} individuals visited)".format(args.state_path, searcher.epoch, len(searcher.vis_dict)))
    searcher.search(state_path=args.state_path if (args.state_path and rank == 0) else None)
    log("total searching time = {:.2f} hours".format((time.time() - t0) / 3600))
    log("evaluated {} individuals, {:.3f} individuals/s".format(len(searcher.vis_dict), len(searcher.vis_dict) / (time.time() - t0)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
