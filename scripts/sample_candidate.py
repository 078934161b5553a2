#!/usr/bin/env python
"""Replay a searched candidate and write samples_{N}x{H}x{W}x3.npz — the B200 twin of the reference's
stand-alone sampler GD/scripts/classifier_sample_prunedUNET.py (:85-219): same flags
(`--use_timestep`, `--skip_layers`, model flags of script_util.model_and_diffusion_defaults,
`--batch_size`, `--num_samples`, `--model_path`, `--save_dir`), same output file layout
(arr_0 = uint8 NHWC images, arr_1 = int64 labels).

Differences, all deliberate: candidates are parsed with ast.literal_eval (the reference eval()s them,
:158-165); one process per GPU via torchrun shards the batches (rank r takes batches r, r+world, ...)
with per-batch seeds, so the images do not depend on the number of GPUs; no classifier is loaded here
(classifier guidance enters through `--classifier_module pkg.mod:factory`, a caller-supplied cond_fn).

    torchrun --nproc-per-node 8 scripts/sample_candidate.py --class_cond True --image_size 64 ... \
        --use_timestep '[744,137,...]' --skip_layers '[[],[],...]' --num_samples 50000 --batch_size 256
"""
import argparse
import ast
import importlib
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from autodiffusion_b200 import (add_dict_to_argparser, args_to_dict, create_model_and_diffusion,  # noqa: E402
                                model_and_diffusion_defaults)
from autodiffusion_b200.evaluator import batch_seed, shard_batches  # noqa: E402
from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate  # noqa: E402

NUM_CLASSES = 1000


def main():
    defaults = dict(clip_denoised=True, num_samples=10000, batch_size=16, use_ddim=True, model_path="",
                    save_dir="./samples", use_timestep=None, skip_layers=None, seed=0, classifier_module="",
                    classifier_scale=1.0)
    defaults.update(model_and_diffusion_defaults())
    parser = argparse.ArgumentParser()
    add_dict_to_argparser(parser, defaults)
    args = parser.parse_args()
    if not args.use_ddim:
        raise SystemExit("only DDIM sampling is on the evaluator path (use_ddim=True)")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model, diffusion = create_model_and_diffusion(**args_to_dict(args, model_and_diffusion_defaults().keys()))
    if args.model_path:
        model.load_state_dict(th.load(args.model_path, map_location="cpu"))
    model.to(dev)
    if args.use_fp16:
        model.convert_to_fp16()
    model.eval()

    timesteps = ast.literal_eval(args.use_timestep) if args.use_timestep else sorted(diffusion.use_timesteps)
    cand = {"timesteps": timesteps}
    if args.skip_layers:
        cand["skip_layers"] = ast.literal_eval(args.skip_layers)
    cond_fn = None
    if args.classifier_module:
        mod, fn = args.classifier_module.split(":")
        cond_fn = getattr(importlib.import_module(mod), fn)(args, dev)  # -> callable(x, t, y=...) -> grad * scale

    active, per_step = resolve_candidate(cand, diffusion)
    plan = SchedulePlan(model, active, per_step, args.batch_size, image_size=args.image_size,
                        clip_denoised=args.clip_denoised, cond_fn=cond_fn, pack_uint8=True)
    nb = (args.num_samples + args.batch_size - 1) // args.batch_size
    images, labels = [], []
    for b in shard_batches(nb, rank, world):
        g = th.Generator(device=dev)
        g.manual_seed(batch_seed(args.seed, str(cand), b))
        y = th.randint(0, NUM_CLASSES, (args.batch_size,), generator=g, device=dev) if args.class_cond else None
        noise = th.randn(plan.shape, generator=g, device=dev)
        plan.run(noise, y)
        keep = min(args.batch_size, args.num_samples - b * args.batch_size)
        images.append((b, plan.u8[:keep].cpu().numpy()))
        labels.append((b, (y[:keep].cpu().numpy() if y is not None else np.zeros(keep, dtype=np.int64))))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (images, labels))
        images = [x for part in gathered for x in part[0]]
        labels = [x for part in gathered for x in part[1]]
    if rank == 0:
        arr = np.concatenate([a for _, a in sorted(images, key=lambda t: t[0])], axis=0)[: args.num_samples]
        lab = np.concatenate([a for _, a in sorted(labels, key=lambda t: t[0])], axis=0)[: args.num_samples]
        os.makedirs(args.save_dir, exist_ok=True)
        out_path = os.path.join(args.save_dir, "samples_" + "x".join(str(v) for v in arr.shape) + ".npz")
        print("saving to " + out_path)
        np.savez(out_path, arr, lab)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
