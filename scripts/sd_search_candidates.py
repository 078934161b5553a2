#!/usr/bin/env python
"""Twin of /root/reference/examples/"Stable Diffusion"/scripts/search_ea.py's `__main__` on the B200 path: evolutionary
search over sampling time steps of the SD-v1 UNet with DDIM / PLMS / DPM-Solver++(2M) and classifier-free guidance.

Offline there is no checkpoint, text encoder, VAE or Inception: the UNet is random-init, prompt encodings are synthetic
77x768 contexts, features are a fixed random projection of the latents - the search mechanics, the samplers and the
multi-GPU population sharding are what this exercises. One process per GPU:
    torchrun --nproc-per-node N scripts/sd_search_candidates.py --time_step 10 --population_num 8 ...
"""
import argparse
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch as th
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from autodiffusion_b200.evaluator import FIDStatistics  # noqa: E402
from autodiffusion_b200.sd_ddim import DDIMSampler, DPMSolverSampler, LatentDiffusionUNet, PLMSSampler  # noqa: E402
from autodiffusion_b200.sd_evaluator import SDCandidateEvaluator  # noqa: E402
from autodiffusion_b200.sd_search import EvolutionSearcher  # noqa: E402
from autodiffusion_b200.sd_unet import UNetModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sampler", default="ddim", choices=["ddim", "plms", "dpm"])
    ap.add_argument("--time_step", type=int, default=10)
    ap.add_argument("--max_epochs", type=int, default=2)
    ap.add_argument("--population_num", type=int, default=8)
    ap.add_argument("--select_num", type=int, default=4)
    ap.add_argument("--mutation_num", type=int, default=3)
    ap.add_argument("--crossover_num", type=int, default=2)
    ap.add_argument("--m_prob", type=float, default=0.25)
    ap.add_argument("--num_sample", type=int, default=64)
    ap.add_argument("--n_samples", type=int, default=32, help="batch size (search_ea.py: --n_samples)")
    ap.add_argument("--scale", type=float, default=7.5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--use_ddim_init_x", action="store_true")
    ap.add_argument("--small", action="store_true", help="64-channel UNet (functional check)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mc, cd = (64, 128) if args.small else (320, 768)
    unet = UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=mc, attention_resolutions=[4, 2, 1],
                     num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                     context_dim=cd, use_checkpoint=True, legacy=False)
    g = th.Generator().manual_seed(1)
    with th.no_grad():
        for name, p in unet.named_parameters():  # random-init incl. the zero modules (otherwise the net outputs 0)
            if p.dim() > 1:
                p.copy_(th.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
    unet.to(dev).eval()
    ld = LatentDiffusionUNet(unet)
    sampler = {"ddim": DDIMSampler, "plms": PLMSSampler, "dpm": DPMSolverSampler}[args.sampler](ld)
    d = 256
    proj = (th.randn(4 * 64 * 64, d, generator=th.Generator().manual_seed(7)) / (4 * 64 * 64) ** 0.5).to(dev)
    rs = np.random.RandomState(11)
    a = rs.randn(d, d) / d ** 0.5
    ref_stats = FIDStatistics(0.05 * rs.randn(d), a @ a.T * 0.05 + 0.02 * np.eye(d))

    def contexts(b, n):  # stand-in for get_learned_conditioning(prompts) / (n * [""]) (search_ea.py:519-525)
        gg = th.Generator(device=dev)
        gg.manual_seed(1000 + b)
        cond = th.randn((n, 77, cd), generator=gg, device=dev)
        uncond = th.zeros((n, 77, cd), device=dev) if args.scale != 1.0 else None
        return cond, uncond

    ev = SDCandidateEvaluator(sampler, contexts, lambda z: z.reshape(z.shape[0], -1) @ proj, ref_stats, batch_size=args.n_samples,
                              num_samples=args.num_sample, scale=args.scale, seed=args.seed, dpm_solver=args.sampler == "dpm")
    dpm_params = None
    if args.sampler == "dpm":  # search_ea.py:889-902
        dpm_params = {"full_timesteps": [v.item() for v in list(th.linspace(1.0, 0.001, 1001))],
                      "init_timesteps": [v.item() for v in list(th.linspace(1.0, 0.001, args.time_step + 1))]}
    opt = types.SimpleNamespace(max_epochs=args.max_epochs, select_num=args.select_num, population_num=args.population_num,
                                m_prob=args.m_prob, crossover_num=args.crossover_num, mutation_num=args.mutation_num,
                                use_ddim_init_x=args.use_ddim_init_x, dpm_solver=args.sampler == "dpm")
    lines = []
    s = EvolutionSearcher(opt, args.time_step, ev.evaluate, ddpm_num_timesteps=1000, dpm_params=dpm_params,
                          log=(lambda m: (lines.append(m), print(m, file=sys.stderr))) if rank == 0 else lines.append)
    random.seed(args.seed)  # every rank walks the same population (search_ea.py seeds all ranks alike)
    np.random.seed(args.seed)
    th.cuda.synchronize()
    t0 = time.time()
    top = s.search()
    th.cuda.synchronize()
    wall = time.time() - t0
    if rank == 0:
        n = len(s.vis_dict)
        print(json.dumps({"metric": "SD time-step search, candidates/s", "value": n / wall, "unit": "candidates/s", "n_gpus": world,
                          "sampler": args.sampler, "candidates": n, "num_sample": args.num_sample, "batch": args.n_samples,
                          "time_step": args.time_step, "epochs": args.max_epochs, "wall_s": wall,
                          "latent_images_per_s": n * args.num_sample / wall, "best": top[0],
                          "best_fid": s.vis_dict[top[0]]["fid"], "gpu_launches": unet.gpu_launches}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
