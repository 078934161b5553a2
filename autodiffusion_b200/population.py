"""Population evaluation driver (BASELINE.json configs[2]): N candidates x num_samples images, sharded over the ranks
of one box, per-candidate FID from device-side feature moments.

The multi-GPU form of the reference's serial `is_legal -> get_cand_fid` loop
(GD/search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:340-367, 369-445, 447-470). Used by
`scripts/population_eval.py` and by `bench.py --gpus N` (key `population_eval`).
"""
from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np
import torch as th
import torch.distributed as dist

from .evaluator import CandidateEvaluator, FIDStatistics
from .search import draw_population


def projection_features(device, dim: int = 2048, image_size: int = 64, seed: int = 7) -> Callable[[th.Tensor], th.Tensor]:
    """Stand-in for Inception pool_3: a fixed random projection of the uint8 NHWC image to
    `dim` O(1) features. Exercises the statistic and the data path; the numbers are not Inception FIDs."""
    n_in = 3 * image_size * image_size
    proj = (th.randn(n_in, dim, generator=th.Generator().manual_seed(seed)) * (3.0 / n_in ** 0.5)).to(device)

    def feature_fn(u8: th.Tensor) -> th.Tensor:
        return (u8.reshape(u8.shape[0], -1).float() / 255.0 - 0.5) @ proj

    return feature_fn


def synthetic_reference_statistics(dim: int, seed: int = 11) -> FIDStatistics:
    rs = np.random.RandomState(seed)
    a = rs.randn(dim, dim) / dim ** 0.5
    return FIDStatistics(0.05 * rs.randn(dim), a @ a.T * 0.05 + 0.02 * np.eye(dim))


def run_population(model, diffusion, cond_fn, n_candidates: int, num_samples: int = 1000, batch_size: int = 256,
                   time_step: int = 10, max_prun: float = 0.1, seed: int = 0, feature_fn: Optional[Callable] = None,
                   feature_dim: int = 2048, fid_method: str = "eigh", warmup: bool = True, population=None) -> dict:
    """Evaluate a random population (`search.draw_population`: the reference's `sample_active_subnet` with the prune
    range fully open) with `CandidateEvaluator.evaluate_population`; returns the measurement as a dict (identical on
    every rank). Wall time is the max over ranks, bracketed by barriers + device synchronisation."""
    dev = model._device()
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if inited else 1
    if feature_fn is None:
        feature_fn = projection_features(dev, feature_dim, model.image_size)
    else:
        feature_dim = None
    ref_dim = feature_dim or 2048
    ev = CandidateEvaluator(model, diffusion, feature_fn, synthetic_reference_statistics(ref_dim), batch_size=batch_size,
                            num_samples=num_samples, image_size=model.image_size, seed=seed, cond_fn=cond_fn,
                            fid_method=fid_method, max_cached_plans=3)
    if population is None:
        population = draw_population(n_candidates, time_step, model.layer_num, max_prun, seed=seed)
    if warmup:  # one candidate outside the timed region: kernel attributes, NCCL communicator, allocator, pinned buffers
        ev.evaluate_population(draw_population(max(1, world), time_step, model.layer_num, max_prun, seed=seed + 991),
                               None)
        if world > 1:  # the moment all-reduce itself (NCCL connection set-up for a 33.6 MB fp64 buffer: ~0.35 s the first time)
            from .evaluator import MomentAccumulator

            MomentAccumulator(ref_dim, dev).all_reduce()

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    barrier()
    t0 = time.time()
    fids = ev.evaluate_population(population)
    barrier()
    wall = time.time() - t0
    info = dict(ev.last_population)
    if world > 1:
        t = th.tensor([wall, info["plan_build_overlapped_s"], info["plan_build_first_s"]], device=dev, dtype=th.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, info["plan_build_overlapped_s"], info["plan_build_first_s"] = (float(v) for v in t.tolist())
        f = th.tensor(fids, device=dev, dtype=th.float64)
        f0 = f.clone()
        dist.broadcast(f0, 0)
        assert th.equal(f, f0), "ranks disagree on the population's FIDs"
        ar = th.tensor([max(info["allreduce_ms"]) if info["allreduce_ms"] else 0.0], device=dev, dtype=th.float64)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        info["allreduce_ms"] = float(ar.item())  # the first tail all-reduce also absorbs the ranks' arrival skew
        # the collective alone: the same 33.6 MB fp64 buffer, every rank released by a barrier just before
        from .evaluator import MomentAccumulator

        probe = MomentAccumulator(ref_dim, dev)
        barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record()
        probe.all_reduce()
        e1.record()
        th.cuda.synchronize()
        pure = th.tensor([e0.elapsed_time(e1)], device=dev, dtype=th.float64)
        dist.all_reduce(pure, op=dist.ReduceOp.MAX)
        info["allreduce_alone_ms"] = float(pure.item())
    else:
        info["allreduce_ms"] = 0.0
        info["allreduce_alone_ms"] = 0.0
    n = len(population)
    return {
        "candidates": n, "num_samples": num_samples, "batch_size": batch_size, "n_gpus": world,
        "candidates_per_s": n / wall, "images_per_s": n * num_samples / wall, "wall_s": wall,
        "allreduce_ms": info["allreduce_ms"], "allreduce_alone_ms": round(info["allreduce_alone_ms"], 3), "plan_build_s": {"first": round(info["plan_build_first_s"], 3),
                                                               "overlapped_total_max_rank": round(info["plan_build_overlapped_s"], 3)},
        "schedule": {"whole_per_rank": info["whole_per_rank"], "batch_sharded_tail": info["shared"]},
        "fid_method": fid_method, "feature_dim": ref_dim, "guided": cond_fn is not None,
        "masks": f"sample_active_subnet, max_prun {max_prun}, {time_step} steps", "fids": fids,
    }
