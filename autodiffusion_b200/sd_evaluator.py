"""Candidate evaluation for the Stable-Diffusion search: the drop-in for `EvolutionSearcher.get_cand_fid` of
/root/reference/examples/"Stable Diffusion"/scripts/search_ea.py:504-566.

The reference loops over a validation loader of prompts, encodes them (`get_learned_conditioning`), samples latents with
the searched time steps through the chosen sampler, decodes them with the VAE, collects images on the host and computes
Inception activations and the Fréchet distance. Here the searched path - sampling - runs on the fused plan
(`sd_ddim.*Sampler`), and the surrounding pieces are callables supplied by the caller because they are outside the
searched path: `contexts(batch_index, n) -> (cond, uncond)` (text encoder), `decode(latents) -> anything the feature
extractor accepts` (VAE; identity if features are taken from latents), `feature_fn(decoded) -> fp32 [n, d]` on the device.
Features go straight into the fp64 moment kernel (`evaluator.MomentAccumulator`); nothing is gathered on the host.

`evaluate(cands)` scores a whole population: candidate i is handled entirely by rank i % world (one all-reduce of the
FID values at the end), which is the form `sd_search.EvolutionSearcher` calls once per generation.
"""
from __future__ import annotations

import zlib
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch as th
import torch.distributed as dist

from .evaluator import FIDStatistics, MomentAccumulator


class SDCandidateEvaluator:
    def __init__(self, sampler, contexts: Callable[[int, int], Tuple[th.Tensor, Optional[th.Tensor]]],
                 feature_fn: Callable[[th.Tensor], th.Tensor], ref_stats: FIDStatistics, batch_size: int, num_samples: int,
                 shape: Sequence[int] = (4, 64, 64), scale: float = 7.5, decode: Optional[Callable[[th.Tensor], th.Tensor]] = None,
                 seed: int = 0, dpm_solver: bool = False, rank: Optional[int] = None, world_size: Optional[int] = None, group=None):
        self.sampler, self.contexts, self.feature_fn, self.ref_stats = sampler, contexts, feature_fn, ref_stats
        self.decode = decode if decode is not None else (lambda z: z)
        self.batch_size, self.num_samples, self.shape, self.scale = batch_size, num_samples, tuple(shape), float(scale)
        self.seed, self.dpm_solver = seed, dpm_solver
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world_size = world_size if world_size is not None else (dist.get_world_size(group) if inited else 1)
        self.group = group
        self._acc: Optional[MomentAccumulator] = None
        self._ref_sqrt = None

    def _seed(self, cand_key: str, batch_index: int) -> int:
        return (self.seed * 0x9E3779B1 + zlib.crc32(cand_key.encode()) * 1000003 + batch_index * 7919 + 4242) % (2 ** 63 - 1)

    @th.no_grad()
    def sample_candidate(self, cand):
        """Yields (latents, decoded) per batch: search_ea.py:515-540 with `fixed_code`-style seeded start codes."""
        dev = self.sampler.model.device
        key = str(list(cand))
        nb = (self.num_samples + self.batch_size - 1) // self.batch_size
        S = len(cand) - 1 if self.dpm_solver else len(cand)
        for b in range(nb):
            g = th.Generator(device=dev)
            g.manual_seed(self._seed(key, b))
            x_T = th.randn((self.batch_size,) + self.shape, generator=g, device=dev)
            cond, uncond = self.contexts(b, self.batch_size)
            z, _ = self.sampler.sample(S=S, conditioning=cond, batch_size=self.batch_size, shape=list(self.shape), verbose=False,
                                       unconditional_guidance_scale=self.scale, unconditional_conditioning=uncond, eta=0.0,
                                       x_T=x_T, sampled_timestep=list(cand))
            keep = min(self.batch_size, self.num_samples - b * self.batch_size)
            yield z[:keep], self.decode(z[:keep])

    def get_cand_fid(self, cand) -> float:
        acc = None
        for _, img in self.sample_candidate(cand):
            feats = self.feature_fn(img).float()
            if acc is None:
                if self._acc is None or self._acc.dim != feats.shape[1]:
                    self._acc = MomentAccumulator(feats.shape[1], feats.device)
                acc = self._acc
                acc.reset()
            acc.add(feats)
        mu, sigma = acc.statistics()
        if self._ref_sqrt is None:
            import scipy.linalg as sl

            w, v = sl.eigh(np.atleast_2d(self.ref_stats.sigma))
            self._ref_sqrt = (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T
        return float(FIDStatistics(mu, sigma).frechet_distance_eigh(self.ref_stats, self._ref_sqrt))

    def evaluate(self, cands: List[list]) -> List[float]:
        """FIDs of a population, every rank returning all of them; candidate i is sampled and scored by rank i % world."""
        vals = [self.get_cand_fid(c) if i % self.world_size == self.rank else 0.0 for i, c in enumerate(cands)]
        if self.world_size > 1 and cands:
            nccl = dist.get_backend(self.group) == "nccl"
            t = th.tensor(vals, dtype=th.float64, device=self.sampler.model.device if nccl else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            vals = t.cpu().tolist()
        return vals
