"""Candidate evaluation: sample `num_samples` images for a candidate, reduce feature moments, FID.

Drop-in for `EvolutionSearcher.get_cand_fid` (…progressive.py:369-445; timestep-only twin
search_imagenet64_classifier_guidance.py:308-376):

    reference                                   here
    ---------                                   ----
    reset_diffusion(use_timesteps)              sampler.resolve_candidate (same tables, same map)
    while len(all_images)*B < num_samples:      static shard of batch indices over ranks
      classes = th.randint(...)                 per-(seed, candidate, batch) generator -> same images
      sample = ddim_sample_loop(model_fn, ...)  SchedulePlan.run (one CUDA graph per batch)
      uint8 NHWC pack                           last node of that graph
      dist.all_gather(images), (labels)         -- nothing is gathered --
    arr[:num_samples]                           rows beyond num_samples are dropped before reduction
    dist.barrier()                              subsumed by the all-reduce
    cal_fid(arr, ...) : Inception -> np.mean /  feature_fn (caller's extractor, on device) ->
      np.cov -> frechet_distance                adb_moments_accumulate (fp64 sum_x, sum_xx) ->
                                                ONE all-reduce of [n | sum_x | sum_xx] -> mu, sigma (fp64,
                                                N-1 denominator as np.cov) -> frechet_distance (scipy sqrtm)

The Inception-V3 pool_3 extractor itself is row N2 of SURVEY.md §8(f) (its TensorFlow graph file is
not available offline); `feature_fn` is any callable uint8 NHWC images -> fp32 [B, d] features.
"""
from __future__ import annotations

import time
import warnings
import zlib
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch as th
import torch.distributed as dist

from . import ops
from .sampler import SchedulePlan, resolve_candidate

NUM_CLASSES = 1000


class FIDStatistics:
    """evaluations/evaluator_v1.py:104-157 (duplicate: …progressive.py:104-153)."""

    def __init__(self, mu: np.ndarray, sigma: np.ndarray):
        self.mu = mu
        self.sigma = sigma

    def frechet_distance_eigh(self, other, other_sqrt=None):
        """Same statistic through a symmetric eigenproblem: tr sqrtm(S1 S2) = sum sqrt(eig(S2^1/2 S1 S2^1/2)).
        ~6x cheaper than the general sqrtm and always real; agrees with `frechet_distance` to rounding when both
        covariances are well conditioned. Not the default: the default follows the reference's arithmetic."""
        import scipy.linalg as sl

        mu1, sigma1 = np.atleast_1d(self.mu), np.atleast_2d(self.sigma)
        mu2, sigma2 = np.atleast_1d(other.mu), np.atleast_2d(other.sigma)
        if other_sqrt is None:
            w, v = sl.eigh(sigma2)
            other_sqrt = (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T
        ev = sl.eigvalsh(other_sqrt @ sigma1 @ other_sqrt)
        diff = mu1 - mu2
        return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.sqrt(np.clip(ev, 0.0, None)).sum()

    def frechet_distance(self, other, eps=1e-6):
        from scipy import linalg

        mu1, sigma1 = np.atleast_1d(self.mu), np.atleast_2d(self.sigma)
        mu2, sigma2 = np.atleast_1d(other.mu), np.atleast_2d(other.sigma)
        assert mu1.shape == mu2.shape, \
            f"Training and test mean vectors have different lengths: {mu1.shape}, {mu2.shape}"
        assert sigma1.shape == sigma2.shape, \
            f"Training and test covariances have different dimensions: {sigma1.shape}, {sigma2.shape}"
        diff = mu1 - mu2
        # scipy >= 1.16 removed sqrtm's `disp` argument; the reference discarded its second result
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
        if not np.isfinite(covmean).all():
            warnings.warn("fid calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
            offset = np.eye(sigma1.shape[0]) * eps
            covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
        if np.iscomplexobj(covmean):
            if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
                raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
            covmean = covmean.real
        return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


class MomentAccumulator:
    """n, sum_x[d], sum_xx[d,d] in fp64, packed in one buffer so a single all-reduce merges ranks."""

    def __init__(self, dim: int, device):
        self.dim = dim
        self.buf = th.zeros(1 + dim + dim * dim, dtype=th.float64, device=device)

    @property
    def n(self) -> th.Tensor:
        return self.buf[0:1]

    @property
    def sum_x(self) -> th.Tensor:
        return self.buf[1:1 + self.dim]

    @property
    def sum_xx(self) -> th.Tensor:
        return self.buf[1 + self.dim:].view(self.dim, self.dim)

    def reset(self):
        self.buf.zero_()

    def add(self, feats: th.Tensor):
        """feats fp32 [n, d] on the device (CUDA kernel; no CPU path)."""
        feats = feats.float().contiguous()
        assert feats.dim() == 2 and feats.shape[1] == self.dim
        if feats.shape[0] == 0:
            return
        ops.moments_accumulate(feats, self.sum_x, self.sum_xx)
        self.n.add_(float(feats.shape[0]))

    def load_partial(self, n: int, sum_x, sum_xx):
        """Install externally computed partial sums (used by the CPU/gloo reduction tests)."""
        self.buf[0] = float(n)
        self.sum_x.copy_(th.as_tensor(sum_x, dtype=th.float64))
        self.sum_xx.copy_(th.as_tensor(sum_xx, dtype=th.float64))

    def all_reduce(self, group=None):
        """The path's one collective: sum over ranks of [n | sum_x | sum_xx] (33.6 MB at d=2048)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)

    def statistics(self):
        """mu = mean, sigma = unbiased covariance (np.cov's N-1, evaluator_v1.py:219-220), fp64 on the host."""
        return self.statistics_from(self.buf.detach().cpu().numpy(), self.dim)

    @staticmethod
    def statistics_from(buf: np.ndarray, d: int):
        n = float(buf[0])
        if n < 2:
            raise ValueError(f"need at least 2 samples for a covariance, have {n}")
        sx = buf[1:1 + d]
        sxx = buf[1 + d:].reshape(d, d)
        mu = sx / n
        sigma = (sxx - n * np.outer(mu, mu)) / (n - 1.0)
        return mu, sigma


class _RemoteFid:
    """Placeholder for a candidate whose host-side FID another rank computes (`CandidateEvaluator.resolve`)."""

    def __init__(self, owner: int):
        self.owner = owner

    def result(self):
        raise RuntimeError(f"this candidate's FID is computed on rank {self.owner}: use CandidateEvaluator.resolve(futures)")


def batch_seed(seed: int, cand_key: str, batch_index: int) -> int:
    """Noise/label seed of one batch: a function of (global seed, candidate, batch index) only, so a
    candidate's images do not depend on how many ranks share the work (the reference seeds every
    rank identically, …progressive.py:762-765, which would duplicate images across ranks)."""
    return (seed * 0x9E3779B1 + zlib.crc32(cand_key.encode()) * 1000003 + batch_index * 7919 + 12345) % (2 ** 63 - 1)


def shard_batches(num_batches: int, rank: int, world_size: int) -> Sequence[int]:
    """Static round-robin of a candidate's batches over ranks; no data-path communication."""
    return list(range(rank, num_batches, world_size))


class CandidateEvaluator:
    """Owns the model, the base diffusion, the feature extractor and the reference statistics.

    `get_cand_fid(cand, args)` keeps the reference's signature; `args` may carry `batch_size`,
    `num_samples`, `image_size`, `class_cond`, `clip_denoised` (…progressive.py:402-420) and
    overrides the constructor's values when given.
    """

    def __init__(self, model, base_diffusion, feature_fn: Callable[[th.Tensor], th.Tensor], ref_stats: FIDStatistics,
                 batch_size: int = 100, num_samples: int = 1000, image_size: int = 64, class_cond: bool = True,
                 clip_denoised: bool = True, cond_fn: Optional[Callable] = None, seed: int = 0,
                 rank: Optional[int] = None, world_size: Optional[int] = None, group=None, max_cached_plans: int = 8,
                 fid_method: str = "eigh", fid_threads: Optional[int] = None, shard_fid: bool = True):
        self.model = model
        self.base_diffusion = base_diffusion
        self.feature_fn = feature_fn
        self.ref_stats = ref_stats
        self.batch_size, self.num_samples, self.image_size = batch_size, num_samples, image_size
        self.class_cond, self.clip_denoised, self.cond_fn = class_cond, clip_denoised, cond_fn
        self.seed = seed
        inited = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if inited else 0)
        self.world_size = world_size if world_size is not None else (dist.get_world_size(group) if inited else 1)
        self.group = group
        self._plans: Dict[tuple, SchedulePlan] = {}
        self._max_cached = max_cached_plans
        self._acc: Optional[MomentAccumulator] = None
        self.last_times: Dict[str, float] = {}
        self.vis_dict: Dict[str, dict] = {}
        # deferred FID: the host-side sqrtm of candidate i runs on this worker while candidate i+1 samples
        self._fid_pool: Optional[ThreadPoolExecutor] = None
        self._host_bufs: list = []
        assert fid_method in ("sqrtm", "eigh")
        # "eigh" (default): frechet_distance_eigh - the same statistic through a symmetric eigenproblem, equal to the
        # reference's sqrtm arithmetic to rounding (tests/test_evaluator_cpu.py; 4+ decimals on the d = 2048 population
        # runs under profiles/) and ~30x cheaper on the host. "sqrtm": scipy.linalg.sqrtm exactly as the reference
        # (evaluator_v1.py:114-157); its Python-level Schur loops hold the GIL and were measured to stretch the
        # sampling thread's plan building 8x when run beside it.
        self.fid_method = fid_method
        self._ref_sqrt = None
        # BLAS threads for the host-side FID. torchrun exports OMP_NUM_THREADS=1, which would make the 2048x2048
        # sqrtm take ~10 s; the worker lifts the limit for its own calls (threadpoolctl) to this many threads.
        import os

        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        self.fid_threads = fid_threads if fid_threads is not None else max(1, (os.cpu_count() or 1) // max(1, local_world))
        # The host-side sqrtm (several seconds at d=2048) is the serial term once sampling is sharded: every rank
        # holds the all-reduced moments, so candidate number i is finished by rank i % world only and the values
        # are exchanged in `resolve` (one tiny all-reduce per batch of candidates).
        self.shard_fid = shard_fid
        self._seq = 0

    # ---- plan cache keyed by what the launch schedule depends on ----
    def _plan_for(self, cand, batch: int) -> SchedulePlan:
        active, per_step = resolve_candidate(cand, self.base_diffusion)
        key = (tuple(active.timestep_map), tuple(tuple(s) for s in per_step), batch, self.image_size, self.clip_denoised)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= self._max_cached:
                self._plans.pop(next(iter(self._plans)))
            plan = SchedulePlan(self.model, active, per_step, batch, image_size=self.image_size,
                                clip_denoised=self.clip_denoised, cond_fn=self.cond_fn, pack_uint8=True)
            self._plans[key] = plan
        return plan

    def sample_batch(self, plan: SchedulePlan, cand_key: str, batch_index: int):
        """-> (uint8 NHWC images on the device, labels)."""
        dev = self.model._device()
        g = th.Generator(device=dev)
        g.manual_seed(batch_seed(self.seed, cand_key, batch_index))
        y = th.randint(0, NUM_CLASSES, (plan.B,), generator=g, device=dev) if self.class_cond else None
        noise = th.randn(plan.shape, generator=g, device=dev)
        plan.run(noise, y)
        return plan.u8, y

    def get_cand_fid(self, cand=None, args=None) -> float:
        """…progressive.py:369-445: sample, reduce, FID - blocking, like the reference call."""
        return self.resolve([self.submit_cand_fid(cand, args)])[0]

    def resolve(self, futures) -> list:
        """FID values of `submit_cand_fid` futures, in order, on every rank. Each rank waits for the candidates it
        owns; with more than one rank the values are exchanged by one all-reduce of len(futures) doubles (every
        rank must call this with the futures of the same candidates - they do: all ranks walk the same population)."""
        vals = [0.0 if isinstance(f, _RemoteFid) else float(f.result()) for f in futures]
        if any(isinstance(f, _RemoteFid) for f in futures) or (self.shard_fid and self.world_size > 1 and futures):
            nccl = dist.get_backend(self.group) == "nccl"
            t = th.tensor(vals, dtype=th.float64, device=self.model._device() if nccl else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            vals = t.cpu().tolist()
        return vals

    def evaluate_population(self, cands, args=None) -> list:
        """FID of every candidate in `cands`, on every rank, with the POPULATION sharded over ranks: candidate i is
        sampled (all of its batches), reduced and scored by rank i % world alone - no per-candidate collective, plan
        building and the host-side sqrtm parallelise with the sampling - and the values are exchanged by one
        all-reduce of len(cands) doubles at the end. Images are those of the batch-sharded path (seeds depend on
        (seed, candidate, batch index) only). Every rank must pass the same list."""
        futures = [self.submit_cand_fid(c, args, _whole_on=i % self.world_size) for i, c in enumerate(cands)]
        return self.resolve(futures)

    def submit_cand_fid(self, cand=None, args=None, _whole_on: Optional[int] = None) -> Future:
        """Same work as `get_cand_fid`, but only the device part (sampling, moments, all-reduce, D2H of the
        33.6 MB moment buffer) happens before this returns; mu / sigma / sqrtm run on a host worker thread.
        The search driver keeps sampling the next candidate meanwhile (`fid_time` was serial in the
        reference, :437-443)."""
        if args is not None:
            for k in ("batch_size", "num_samples", "image_size", "class_cond", "clip_denoised"):
                if hasattr(args, k):
                    setattr(self, k, getattr(args, k))
        if _whole_on is not None and _whole_on != self.rank:  # another rank evaluates this candidate entirely
            self.last_times = dict(reset_time=0.0, sample_time=0.0, fid_time=0.0)
            return _RemoteFid(_whole_on)
        t0 = time.time()
        plan = self._plan_for(cand, self.batch_size)
        reset_time = time.time() - t0
        t0 = time.time()
        cand_key = str(cand)
        num_batches = (self.num_samples + self.batch_size - 1) // self.batch_size
        acc = None
        my_batches = range(num_batches) if _whole_on is not None else shard_batches(num_batches, self.rank, self.world_size)
        for b in my_batches:
            images, _ = self.sample_batch(plan, cand_key, b)
            keep = min(self.batch_size, self.num_samples - b * self.batch_size)  # arr[:num_samples], :432-433
            feats = self.feature_fn(images[:keep])
            if acc is None:
                if self._acc is None or self._acc.dim != feats.shape[1]:
                    self._acc = MomentAccumulator(feats.shape[1], feats.device)
                acc = self._acc
                acc.reset()
            acc.add(feats)
        if acc is None:  # this rank had no batch of this candidate
            dim = self.ref_stats.mu.shape[0]
            if self._acc is None or self._acc.dim != dim:
                self._acc = MomentAccumulator(dim, self.model._device())
            acc = self._acc
            acc.reset()
        if _whole_on is None:
            acc.all_reduce(self.group)
        seq = self._seq
        self._seq += 1
        if _whole_on is None and self.shard_fid and self.world_size > 1 and seq % self.world_size != self.rank:
            th.cuda.current_stream().synchronize() if acc.buf.is_cuda else None
            self.last_times = dict(reset_time=reset_time, sample_time=time.time() - t0, fid_time=0.0)
            return _RemoteFid(seq % self.world_size)
        host = self._host_bufs.pop() if self._host_bufs and self._host_bufs[-1].numel() == acc.buf.numel() \
            else th.empty(acc.buf.numel(), dtype=th.float64).pin_memory()
        host.copy_(acc.buf, non_blocking=True)
        done = th.cuda.Event()
        done.record()
        th.cuda.current_stream().synchronize()  # sample_time as the reference logs it (:435)
        sample_time = time.time() - t0
        dim = acc.dim
        times = dict(reset_time=reset_time, sample_time=sample_time, fid_time=0.0)
        self.last_times = times

        def finish() -> float:
            t1 = time.time()
            done.synchronize()
            mu, sigma = MomentAccumulator.statistics_from(host.numpy(), dim)
            self._host_bufs.append(host)
            try:
                from threadpoolctl import threadpool_limits
                ctx = threadpool_limits(limits=self.fid_threads, user_api="blas")
            except Exception:  # threadpoolctl missing: run with whatever the BLAS was started with
                import contextlib
                ctx = contextlib.nullcontext()
            with ctx:
                if self.fid_method == "eigh":
                    if self._ref_sqrt is None:
                        import scipy.linalg as sl
                        w, v = sl.eigh(np.atleast_2d(self.ref_stats.sigma))
                        self._ref_sqrt = (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T
                    fid = float(FIDStatistics(mu, sigma).frechet_distance_eigh(self.ref_stats, self._ref_sqrt))
                else:
                    fid = float(FIDStatistics(mu, sigma).frechet_distance(self.ref_stats))
            times["fid_time"] = time.time() - t1
            return fid

        if self._fid_pool is None:
            self._fid_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="adb-fid")
        return self._fid_pool.submit(finish)

    def is_legal(self, cand: str, log: Callable[[str], None] = print) -> bool:
        """…progressive.py:355-367: candidates are `str(dict)` keys; same log line format."""
        import ast

        info = self.vis_dict.setdefault(cand, {})
        if "visited" in info:
            log("cand: {} has visited!".format(cand))
            return False
        info["fid"] = self.get_cand_fid(cand=ast.literal_eval(cand))
        log("cand: {}, fid: {}".format(cand, info["fid"]))
        info["visited"] = True
        return True
