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


def fid_owner(cand_key: str, world_size: int) -> int:
    """The rank that finishes the host-side FID of a batch-sharded candidate: a function of the candidate alone, so
    every rank names the same owner whatever it evaluated before."""
    return zlib.crc32(cand_key.encode()) % max(1, world_size)


def schedule_population(costs: Sequence[float], num_batches: int, world_size: int):
    """Static schedule of a population over the ranks of one box (longest first).

    `costs[i]` = estimated work of candidate i (any unit; equal for all of its `num_batches` batches).
    Returns (whole, shared):
      whole  = {rank: [candidate indices, in the order that rank runs them]} - these candidates are sampled, reduced
               and scored entirely by one rank: no collective, plan building and the host FID stay rank-local;
      shared = [(candidate index, {rank: [batch indices]})] - the tail that does not divide by the world size is
               split at batch granularity; each of these candidates costs one all-reduce of its moment buffer.
    Whole candidates are placed longest-processing-time-first on the least loaded rank; the `n mod world` cheapest
    ones form the tail and their batches fill the ranks up in the same greedy way, so 50 candidates on 8 ranks end at
    25 batches per rank instead of 7-vs-6 whole candidates (89 %). Every rank computes the same schedule."""
    n = len(costs)
    world = max(1, int(world_size))
    order = sorted(range(n), key=lambda i: (-float(costs[i]), i))
    n_whole = n if world == 1 else (n // world) * world
    load = [0.0] * world
    whole = {r: [] for r in range(world)}
    for i in order[:n_whole]:
        r = min(range(world), key=lambda k: (load[k], k))
        whole[r].append(i)
        load[r] += float(costs[i])
    shared = []
    for i in order[n_whole:]:
        per = {}
        for b in range(num_batches):
            r = min(range(world), key=lambda k: (load[k], k))
            per.setdefault(r, []).append(b)
            load[r] += float(costs[i]) / max(1, num_batches)
        shared.append((i, per))
    return whole, shared


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
                 fid_method: str = "sqrtm", fid_threads: Optional[int] = None, shard_fid: bool = True):
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
        self.last_population: Dict[str, object] = {}
        self.vis_dict: Dict[str, dict] = {}
        # deferred FID: the host-side sqrtm of candidate i runs on this worker while candidate i+1 samples
        self._fid_pool: Optional[ThreadPoolExecutor] = None
        self._host_bufs: list = []
        assert fid_method in ("sqrtm", "eigh")
        # "sqrtm" (default): scipy.linalg.sqrtm exactly as the reference (evaluator_v1.py:114-157), including the
        # eps-regularised retry and the complex-result handling. "eigh": frechet_distance_eigh - the same statistic
        # through a symmetric eigenproblem (tests/test_evaluator_cpu.py bounds the difference, also for n < d), ~30x
        # cheaper on the host and free of sqrtm's GIL-holding Python-level Schur loops; an explicit opt-in for
        # throughput runs (scripts/population_eval.py, scripts/search_candidates.py --fid_method eigh).
        self.fid_method = fid_method
        self._ref_sqrt = None
        # BLAS threads for the host-side FID. torchrun exports OMP_NUM_THREADS=1, which would make the 2048x2048
        # sqrtm take ~10 s; the worker lifts the limit for its own calls (threadpoolctl) to this many threads.
        import os

        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        self.fid_threads = fid_threads if fid_threads is not None else max(1, (os.cpu_count() or 1) // max(1, local_world))
        # The host-side sqrtm (several seconds at d=2048) is the serial term once sampling is sharded: every rank
        # holds the all-reduced moments, so a batch-sharded candidate is finished by rank `fid_owner(candidate)` only
        # and the values are exchanged in `resolve` (one tiny all-reduce per batch of candidates).
        self.shard_fid = shard_fid
        self._build_stream: Optional[th.cuda.Stream] = None

    # ---- plan cache keyed by what the launch schedule depends on ----
    def _plan_for(self, cand, batch: int, overlapped: bool = False) -> SchedulePlan:
        """`overlapped`: build without waiting for the device (previous candidates may still be sampling): recording,
        weight packing of skipped-block variants and the graph capture run under a side stream, nothing executes."""
        active, per_step = resolve_candidate(cand, self.base_diffusion)
        key = (tuple(active.timestep_map), tuple(tuple(s) for s in per_step), batch, self.image_size, self.clip_denoised)
        plan = self._plans.get(key)
        if plan is None:
            while len(self._plans) >= max(1, self._max_cached):
                self._plans.pop(next(iter(self._plans)))
            kw = dict(image_size=self.image_size, clip_denoised=self.clip_denoised, cond_fn=self.cond_fn, pack_uint8=True)
            dev = self.model._device()
            if overlapped and dev.type == "cuda":
                if self._build_stream is None:
                    self._build_stream = th.cuda.Stream(device=dev)
                with th.cuda.stream(self._build_stream):  # H2D uploads of freshly packed operands do not queue behind sampling
                    plan = SchedulePlan(self.model, active, per_step, batch, no_sync=True, **kw)
                th.cuda.current_stream().wait_stream(self._build_stream)
            else:
                plan = SchedulePlan(self.model, active, per_step, batch, **kw)
            self._plans[key] = plan
        return plan

    def candidate_cost(self, cand) -> float:
        """Estimated work of one image of `cand` in FLOP (F(cand) = sum over its steps of the forward's
        FLOPs minus what its skip list removes; the guidance pass adds the same amount to every step)."""
        _, per_step = resolve_candidate(cand, self.base_diffusion)
        ff = getattr(self.model, "forward_flops", None)
        if not callable(ff):
            return float(len(per_step))
        guidance = 0.37 * ff(()) if self.cond_fn is not None else 0.0  # classifier fwd + input gradient ~ 80.8 / 219.4
        return float(sum(ff(s, self.image_size, self.image_size) + guidance for s in per_step))

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
        owns; with more than one rank the values are exchanged by one all-reduce of 2 x len(futures) doubles (value and
        the number of ranks that contributed it: exactly one, or the ranks disagree about who owns a candidate). Every
        rank must call this with the futures of the same candidates - they do: all ranks walk the same population."""
        vals = [0.0 if isinstance(f, _RemoteFid) else float(f.result()) for f in futures]
        if any(isinstance(f, _RemoteFid) for f in futures) or (self.shard_fid and self.world_size > 1 and futures):
            mine = [0.0 if isinstance(f, _RemoteFid) else 1.0 for f in futures]
            nccl = dist.get_backend(self.group) == "nccl"
            t = th.tensor(vals + mine, dtype=th.float64, device=self.model._device() if nccl else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            out = t.cpu().tolist()
            vals, owners = out[:len(futures)], out[len(futures):]
            bad = [i for i, c in enumerate(owners) if c != 1.0]
            if bad:
                raise RuntimeError(f"FID exchange: candidates {bad} were finished by {[owners[i] for i in bad]} ranks instead "
                                   "of exactly one (ranks disagree on candidate ownership)")
        return vals

    def _apply_args(self, args):
        if args is not None:
            for k in ("batch_size", "num_samples", "image_size", "class_cond", "clip_denoised"):
                if hasattr(args, k):
                    setattr(self, k, getattr(args, k))

    def _accumulator(self, dim: int) -> MomentAccumulator:
        if self._acc is None or self._acc.dim != dim:
            self._acc = MomentAccumulator(dim, self.model._device())
        return self._acc

    def _sample_into(self, plan: SchedulePlan, cand_key: str, batches, own: bool = False) -> MomentAccumulator:
        """Enqueue sampling + features + moment accumulation of `batches`; nothing here waits for the device.
        own: accumulate into a buffer of this candidate's own (it outlives the next candidate's sampling) instead of the
        evaluator's reused one."""
        new_acc = (lambda d: MomentAccumulator(d, self.model._device())) if own else self._accumulator
        acc = None
        for b in batches:
            images, _ = self.sample_batch(plan, cand_key, b)
            keep = min(self.batch_size, self.num_samples - b * self.batch_size)  # arr[:num_samples], :432-433
            feats = self.feature_fn(images[:keep])
            if acc is None:
                acc = new_acc(feats.shape[1])
                acc.reset()
            acc.add(feats)
        if acc is None:  # this rank had no batch of this candidate
            acc = new_acc(self.ref_stats.mu.shape[0])
            acc.reset()
        return acc

    def _finish_async(self, acc: MomentAccumulator, times: dict, sync: bool, t0: float) -> Future:
        """D2H of the moment buffer (pinned, async) and the host-side statistics / Frechet distance on the worker."""
        host = self._host_bufs.pop() if self._host_bufs and self._host_bufs[-1].numel() == acc.buf.numel() \
            else th.empty(acc.buf.numel(), dtype=th.float64).pin_memory()
        host.copy_(acc.buf, non_blocking=True)
        done = th.cuda.Event()
        done.record()
        if sync:
            th.cuda.current_stream().synchronize()  # sample_time as the reference logs it (:435)
            times["sample_time"] = time.time() - t0
        dim = acc.dim

        def finish() -> float:
            done.synchronize()
            t1 = time.time()
            if not sync:
                times["sample_time"] = t1 - t0  # includes queueing behind earlier candidates of a pipelined population
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

    def evaluate_population(self, cands, args=None) -> list:
        """FID of every candidate in `cands`, on every rank, with the POPULATION sharded over ranks (BASELINE
        configs[2]). `schedule_population` places whole candidates longest-first (no per-candidate collective: plan
        building, sampling, moments and the host FID of such a candidate stay on one rank) and splits the `n mod world`
        tail at batch granularity, each tail candidate merged by ONE all-reduce of its [n | sum_x | sum_xx] buffer
        (NCCL over NVLink). Inside a rank the launch plan of candidate i+1 is recorded and captured while candidate i
        is still sampling (nothing in the build waits for the device). The FID values are exchanged by one all-reduce
        at the end. Images are those of the single-rank path (seeds depend on (seed, candidate, batch index) only).
        Every rank must pass the same list."""
        self._apply_args(args)
        cands = list(cands)
        n = len(cands)
        num_batches = (self.num_samples + self.batch_size - 1) // self.batch_size
        costs = [self.candidate_cost(c) for c in cands]
        whole, shared = schedule_population(costs, num_batches, self.world_size)
        mine = whole[self.rank]
        todo = [(i, range(num_batches), False) for i in mine] + [(i, per.get(self.rank, []), True) for i, per in shared]
        futures: list = [None] * n
        for r, idxs in whole.items():
            if r != self.rank:
                for i in idxs:
                    futures[i] = _RemoteFid(r)
        t_start = time.time()
        build_s, allreduce_ms = 0.0, []
        t0 = time.time()
        # a rank without a batch of a shared candidate only joins its all-reduce: no plan needed
        build = lambda k, **kw: self._plan_for(cands[todo[k][0]], self.batch_size, **kw) if len(todo[k][1]) else None
        nxt = build(0) if todo else None
        build_first = time.time() - t0
        tail = []
        for j, (i, batches, is_shared) in enumerate(todo):
            plan, cand_key = nxt, str(cands[i])
            times = dict(reset_time=0.0, sample_time=0.0, fid_time=0.0)
            t0 = time.time()
            acc = self._sample_into(plan, cand_key, batches, own=is_shared)
            if is_shared:
                tail.append((i, acc, cand_key, times, t0))
            else:
                futures[i] = self._finish_async(acc, times, sync=False, t0=t0)
            if j + 1 < len(todo):  # host-side build of the next candidate overlaps this one's sampling on the device
                t1 = time.time()
                nxt = build(j + 1, overlapped=True)
                build_s += time.time() - t1
        # The path's data-plane collective: one all-reduce of each tail candidate's moments, issued only after this rank
        # has enqueued ALL of its sampling. (An all-reduce right after the candidate's own batches made every rank wait
        # there for the most loaded one - which holds none of those batches - before it could start its share of the
        # next tail candidate: 1.2 s of a 26 s population at 8 ranks.) Same order on every rank: `shared` is.
        for (i, acc, cand_key, times, t0) in tail:
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            acc.all_reduce(self.group)
            e1.record()
            allreduce_ms.append((e0, e1))
            owner = fid_owner(cand_key, self.world_size)
            if owner == self.rank:
                futures[i] = self._finish_async(acc, times, sync=False, t0=t0)
            else:
                futures[i] = _RemoteFid(owner)
        vals = self.resolve(futures)
        th.cuda.current_stream().synchronize() if th.cuda.is_available() else None
        self.last_population = dict(
            candidates=n, whole_per_rank=[len(whole[r]) for r in range(self.world_size)], shared=len(shared),
            plan_build_first_s=build_first, plan_build_overlapped_s=build_s, wall_s=time.time() - t_start,
            allreduce_ms=[a.elapsed_time(b) for a, b in allreduce_ms], est_cost_min_max=(min(costs), max(costs)) if costs else None)
        return vals

    def submit_cand_fid(self, cand=None, args=None, _whole_on: Optional[int] = None) -> Future:
        """Same work as `get_cand_fid`, but only the device part (sampling, moments, all-reduce, D2H of the
        33.6 MB moment buffer) happens before this returns; mu / sigma / sqrtm run on a host worker thread.
        The search driver keeps sampling the next candidate meanwhile (`fid_time` was serial in the
        reference, :437-443). `_whole_on=r`: rank r alone samples and scores the candidate (no collective)."""
        self._apply_args(args)
        if _whole_on is not None and _whole_on != self.rank:  # another rank evaluates this candidate entirely
            self.last_times = dict(reset_time=0.0, sample_time=0.0, fid_time=0.0)
            return _RemoteFid(_whole_on)
        t0 = time.time()
        plan = self._plan_for(cand, self.batch_size)
        times = dict(reset_time=time.time() - t0, sample_time=0.0, fid_time=0.0)
        self.last_times = times
        t0 = time.time()
        cand_key = str(cand)
        num_batches = (self.num_samples + self.batch_size - 1) // self.batch_size
        my_batches = range(num_batches) if _whole_on is not None else shard_batches(num_batches, self.rank, self.world_size)
        acc = self._sample_into(plan, cand_key, my_batches)
        if _whole_on is None:
            acc.all_reduce(self.group)
        owner = fid_owner(cand_key, self.world_size)
        if _whole_on is None and self.shard_fid and self.world_size > 1 and owner != self.rank:
            th.cuda.current_stream().synchronize() if acc.buf.is_cuda else None
            times["sample_time"] = time.time() - t0
            return _RemoteFid(owner)
        return self._finish_async(acc, times, sync=True, t0=t0)

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
