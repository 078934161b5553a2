"""Whole-candidate sampling plans: a searched DDIM schedule as one recorded launch sequence.

A candidate (…progressive.py:369-373) is `{'timesteps': [...], 'skip_layers': [[...], ...]}`.
The reference evaluates it through closures and a Python loop
(`model_fn`/`cond_fn` :383-397, `ddim_sample_loop` gaussian_diffusion.py:664-716): per step one
GPU->CPU sync (`timestep_map.index(t[0])`), ~8 host->device table uploads and ~25 elementwise
launches around an ~800-launch eager UNet forward.

`SchedulePlan` resolves all of that when the candidate arrives:
  * timesteps -> `set` -> ascending `timestep_map` and float64 tables (respace.reset_diffusion),
  * the skip list of step i is `skip_layers[i]` for the i-th SMALLEST timestep (the evaluator's
    sorted-rank indexing, :394-396), entries beyond K' unused,
  * every step's UNet forward (skipped blocks elided) + fused guidance/DDIM update is recorded
    into ONE plan, in sampling order K'-1 .. 0, and captured in ONE CUDA graph when there is no
    caller-supplied `cond_fn`. With a `cond_fn` (an arbitrary torch callable returning
    grad log p(y|x) * scale) the graph is cut at each step around that call.
  * the final `((x+1)*127.5).clamp(0,255).to(uint8)` NHWC pack (:421-423) is the last node.
"""
from __future__ import annotations

import copy
import os
from typing import Callable, List, Optional, Sequence

import torch as th

from . import ops
from .gaussian_diffusion import ModelMeanType, ddim_coefficients
from .respace import reset_diffusion


def resolve_candidate(cand, base_diffusion, active_diffusion=None):
    """-> (active diffusion rebuilt for the candidate, per-step skip lists in timestep_map order).

    `cand` is a dict with 'timesteps' and optional 'skip_layers', or a bare list of timesteps
    (search_imagenet64_classifier_guidance.py:308-311)."""
    if isinstance(cand, dict):
        timesteps = list(cand["timesteps"])
        skip_layers = cand.get("skip_layers")
    else:
        timesteps, skip_layers = list(cand), None
    active = active_diffusion if active_diffusion is not None else copy.deepcopy(base_diffusion)
    reset_diffusion(timesteps, active, base_diffusion)
    k = active.num_timesteps
    if skip_layers is None:
        per_step = [[] for _ in range(k)]
    else:
        if len(skip_layers) < k:
            raise IndexError(f"candidate has {len(skip_layers)} skip lists for {k} distinct timesteps")
        per_step = [sorted(set(int(s) for s in skip_layers[i])) for i in range(k)]  # sorted-rank indexing
    return active, per_step


class SchedulePlan:
    """All K' steps of a candidate for a fixed batch, recorded once and replayed per batch."""

    def __init__(self, model, active_diffusion, per_step_skips: Sequence[Sequence[int]], batch: int,
                 image_size: Optional[int] = None, clip_denoised: bool = True, cond_fn: Optional[Callable] = None,
                 pack_uint8: bool = True, use_graph: Optional[bool] = None):
        if active_diffusion.model_mean_type != ModelMeanType.EPSILON:
            raise NotImplementedError("SchedulePlan covers epsilon-prediction models")
        if active_diffusion.rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True is not used by any reference config")
        dev = model._device()
        if dev.type != "cuda":
            raise RuntimeError("SchedulePlan needs the model on a CUDA device (no CPU path)")
        self.model = model
        self.diffusion = active_diffusion
        self.timestep_map = list(active_diffusion.timestep_map)
        self.per_step_skips = [list(s) for s in per_step_skips]
        self.K = active_diffusion.num_timesteps
        assert len(self.per_step_skips) == self.K
        self.B = batch
        hw = image_size or model.image_size
        self.shape = (batch, model.in_channels, hw, hw)
        self.cond_fn = cond_fn
        self.class_cond = model.num_classes is not None
        if use_graph is None:
            use_graph = os.environ.get("ADB_NO_GRAPH", "0") != "1"

        f32 = dict(dtype=th.float32, device=dev)
        self.x = [th.zeros(self.shape, **f32), th.zeros(self.shape, **f32)]  # ping-pong x_t
        self.y = th.zeros((batch,), dtype=th.int64, device=dev) if self.class_cond else None
        self.model_out = th.empty((batch, model.out_channels, hw, hw), **f32)
        self.grad = th.zeros(self.shape, **f32) if cond_fn is not None else None
        self.t_tensors = [th.full((batch,), int(t), dtype=th.int64, device=dev) for t in self.timestep_map]
        self.u8 = th.empty((batch, hw, hw, model.in_channels), dtype=th.uint8, device=dev) if pack_uint8 else None

        # segments: with a cond_fn each step is [unet] -> cond_fn (eager torch) -> [ddim_step]
        self.segments: List[ops.Plan] = []
        self.graphs: List[Optional[th.cuda.CUDAGraph]] = []
        self.launches = 0
        cur = 0
        order = list(range(self.K))[::-1]  # sampling runs high -> low (gaussian_diffusion.py:690)
        with th.no_grad():
            plan = ops.Plan()
            for n, i in enumerate(order):
                model.record_forward(plan, self.x[cur], self.t_tensors[i], self.y, self.model_out,
                                     self.per_step_skips[i])
                if cond_fn is not None:
                    self.segments.append(plan)
                    plan = ops.Plan()
                ops.ddim_step(self.x[cur], self.model_out, self.grad, ddim_coefficients(active_diffusion, i),
                              clip_denoised, x_prev=self.x[1 - cur], plan=plan)
                cur = 1 - cur
            self.final = self.x[cur]
            if self.u8 is not None:
                ops.pack_uint8(self.final, out=self.u8, plan=plan)
            self.segments.append(plan)
            self._order = order
            # validation run (sets kernel attributes) + capture
            if cond_fn is None:
                for seg in self.segments:
                    self.launches += seg.run()
                self.graphs = [self._capture(seg) if use_graph else None for seg in self.segments]
            else:
                self.graphs = [None] * len(self.segments)
                self._use_graph = use_graph
                self._captured = False

    @staticmethod
    def _capture(seg: ops.Plan) -> th.cuda.CUDAGraph:
        th.cuda.current_stream().synchronize()
        g = th.cuda.CUDAGraph()
        with th.cuda.graph(g):
            seg.run()
        return g

    def _run_segment(self, k: int):
        if self.graphs[k] is not None:
            self.graphs[k].replay()
        else:
            n = self.segments[k].run()
            if self.cond_fn is not None and not self._captured:
                self.launches += n

    def run(self, noise: th.Tensor, y: Optional[th.Tensor] = None, model_kwargs: Optional[dict] = None) -> th.Tensor:
        """x_T = noise (fp32 [B,C,H,W]), labels y -> x_0 (a view of an internal buffer; clone to keep)."""
        assert tuple(noise.shape) == self.shape
        self.x[0].copy_(noise, non_blocking=True)
        if self.class_cond:
            assert y is not None and y.shape == (self.B,)
            self.y.copy_(y, non_blocking=True)
        if self.cond_fn is None:
            self._run_segment(0)
            self.model.gpu_launches += self.launches
            return self.final
        kwargs = dict(model_kwargs or {})
        if self.class_cond:
            kwargs.setdefault("y", self.y)
        cur = 0
        for n, i in enumerate(self._order):
            self._run_segment(n)  # UNet forward of step i (and the previous step's DDIM update)
            g = self.cond_fn(self.x[cur], self.t_tensors[i], **kwargs)  # original timestep, as _WrappedModel passes it
            self.grad.copy_(g)
            cur = 1 - cur
        self._run_segment(len(self._order))
        if not self._captured:
            self._captured = True
            if self._use_graph:
                self.graphs = [self._capture(seg) for seg in self.segments]
        return self.final


def sample_candidate(model, base_diffusion, cand, shape, noise: th.Tensor, y: Optional[th.Tensor] = None,
                     clip_denoised: bool = True, cond_fn: Optional[Callable] = None) -> th.Tensor:
    """One-shot convenience: build the plan for `cand` and sample one batch (fp32 NCHW result)."""
    active, per_step = resolve_candidate(cand, base_diffusion)
    plan = SchedulePlan(model, active, per_step, shape[0], image_size=shape[2], clip_denoised=clip_denoised,
                        cond_fn=cond_fn, pack_uint8=False)
    return plan.run(noise, y).clone()
