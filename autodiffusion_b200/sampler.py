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
  * every step's UNet forward (skipped blocks elided; one cached CUDA graph per (batch, skip set)) +
    fused guidance/DDIM update is chained in sampling order K'-1 .. 0 and captured in ONE CUDA graph
    when there is no caller-supplied `cond_fn`, or when the `cond_fn` is this package's
    `classifier.ClassifierGuidance` (the noisy classifier's forward + input-gradient is then recorded
    into the same graph, one instance per step). With any other `cond_fn` (an arbitrary torch callable
    returning grad log p(y|x) * scale) the chain is run step by step around that call.
  * the final `((x+1)*127.5).clamp(0,255).to(uint8)` NHWC pack (:421-423) is the last node.
"""
from __future__ import annotations

import copy
import os
from typing import Callable, List, Optional, Sequence

import torch as th

from . import ops
from .classifier import ClassifierGuidance
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
    """All K' steps of a candidate for a fixed batch: cached per-mask UNet graphs chained into one graph.

    Building one is cheap: a UNet forward is recorded (and graph-captured) once per (batch, skip set)
    and cached on the model - most steps of most candidates share the empty mask - so a new candidate
    only costs its K' coefficient sets, K' tiny nodes and the capture of the chain. All forwards use the
    model's shared input/output buffers: the fused DDIM update writes x_{t-1} in place into the UNet's
    input buffer.
    """

    def __init__(self, model, active_diffusion, per_step_skips: Sequence[Sequence[int]], batch: int,
                 image_size: Optional[int] = None, clip_denoised: bool = True, cond_fn: Optional[Callable] = None,
                 pack_uint8: bool = True, use_graph: Optional[bool] = None, no_sync: bool = False):
        """`no_sync`: build without waiting for (or adding work to) the device - no validation run of freshly recorded
        forwards, graph capture on the current (side) stream without the device-wide synchronize of `torch.cuda.graph`.
        Used by `CandidateEvaluator.evaluate_population` to plan candidate i+1 while candidate i samples."""
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
        self.clip_denoised = clip_denoised
        if use_graph is None:
            use_graph = os.environ.get("ADB_NO_GRAPH", "0") != "1"
        self.use_graph = use_graph

        self._order = list(range(self.K))[::-1]  # sampling runs high -> low (gaussian_diffusion.py:690)
        with th.no_grad():
            self.steps = [model.get_plan(batch, hw, hw, self.per_step_skips[i], validate=not no_sync) for i in self._order]
        io = model.io_buffers(batch, hw, hw)
        self.x, self.t_in, self.y, self.model_out = io.x_in, io.t_in, io.y_in, io.out
        self.final = self.x  # x_0 ends up in the shared input buffer
        self.coefs = [ddim_coefficients(active_diffusion, i) for i in self._order]
        self.t_values = [int(self.timestep_map[i]) for i in self._order]
        self.grad = None
        if cond_fn is not None:  # one guidance-gradient buffer per model geometry, shared like the IO buffers
            grads = model.__dict__.setdefault("_guidance_grad", {})
            key = (self.shape, str(dev))
            if key not in grads:
                grads[key] = th.zeros(self.shape, dtype=th.float32, device=dev)
            self.grad = grads[key]
        self.u8 = th.empty((batch, hw, hw, model.in_channels), dtype=th.uint8, device=dev) if pack_uint8 else None
        self._count_launches()
        # native classifier guidance: forward + input-gradient recorded once over the shared x / t / y buffers
        self.guidance: Optional[ops.Plan] = None
        if isinstance(cond_fn, ClassifierGuidance):
            if self.y is None:  # unconditional UNet guided by a classifier: labels still drive the guidance
                self.y = th.zeros((batch,), dtype=th.int64, device=dev)
            self.guidance = cond_fn.shared_plan(self.x, self.t_in, self.y, self.grad)
            self._count_launches()
        # second stream for the guidance branch of each step (created here, outside any capture)
        self._side = th.cuda.Stream(device=dev) if self.guidance is not None else None
        self.graph: Optional[th.cuda.CUDAGraph] = None
        if (cond_fn is None or self.guidance is not None) and use_graph:
            g = th.cuda.CUDAGraph()
            if no_sync:
                # capture on the caller's side stream: nothing executes and nothing inside allocates, so neither the
                # device-wide synchronize nor the private memory pool handling of `torch.cuda.graph` is needed
                if th.cuda.current_stream() == th.cuda.default_stream():
                    raise RuntimeError("SchedulePlan(no_sync=True) must be built under a non-default torch.cuda.stream")
                g.capture_begin(capture_error_mode="thread_local")
                try:
                    self._run_chain()
                finally:
                    g.capture_end()
            else:
                th.cuda.current_stream().synchronize()
                with th.cuda.graph(g):
                    self._run_chain()
            self.graph = g
            self._count_launches()  # plans built without a validation run learnt their launch counts during capture

    def _count_launches(self):
        self.launches = sum(up.launches for up in self.steps) + self.K + (1 if self.u8 is not None else 0)
        if getattr(self, "guidance", None) is not None:
            self.launches += self.K * self.guidance.launches_per_run

    def _step(self, n: int, fill: bool = True):
        """UNet forward of the n-th sampled step (cached graph as a child node when capturing)."""
        if fill:
            self.t_in.fill_(self.t_values[n])  # original timestep, as _WrappedModel maps it (respace.py:122-127)
        if th.cuda.is_current_stream_capturing():
            self.steps[n].launches = self.steps[n].plan.run()  # re-issue the recorded launches into the schedule's own graph
        else:
            self.steps[n].replay()

    def _update(self, n: int):
        ops.ddim_step(self.x, self.model_out, self.grad, self.coefs[n], self.clip_denoised, x_prev=self.x)

    def _run_chain(self):
        # eps(x_t, t) and grad log p(y | x_t) both read x_t / t only and write different buffers (model_out / grad) from
        # private activation pools: the guidance plan is issued on a second stream, forked after t is set and joined
        # before the DDIM update, so that its kernels fill the tails of the UNet's persistent kernels (and vice versa):
        # +2 % images/s, bit-identical samples. Inside a capture the fork / join become graph edges.
        # ADB_CONCURRENT_GUIDANCE=0 restores the single-stream order (A/B testing).
        conc = self.guidance is not None and os.environ.get("ADB_CONCURRENT_GUIDANCE", "1") != "0"
        for n in range(self.K):
            if conc:
                main = th.cuda.current_stream()
                self.t_in.fill_(self.t_values[n])
                fork, join = th.cuda.Event(), th.cuda.Event()
                fork.record(main)
                self._side.wait_event(fork)
                with th.cuda.stream(self._side):
                    self.guidance.run()
                    join.record(self._side)
                self._step(n, fill=False)
                main.wait_event(join)
            else:
                self._step(n)
                if self.guidance is not None:
                    self.guidance.run()  # grad log p(y | x_t) * scale at the ORIGINAL timestep (t_in), into self.grad
            self._update(n)
        if self.u8 is not None:
            ops.pack_uint8(self.final, out=self.u8)

    def run(self, noise: th.Tensor, y: Optional[th.Tensor] = None, model_kwargs: Optional[dict] = None) -> th.Tensor:
        """x_T = noise (fp32 [B,C,H,W]), labels y -> x_0 (a view of the shared buffer; clone to keep)."""
        assert tuple(noise.shape) == self.shape
        self.x.copy_(noise, non_blocking=True)
        if self.class_cond or self.guidance is not None:
            assert y is not None and y.shape == (self.B,)
            self.y.copy_(y, non_blocking=True)
        self.model.gpu_launches += self.launches
        if self.cond_fn is None or self.guidance is not None:
            if self.graph is not None:
                self.graph.replay()
            else:
                self._run_chain()
            return self.final
        kwargs = dict(model_kwargs or {})
        if self.class_cond:
            kwargs.setdefault("y", self.y)
        for n in range(self.K):
            self._step(n)
            # caller-supplied guidance: grad log p(y|x_t) * scale at the ORIGINAL timestep
            g = self.cond_fn(self.x, self.t_in, **kwargs)
            self.grad.copy_(g)
            self._update(n)
        if self.u8 is not None:
            ops.pack_uint8(self.final, out=self.u8)
        return self.final


def sample_candidate(model, base_diffusion, cand, shape, noise: th.Tensor, y: Optional[th.Tensor] = None,
                     clip_denoised: bool = True, cond_fn: Optional[Callable] = None) -> th.Tensor:
    """One-shot convenience: build the plan for `cand` and sample one batch (fp32 NCHW result)."""
    active, per_step = resolve_candidate(cand, base_diffusion)
    plan = SchedulePlan(model, active, per_step, shape[0], image_size=shape[2], clip_denoised=clip_denoised,
                        cond_fn=cond_fn, pack_uint8=False)
    return plan.run(noise, y).clone()
