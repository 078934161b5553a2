"""Searched-timestep DDIM sampling with classifier-free guidance for the Stable-Diffusion family.

Drop-in for `ldm.models.diffusion.ddim.DDIMSampler` as the search calls it (reference:
/root/reference/examples/"Stable Diffusion"/ldm/models/diffusion/ddim.py:13-217 and scripts/search_ea.py:504-538):
`DDIMSampler(model).sample(S, batch_size, shape, conditioning, eta=0, x_T, unconditional_guidance_scale,
unconditional_conditioning, sampled_timestep)` -> (samples, intermediates). `model` is anything with the attributes
the reference sampler reads from LatentDiffusion (num_timesteps, betas, alphas_cumprod, alphas_cumprod_prev, device,
apply_model); `LatentDiffusionUNet` below provides them around `sd_unet.UNetModel` with the v1 schedule.

When the model is ours, a whole candidate - context K/V projections once, then per searched step one batched
[uncond | cond] UNet forward and the fused CFG + DDIM update - is recorded into one launch plan and captured in
one CUDA graph (`CandidatePlan`); any other `apply_model` goes through the generic loop with the same fused update.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch as th

from . import ops
from .sd_unet import CTX_ROWS, UNetModel


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """ldm/modules/diffusionmodules/util.py:21-43 ("linear" is what v1-inference*.yaml uses)."""
    if schedule != "linear":
        raise NotImplementedError(f"schedule '{schedule}' is not used by the Stable-Diffusion search")
    betas = th.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=th.float64) ** 2
    return betas.numpy()


def ddim_tables(alphas_cumprod: th.Tensor, ddim_timesteps: Sequence[int]):
    """make_ddim_sampling_parameters (util.py:63-75) at eta = 0, as fp32 tensors:
    alphas = acp[steps], alphas_prev = [acp[0]] + acp[steps[:-1]], sqrt(1 - alphas)."""
    steps = [int(t) for t in ddim_timesteps]
    acp = alphas_cumprod.detach().float().cpu()
    alphas = acp[steps]
    alphas_prev = th.tensor([acp[0].item()] + acp[steps[:-1]].tolist(), dtype=th.float32)
    return alphas, alphas_prev, th.sqrt(1.0 - alphas)


def ddim_coefficients(alphas, alphas_prev, sqrt_one_minus_alphas, index: int):
    """The four fp32 scalars p_sample_ddim broadcasts at `index` (ddim.py:199-213, sigma_t = 0):
    sqrt(1 - a_t), sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev) - each rounded to fp32 like the reference's tensor ops."""
    a_t = np.float32(alphas[index])
    a_prev = np.float32(alphas_prev[index])
    return (float(np.float32(sqrt_one_minus_alphas[index])), float(np.sqrt(a_t)), float(np.sqrt(a_prev)),
            float(np.sqrt(np.float32(1.0) - a_prev)))


class LatentDiffusionUNet:
    """What DDIMSampler needs from LatentDiffusion (ldm/models/diffusion/ddpm.py:117-134 register_schedule;
    apply_model -> DiffusionWrapper 'crossattn' -> unet(x, t, context=c), ddpm.py:1410-1417) around our UNet.
    The text encoder and the VAE are outside the searched path (scripts/search_ea.py:519-525,539)."""

    def __init__(self, unet: UNetModel, timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.0120):
        self.unet = unet
        betas = make_beta_schedule("linear", timesteps, linear_start, linear_end)
        acp = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = int(timesteps)
        self.parameterization = "eps"
        dev = unet._device()
        self.betas = th.tensor(betas, dtype=th.float32, device=dev)
        self.alphas_cumprod = th.tensor(acp, dtype=th.float32, device=dev)
        self.alphas_cumprod_prev = th.tensor(np.append(1.0, acp[:-1]), dtype=th.float32, device=dev)

    @property
    def device(self):
        return self.unet._device()

    def apply_model(self, x_noisy, t, cond):
        if isinstance(cond, dict):
            cond = th.cat(cond["c_crossattn"], 1)
        return self.unet(x_noisy, t, context=cond)


class _SharedForward:
    """Buffers and recorded plans every candidate of one (model, batch, latent shape, CFG, context length, timestep dtype)
    shares: a UNet forward does not depend on the searched schedule (the timestep is a device buffer), so a new candidate
    only costs its coefficient sets and the capture of its own chain."""

    def __init__(self, unet: UNetModel, n: int, shape, ctx_tokens: int, t_dtype):
        dev = unet._device()
        C, H, W = shape
        self.x2 = th.zeros((n, C, H, W), dtype=th.float32, device=dev)
        self.t_in = th.zeros((n,), dtype=t_dtype, device=dev)
        self.ctx = th.zeros((n, ctx_tokens, unet.context_dim), dtype=th.float32, device=dev)  # [uncond | cond] under CFG
        self.eps = th.empty((n, unet.out_channels, H, W), dtype=th.float32, device=dev)
        self.plan_ctx = ops.Plan()
        cpad = ops.pad_context(self.ctx, CTX_ROWS, plan=self.plan_ctx)
        kvs = unet.record_context(self.plan_ctx, cpad)
        self.plan_fwd = ops.Plan()  # one forward; every step of every candidate replays it
        unet.record_forward(self.plan_fwd, self.x2, self.t_in, kvs, self.eps, ctx_tokens=ctx_tokens)
        with th.no_grad():  # first run outside any capture: sets kernel attributes, validates the schedule
            self.ctx_launches = self.plan_ctx.run()
            self.fwd_launches = self.plan_fwd.run()


def shared_forward(unet: UNetModel, n: int, shape, ctx_tokens: int, t_dtype) -> _SharedForward:
    unet._ready()  # re-packs (and clears unet._plans) after a weight change
    key = ("shared_forward", n, tuple(shape), ctx_tokens, t_dtype)
    sf = unet._plans.get(key)
    if sf is None:
        sf = _SharedForward(unet, n, shape, ctx_tokens, t_dtype)
        unet._plans[key] = sf
    return sf


class CandidatePlan:
    """One searched candidate (sorted timesteps) as ONE recorded schedule / CUDA graph at a fixed batch:
    pad + project the contexts once, then for every step (descending): t, [x | x] -> eps_uncond, eps_cond -> fused
    CFG + DDIM update written back into both halves of the UNet's input buffer."""

    def __init__(self, unet: UNetModel, alphas_cumprod: th.Tensor, sampled_timestep: Sequence[int], batch: int, shape,
                 scale: float, cfg: bool, ctx_tokens: int = 77, use_graph: Optional[bool] = None, method: str = "ddim"):
        dev = unet._device()
        if dev.type != "cuda":
            raise RuntimeError("CandidatePlan needs the model on a CUDA device (no CPU path)")
        C, H, W = shape
        assert method in ("ddim", "plms")
        self.method = method
        self.unet, self.B, self.cfg, self.scale = unet, batch, bool(cfg), float(scale)
        self.steps = sorted(int(t) for t in sampled_timestep)  # ddim.py:93-94
        self.alphas, self.alphas_prev, self.s1m = ddim_tables(alphas_cumprod, self.steps)
        n = batch * (2 if cfg else 1)
        sf = shared_forward(unet, n, (C, H, W), ctx_tokens, th.int64)
        self.x2, self.t_in, self.ctx, self.eps, self.plan_ctx, self.plan_fwd = sf.x2, sf.t_in, sf.ctx, sf.eps, sf.plan_ctx, sf.plan_fwd
        self.x = self.x2[:batch]
        if method == "plms":  # eps history ring (p_sample_plms keeps the last three), e_t_next and x_t of the first step
            self.e_ring = [th.empty((batch, C, H, W), dtype=th.float32, device=dev) for _ in range(4)]
            self.e_next = th.empty((batch, C, H, W), dtype=th.float32, device=dev)
            self.x_keep = th.empty((batch, C, H, W), dtype=th.float32, device=dev)
        self.coefs = [ddim_coefficients(self.alphas, self.alphas_prev, self.s1m, i) for i in range(len(self.steps))]
        self.launches, per_fwd = sf.ctx_launches, sf.fwd_launches
        self.launches += len(self.steps) * (per_fwd + (1 if method == "ddim" else 2)) + (per_fwd + 2 if method == "plms" else 0)
        self.launches_per_forward = per_fwd
        if use_graph is None:
            use_graph = os.environ.get("ADB_NO_GRAPH", "0") != "1"
        self.graph: Optional[th.cuda.CUDAGraph] = None
        if use_graph:
            th.cuda.current_stream().synchronize()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                self._chain()
            self.graph = g

    def _dup(self):
        if self.cfg:
            self.x2[self.B:].copy_(self.x)

    def _chain_plms(self):
        """plms.py:152-187 + p_sample_plms (:190-257)."""
        self.plan_ctx.run()
        time_range = list(reversed(self.steps))
        for i, step in enumerate(time_range):
            index = len(self.steps) - i - 1
            self.t_in.fill_(step)
            self.plan_fwd.run()
            e_t = ops.cfg_combine(self.eps, scale=self.scale, cfg=self.cfg, out=self.e_ring[i % 4])
            if i == 0:  # pseudo improved Euler: a provisional x_prev, the model again at t_next, then the real update from x_t
                self.x_keep.copy_(self.x)
                ops.plms_update(self.x, e_t, [], 0, self.coefs[index], x_prev=self.x)
                self._dup()
                self.t_in.fill_(time_range[min(i + 1, len(time_range) - 1)])
                self.plan_fwd.run()
                ops.cfg_combine(self.eps, scale=self.scale, cfg=self.cfg, out=self.e_next)
                ops.plms_update(self.x_keep, e_t, [self.e_next], 1, self.coefs[index], x_prev=self.x)
            else:
                n_old = min(i, 3)
                olds = [self.e_ring[(i - k) % 4] for k in range(1, n_old + 1)]  # newest first
                ops.plms_update(self.x, e_t, olds, 1 + n_old, self.coefs[index], x_prev=self.x)
            self._dup()

    def _chain(self):
        if self.method == "plms":
            return self._chain_plms()
        self.plan_ctx.run()
        for i, step in enumerate(reversed(self.steps)):
            index = len(self.steps) - i - 1
            self.t_in.fill_(step)
            self.plan_fwd.run()
            ops.cfg_ddim_step(self.x, self.eps, self.coefs[index], scale=self.scale, cfg=self.cfg, x_prev=self.x)
            if self.cfg:
                self.x2[self.B:].copy_(self.x)

    @th.no_grad()
    def run(self, x_T: th.Tensor, cond: th.Tensor, uncond: Optional[th.Tensor]) -> th.Tensor:
        """-> x_0 latents (a view of the plan's buffer; clone to keep)."""
        self.x.copy_(x_T, non_blocking=True)
        if self.cfg:
            self.x2[self.B:].copy_(x_T, non_blocking=True)
            self.ctx[:self.B].copy_(uncond, non_blocking=True)
            self.ctx[self.B:].copy_(cond, non_blocking=True)
        else:
            self.ctx.copy_(cond, non_blocking=True)
        self.unet.gpu_launches += self.launches
        if self.graph is not None:
            self.graph.replay()
        else:
            self._chain()
        return self.x


class DDIMSampler(object):
    """ddim.py:13-217 restricted to what the search calls: eta = 0, no mask / x0 / score corrector / quantisation."""

    method = "ddim"

    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self._plans: Dict[tuple, CandidatePlan] = {}

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True, sampled_timestep=None):
        if ddim_eta != 0.0:
            raise NotImplementedError("eta != 0 is not used by the search (scripts/search_ea.py: ddim_eta default 0.0)")
        if sampled_timestep is None:  # make_ddim_timesteps, util.py:46-61
            if ddim_discretize != "uniform":
                raise NotImplementedError(ddim_discretize)
            c = round(self.ddpm_num_timesteps / ddim_num_steps)
            self.ddim_timesteps = np.asarray(list(range(0, self.ddpm_num_timesteps, c))) + 1
        else:
            self.ddim_timesteps = sampled_timestep
        acp = self.model.alphas_cumprod
        assert acp.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas = ddim_tables(acp, list(self.ddim_timesteps))
        self.ddim_sigmas = th.zeros_like(self.ddim_alphas)

    @th.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0, score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, sampled_timestep=None, **kwargs):
        if mask is not None or x0 is not None or score_corrector is not None or quantize_x0 or noise_dropout > 0.0:
            raise NotImplementedError("inpainting / score correction / quantisation are not on the searched path")
        if conditioning is not None and not isinstance(conditioning, dict) and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        if sampled_timestep is not None:
            sampled_timestep = sorted(int(t) for t in sampled_timestep)
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose, sampled_timestep=sampled_timestep)
        C, H, W = shape
        dev = self.model.device
        img = th.randn((batch_size, C, H, W), device=dev) if x_T is None else x_T
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.0)  # ddim.py:184
        unet = getattr(self.model, "unet", None)
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        if isinstance(unet, UNetModel) and not isinstance(conditioning, dict) and callback is None and img_callback is None:
            steps = [int(t) for t in self.ddim_timesteps]
            key = (self.method, tuple(steps), batch_size, C, H, W, float(unconditional_guidance_scale), cfg, conditioning.shape[1])
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 4:
                    self._plans.pop(next(iter(self._plans)))
                plan = CandidatePlan(unet, self.model.alphas_cumprod, steps, batch_size, (C, H, W),
                                     unconditional_guidance_scale, cfg, ctx_tokens=conditioning.shape[1], method=self.method)
                self._plans[key] = plan
            out = plan.run(img, conditioning, unconditional_conditioning).clone()
            return out, intermediates
        # generic apply_model: the reference's loop (ddim.py:143-172) with the fused update
        steps = np.asarray(self.ddim_timesteps)
        total = steps.shape[0]
        time_range = [int(t) for t in np.flip(steps)]

        def model_eps(x, ts):
            if cfg:
                return self.model.apply_model(th.cat([x] * 2), th.cat([ts] * 2),
                                              th.cat([unconditional_conditioning, conditioning])).float().contiguous()
            return self.model.apply_model(x, ts, conditioning).float().contiguous()

        old_eps: List[th.Tensor] = []
        for i, step in enumerate(time_range):
            index = total - i - 1
            ts = th.full((batch_size,), int(step), device=dev, dtype=th.long)
            coef = ddim_coefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas, index)
            pred = th.empty_like(img)
            img = img.contiguous()
            if self.method == "ddim":
                img = ops.cfg_ddim_step(img, model_eps(img, ts), coef, scale=unconditional_guidance_scale, cfg=cfg, pred_x0=pred)
            else:  # plms.py:190-257
                e_t = ops.cfg_combine(model_eps(img, ts), scale=unconditional_guidance_scale, cfg=cfg)
                if not old_eps:
                    ts_next = th.full((batch_size,), time_range[min(i + 1, total - 1)], device=dev, dtype=th.long)
                    x_prov = ops.plms_update(img, e_t, [], 0, coef)
                    e_next = ops.cfg_combine(model_eps(x_prov, ts_next), scale=unconditional_guidance_scale, cfg=cfg)
                    img = ops.plms_update(img, e_t, [e_next], 1, coef, pred_x0=pred)
                else:
                    img = ops.plms_update(img, e_t, old_eps[::-1], 1 + len(old_eps), coef, pred_x0=pred)
                old_eps.append(e_t)
                if len(old_eps) >= 4:
                    old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred, i)
            if index % log_every_t == 0 or index == total - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred)
        return img, intermediates


class PLMSSampler(DDIMSampler):
    """ldm/models/diffusion/plms.py:13-257 as the search calls it (scripts/search_ea.py with the PLMS sampler,
    search_plms.sh): same schedule tables as DDIM, pseudo linear multistep eps combinations, eta must be 0."""

    method = "plms"

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True, sampled_timestep=None):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")  # plms.py:25-26
        return super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose, sampled_timestep)


# ------------------------------------------------------------------------------------------
# DPM-Solver++(2M) with searched time steps (ldm/models/diffusion/dpm_solver/, search_dpm_solver.sh)
# ------------------------------------------------------------------------------------------
class DiscreteNoiseSchedule:
    """NoiseScheduleVP('discrete', alphas_cumprod=...) (dpm_solver.py:97-156): log alpha_t is piecewise linear in t over
    the knots t_n = n / N, n = 1..N (outermost segments extended); every quantity a fp32 torch op on the host, in the
    reference's order, so the per-step scalars equal the reference's."""

    def __init__(self, alphas_cumprod: th.Tensor):
        self.log_alpha = 0.5 * th.log(alphas_cumprod.detach().float().cpu())
        self.total_N = int(self.log_alpha.shape[0])
        self.T = 1.0
        self.t_array = th.linspace(0.0, 1.0, self.total_N + 1)[1:]

    def marginal_log_mean_coeff(self, t: th.Tensor) -> th.Tensor:
        t = t.contiguous()
        lo = th.clamp(th.searchsorted(self.t_array, t, right=False) - 1, 0, self.total_N - 2)
        x0, x1, y0, y1 = self.t_array[lo], self.t_array[lo + 1], self.log_alpha[lo], self.log_alpha[lo + 1]
        return y0 + (t - x0) * (y1 - y0) / (x1 - x0)

    def marginal_alpha(self, t):
        return th.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return th.sqrt(1.0 - th.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * th.log(1.0 - th.exp(2.0 * lm))


def dpm_time_steps(ea_timesteps, total_N: int = 1000) -> th.Tensor:
    """dpm_solver.py:1079-1091: integer candidates index the reversed uniform 1001-point grid between t_T = 1 and
    t_0 = 1/N in the order given; candidates already in (0, 1] are sorted descending."""
    ea = [float(v) for v in ea_timesteps]
    if max(ea) > 1:
        full = list(th.linspace(1.0, 1.0 / total_N, 1000 + 1))
        full.reverse()
        return th.Tensor([full[int(v)].item() for v in ea])
    return th.Tensor(sorted(ea, reverse=True))


def dpm_schedule(ns: DiscreteNoiseSchedule, ts: th.Tensor):
    """Everything the 2M solver needs per model evaluation / update as Python floats (fp32 values):
    model input times, (sigma, alpha) for the data prediction, and the update coefficients
    (order, c0, c1, c2, inv_r0) for steps 1..S (dpm_solver.py:519-533, 770-790, 1099-1121)."""
    S = ts.shape[0] - 1
    assert S >= 2, "the order-2 multistep solver needs at least 2 steps (dpm_solver.py:1077)"
    t_in = ((ts - 1.0 / ns.total_N) * 1000.0).tolist()  # get_model_input_time (:278-286)
    sig, alp = ns.marginal_std(ts).tolist(), ns.marginal_alpha(ts).tolist()
    lam, lm, std = ns.marginal_lambda(ts), ns.marginal_log_mean_coeff(ts), ns.marginal_std(ts)
    upd = []
    for step in range(1, S + 1):
        order = 1 if step == 1 else (min(2, S + 1 - step) if S < 15 else 2)
        s_, t_ = step - 1, step
        alpha_t = th.exp(lm[t_])
        c0 = std[t_] / std[s_]
        if order == 1:
            h = lam[t_] - lam[s_]
            upd.append((1, c0.item(), (alpha_t * th.expm1(-h)).item(), 0.0, 0.0))
        else:
            h_0 = lam[s_] - lam[s_ - 1]
            h = lam[t_] - lam[s_]
            r0 = h_0 / h
            c1 = alpha_t * (th.exp(-h) - 1.0)
            upd.append((2, c0.item(), c1.item(), (0.5 * c1).item(), (1.0 / r0).item()))
    return t_in, sig, alp, upd


class DPMCandidatePlan(CandidatePlan):
    """A DPM-Solver++(2M) candidate (S + 1 time points, S model evaluations) as one CUDA graph: the UNet runs on
    fractional timesteps; two data-prediction buffers alternate as (older, newer)."""

    def __init__(self, unet: UNetModel, alphas_cumprod: th.Tensor, ea_timesteps, batch: int, shape, scale: float, cfg: bool,
                 ctx_tokens: int = 77, use_graph: Optional[bool] = None):
        dev = unet._device()
        if dev.type != "cuda":
            raise RuntimeError("DPMCandidatePlan needs the model on a CUDA device (no CPU path)")
        C, H, W = shape
        self.method = "dpm_solver++"
        self.unet, self.B, self.cfg, self.scale = unet, batch, bool(cfg), float(scale)
        ns = DiscreteNoiseSchedule(alphas_cumprod)
        self.ts = dpm_time_steps(ea_timesteps, ns.total_N)
        self.steps = self.ts.tolist()
        self.t_model, self.sigmas, self.alphas_t, self.updates = dpm_schedule(ns, self.ts)
        n = batch * (2 if cfg else 1)
        sf = shared_forward(unet, n, (C, H, W), ctx_tokens, th.float32)  # fractional model timesteps
        self.x2, self.t_in, self.ctx, self.eps, self.plan_ctx, self.plan_fwd = sf.x2, sf.t_in, sf.ctx, sf.eps, sf.plan_ctx, sf.plan_fwd
        self.x = self.x2[:batch]
        self.m = [th.empty((batch, C, H, W), dtype=th.float32, device=dev) for _ in range(2)]
        S = len(self.updates)
        self.launches, per_fwd = sf.ctx_launches, sf.fwd_launches
        self.launches += S * (per_fwd + 2)
        self.launches_per_forward = per_fwd
        if use_graph is None:
            use_graph = os.environ.get("ADB_NO_GRAPH", "0") != "1"
        self.graph = None
        if use_graph:
            th.cuda.current_stream().synchronize()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                self._chain()
            self.graph = g

    def _model(self, k: int, dst: th.Tensor):
        self.t_in.fill_(self.t_model[k])
        self.plan_fwd.run()
        ops.dpm_x0(self.x, self.eps, self.sigmas[k], self.alphas_t[k], scale=self.scale, cfg=self.cfg, out=dst)

    def _chain(self):
        self.plan_ctx.run()
        S = len(self.updates)
        old, new = 0, 0
        self._model(0, self.m[0])
        for step in range(1, S + 1):
            order, c0, c1, c2, inv_r0 = self.updates[step - 1]
            ops.dpm_update(self.x, self.m[new], self.m[old] if order == 2 else None, order, c0, c1, c2, inv_r0, out=self.x)
            self._dup()
            if step < S:  # the final model value is never needed (dpm_solver.py:1119-1121)
                old, new = new, 1 - new
                self._model(step, self.m[new])


class DPMSolverSampler(object):
    """ldm/models/diffusion/dpm_solver/sampler.py:8-83: DPM-Solver++ (data prediction), multistep, order 2,
    lower_order_final, classifier-free guidance, `sampled_timestep` = S + 1 searched time points."""

    def __init__(self, model, **kwargs):
        self.model = model
        self.alphas_cumprod = model.alphas_cumprod.detach().float()
        self._plans: Dict[tuple, DPMCandidatePlan] = {}

    @th.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0, score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.0,
               unconditional_conditioning=None, sampled_timestep=None, **kwargs):
        if sampled_timestep is None:
            ns_T, t0 = 1.0, 1.0 / int(self.alphas_cumprod.shape[0])
            sampled_timestep = th.linspace(ns_T, t0, S + 1).tolist()  # get_time_steps('time_uniform'), dpm_solver.py:431-432
        assert len(sampled_timestep) - 1 == S  # dpm_solver.py:1097
        C, H, W = shape
        dev = self.model.device
        img = th.randn((batch_size, C, H, W), device=dev) if x_T is None else x_T
        cfg = not (unconditional_guidance_scale == 1.0 or unconditional_conditioning is None)  # dpm_solver.py:337
        unet = getattr(self.model, "unet", None)
        if isinstance(unet, UNetModel) and not isinstance(conditioning, dict):
            key = (tuple(float(v) for v in sampled_timestep), batch_size, C, H, W, float(unconditional_guidance_scale), cfg,
                   conditioning.shape[1])
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 4:
                    self._plans.pop(next(iter(self._plans)))
                plan = DPMCandidatePlan(unet, self.alphas_cumprod, sampled_timestep, batch_size, (C, H, W),
                                        unconditional_guidance_scale, cfg, ctx_tokens=conditioning.shape[1])
                self._plans[key] = plan
            return plan.run(img, conditioning, unconditional_conditioning).clone(), None
        # generic apply_model: dpm_solver.py:1099-1121 with the fused kernels
        ns = DiscreteNoiseSchedule(self.alphas_cumprod)
        ts = dpm_time_steps(sampled_timestep, ns.total_N)
        t_model, sig, alp, upd = dpm_schedule(ns, ts)

        def model(x, k):
            t = th.full((batch_size,), t_model[k], device=dev, dtype=th.float32)
            if cfg:
                eps = self.model.apply_model(th.cat([x] * 2), th.cat([t] * 2), th.cat([unconditional_conditioning, conditioning]))
            else:
                eps = self.model.apply_model(x, t, conditioning)
            return ops.dpm_x0(x, eps.float().contiguous(), sig[k], alp[k], scale=unconditional_guidance_scale, cfg=cfg)

        x = img.contiguous()
        m_old = m_new = model(x, 0)
        for step in range(1, len(upd) + 1):
            order, c0, c1, c2, inv_r0 = upd[step - 1]
            x = ops.dpm_update(x, m_new, m_old if order == 2 else None, order, c0, c1, c2, inv_r0)
            if step < len(upd):
                m_old, m_new = m_new, model(x, step)
        return x, None
