"""Timestep respacing: the searched subsequence -> a K'-step diffusion process.

Mirrors guided_diffusion/respace.py (space_timesteps :7-60, SpacedDiffusion :63-113,
_WrappedModel :116-127) and adds `reset_diffusion`, the in-place rebuild the search drivers
perform per candidate (…progressive.py:219-274; scripts/classifier_sample_prunedUNET.py:28-83).
Integer work (set dedup, ascending timestep_map) is bit-exact by construction; float64
tables follow the reference's betas -> cumprod route.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """respace.py:7-60."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        if section_count <= 1:
            frac_stride = 1
        else:
            frac_stride = (size - 1) / (section_count - 1)
        cur_idx = 0.0
        taken_steps = []
        for _ in range(section_count):
            taken_steps.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken_steps
        start_idx += size
    return set(all_steps)


def respaced_betas(base_alphas_cumprod, use_timesteps):
    """respace.py:76-84: new_beta_i = 1 - acp[t_i] / acp[t_{i-1}] over the ascending kept steps."""
    use = set(use_timesteps)
    last_alpha_cumprod = 1.0
    new_betas = []
    timestep_map = []
    for i, alpha_cumprod in enumerate(base_alphas_cumprod):
        if i in use:
            new_betas.append(1 - alpha_cumprod / last_alpha_cumprod)
            last_alpha_cumprod = alpha_cumprod
            timestep_map.append(i)
    return timestep_map, np.array(new_betas, dtype=np.float64)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:63-113."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base_diffusion = GaussianDiffusion(**kwargs)
        self.timestep_map, new_betas = respaced_betas(base_diffusion.alphas_cumprod, self.use_timesteps)
        kwargs["betas"] = new_betas
        super().__init__(**kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t  # scaling is done by the wrapped model (respace.py:110-112)


class _WrappedModel:
    """respace.py:116-127. The step index -> original timestep gather stays an exact integer op."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps

    def __call__(self, x, ts, **kwargs):
        map_tensor = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = map_tensor[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)


def reset_diffusion(use_timesteps, active_diffusion, base_diffusion):
    """Rebuild `active_diffusion` in place for a candidate's timesteps.

    Same effect as EvolutionSearcher.reset_diffusion (…progressive.py:219-274) and the module-level
    copy in scripts/classifier_sample_prunedUNET.py:28-83, which callers may keep using unchanged
    on our objects (they only touch numpy attributes).
    """
    tmap, new_betas = respaced_betas(base_diffusion.alphas_cumprod, use_timesteps)
    active_diffusion.use_timesteps = set(use_timesteps)
    active_diffusion.timestep_map = tmap
    GaussianDiffusion.__init__(
        active_diffusion,
        betas=new_betas,
        model_mean_type=active_diffusion.model_mean_type,
        model_var_type=active_diffusion.model_var_type,
        loss_type=active_diffusion.loss_type,
        rescale_timesteps=active_diffusion.rescale_timesteps,
    )
    return active_diffusion
