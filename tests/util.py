"""Shared helpers for the tests: fixtures, oracle weights and builders."""
import contextlib
import os

import numpy as np
import torch

from oracle import unet_ref, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ADM_FLAGS = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=192, num_head_channels=64, num_res_blocks=3,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=True, use_scale_shift_norm=True,
                 use_dynamic_unet=True)
SMALL_FLAGS = dict(ADM_FLAGS, num_channels=64, num_res_blocks=1)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def cfg_of(flags):
    return unet_ref.UNetConfig(model_channels=flags["num_channels"], num_res_blocks=flags["num_res_blocks"])


def oracle_weights(flags, seed=0):
    cfg = cfg_of(flags)
    return cfg, weights.make_state_dict(unet_ref.param_shapes(cfg), seed=seed)


def build_ours(flags, sd, device="cuda"):
    from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults

    d = model_and_diffusion_defaults()
    d.update(flags)
    model, diffusion = create_model_and_diffusion(**d)
    model.load_state_dict(sd)
    model.to(device)
    if flags.get("use_fp16"):
        model.convert_to_fp16()
    model.eval()
    return model, diffusion


def parse_skip_list(arr):
    return [[int(v) for v in s.split(",") if v != ""] for s in arr.tolist()]


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


@contextlib.contextmanager
def no_fast_path():
    """Force `ddim_sample_loop` onto the generic per-step loop (fastpath.py honours ADB_NO_FAST_PATH at every call)."""
    old = os.environ.get("ADB_NO_FAST_PATH")
    os.environ["ADB_NO_FAST_PATH"] = "1"
    try:
        yield
    finally:
        if old is None:
            os.environ.pop("ADB_NO_FAST_PATH", None)
        else:
            os.environ["ADB_NO_FAST_PATH"] = old
