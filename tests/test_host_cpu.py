"""CPU: host-side logic of the product (no compute calls): diffusion tables, respacing, coefficient
rounding, state_dict compatibility, flag system, and that the C-ABI library loads and exports every
symbol include/adb200.h declares."""
import copy
import os
import re

import numpy as np
import pytest
import torch

from oracle import unet_ref
from tests.util import ADM_FLAGS, SMALL_FLAGS, golden

CANDS = {
    "cand10": [744, 137, 647, 856, 305, 441, 676, 572, 971, 85],
    "cand4": [153, 424, 926, 690],
    "dedup": [5, 5, 900],
    "single": [500],
}
TABLE_KEYS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]


def _diffusion():
    from autodiffusion_b200 import create_gaussian_diffusion

    return create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="cosine")


@pytest.mark.parametrize("name", list(CANDS))
def test_reset_diffusion_tables_bit_exact(name):
    from autodiffusion_b200.respace import SpacedDiffusion, reset_diffusion
    from autodiffusion_b200 import gaussian_diffusion as gd

    g = golden("tables.npz")
    base = _diffusion()
    assert np.array_equal(base.betas, g["base/betas"])
    assert np.array_equal(base.alphas_cumprod, g["base/alphas_cumprod"])
    active = copy.deepcopy(base)  # the search deep-copies the diffusion (…progressive.py:163)
    reset_diffusion(CANDS[name], active, base)
    assert active.timestep_map == g[f"{name}/timestep_map"].tolist()
    assert active.num_timesteps == len(set(CANDS[name]))
    for k in TABLE_KEYS:
        assert np.array_equal(getattr(active, k), g[f"{name}/{k}"]), k
    sd = SpacedDiffusion(use_timesteps=CANDS[name], betas=base.betas, model_mean_type=gd.ModelMeanType.EPSILON,
                         model_var_type=gd.ModelVarType.LEARNED_RANGE, loss_type=gd.LossType.MSE)
    assert sd.timestep_map == active.timestep_map
    for k in TABLE_KEYS:
        assert np.array_equal(getattr(sd, k), g[f"{name}/{k}"]), k


def test_reference_reset_diffusion_code_runs_on_our_object():
    """Callers keep their own reset_diffusion (attribute mutation on numpy arrays): restated inline
    from …progressive.py:219-231 to show our object accepts it."""
    base = _diffusion()
    active = copy.deepcopy(base)
    use = set([153, 424, 926, 690])
    active.timestep_map = []
    last, nb = 1.0, []
    for i, acp in enumerate(base.alphas_cumprod):
        if i in use:
            nb.append(1 - acp / last)
            last = acp
            active.timestep_map.append(i)
    active.betas = np.array(nb)
    active.num_timesteps = 4
    assert active.timestep_map == [153, 424, 690, 926]
    assert base.original_num_steps == 1000 and base.rescale_timesteps is False


def test_space_timesteps():
    from autodiffusion_b200.respace import space_timesteps

    assert space_timesteps(1000, "ddim10") == set(range(0, 1000, 100))
    assert space_timesteps(300, [10, 15, 20]) == space_timesteps(300, "10,15,20")
    assert len(space_timesteps(1000, [1000])) == 1000
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError):
        space_timesteps(10, [20])


def test_ddim_coefficients_are_fp32_roundings():
    from autodiffusion_b200.gaussian_diffusion import ddim_coefficients
    from autodiffusion_b200.respace import reset_diffusion

    base = _diffusion()
    active = reset_diffusion(CANDS["cand10"], copy.deepcopy(base), base)
    for i in range(active.num_timesteps):
        c = ddim_coefficients(active, i)
        t = torch.tensor([i])
        ex = lambda arr: torch.from_numpy(arr)[t].float()  # _extract_into_tensor, gaussian_diffusion.py:920
        ab, abp = ex(active.alphas_cumprod), ex(active.alphas_cumprod_prev)
        ref = [ex(active.sqrt_recip_alphas_cumprod), ex(active.sqrt_recipm1_alphas_cumprod), (1 - ab).sqrt(),
               torch.sqrt(abp), torch.sqrt(1 - abp - torch.zeros(1) ** 2)]
        assert [np.float32(v) for v in c] == [np.float32(r.item()) for r in ref]
    assert ddim_coefficients(active, 0)[3] == 1.0 and ddim_coefficients(active, 0)[4] == 0.0
    with pytest.raises(NotImplementedError):
        ddim_coefficients(active, 0, eta=0.5)


@pytest.mark.parametrize("flags", [SMALL_FLAGS, ADM_FLAGS])
def test_state_dict_keys_match_reference(flags):
    from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults

    d = model_and_diffusion_defaults()
    d.update(flags)
    with torch.device("meta"):
        model, diffusion = create_model_and_diffusion(**d)
    cfg = unet_ref.UNetConfig(model_channels=flags["num_channels"], num_res_blocks=flags["num_res_blocks"])
    want = unet_ref.param_shapes(cfg)  # verified == reference state_dict in tests/golden/make_golden.py
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == want
    assert model.layer_num == unet_ref.build_arch(cfg).layer_num
    if flags is ADM_FLAGS:
        assert model.layer_num == 58 and sum(v.numel() for v in model.state_dict().values()) == 295904454


def test_defaults_and_argparser():
    import argparse

    from autodiffusion_b200 import add_dict_to_argparser, args_to_dict, model_and_diffusion_defaults

    d = model_and_diffusion_defaults()
    assert d["use_dynamic_unet"] is False and d["rescale_timesteps"] is False and d["num_channels"] == 128
    p = argparse.ArgumentParser()
    add_dict_to_argparser(p, d)
    a = p.parse_args(["--class_cond", "True", "--num_channels", "192", "--use_fp16", "no"])
    dd = args_to_dict(a, d.keys())
    assert dd["class_cond"] is True and dd["num_channels"] == 192 and dd["use_fp16"] is False


def test_cpu_tensors_fail_loudly():
    from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults

    d = model_and_diffusion_defaults()
    d.update(SMALL_FLAGS)
    model, diffusion = create_model_and_diffusion(**d)
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x, torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
    with pytest.raises(AssertionError):
        model.to("meta")(x, torch.zeros(1, dtype=torch.long))  # y required iff class-conditional
    with pytest.raises(RuntimeError, match="CUDA"):
        diffusion.ddim_sample_loop(lambda *a, **k: None, (1, 3, 64, 64), noise=x, device="cpu")


def test_library_loads_and_exports_header_symbols():
    from autodiffusion_b200 import _lib

    _lib.build_library()
    handle = _lib.lib()
    header = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(adb_[a-z0-9_]+)\s*\(", header))
    declared -= {"adb_plan"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.adb_version() == 100
    assert handle.adb_conv_block_n(192) == 192 and handle.adb_conv_block_n(6) == 16
    assert handle.adb_plan_num_ops(None) == 0
    # host-only predicate of the conv epilogue's GroupNorm-backward sums (full tiles on the all-TMA path, include/adb200.h)
    assert handle.adb_conv_gnb_supported(256, 64, 64, 128) == 1 and handle.adb_conv_gnb_supported(256, 8, 8, 512) == 1
    assert handle.adb_conv_gnb_supported(1, 8, 8, 512) == 0    # 64 rows: not a full 128-row tile
    assert handle.adb_conv_gnb_supported(2, 4, 4, 512) == 0    # 16 pixels per image: a 32-row slab would span images
    assert handle.adb_conv_gnb_supported(4, 16, 16, 48) == 0   # channels must divide into 32 groups


def test_fast_path_locates_the_modules_behind_the_reference_closures():
    """fastpath.py (host side): the UNet / classifier behind the search script's closures are found through the closure cells
    (`self.model`, `self.classifier` of the captured searcher, or the module captured directly); a CPU call is never fused."""
    import types

    import torch

    from autodiffusion_b200 import classifier_defaults, create_classifier, create_model_and_diffusion, model_and_diffusion_defaults
    from autodiffusion_b200.classifier import ClassifierGuidance
    from autodiffusion_b200.fastpath import _find_modules, try_fast_path

    d = model_and_diffusion_defaults()
    d.update(attention_resolutions="32,16,8", class_cond=True, image_size=64, learn_sigma=True, num_channels=64, num_head_channels=64,
             num_res_blocks=1, resblock_updown=True, use_new_attention_order=True, use_scale_shift_norm=True, use_dynamic_unet=True)
    model, diffusion = create_model_and_diffusion(**d)
    cd = classifier_defaults()
    cd.update(classifier_depth=1, classifier_width=64)
    clf = create_classifier(**cd)
    self = types.SimpleNamespace(model=model, classifier=clf, active_diffusion=diffusion)
    args = types.SimpleNamespace(classifier_scale=1.0, class_cond=True)

    def cond_fn(x, t, y=None, skip_layers=None, timesteps=None):
        return self.classifier(x, t) * args.classifier_scale

    def model_fn(x, t, y=None, skip_layers=None, timesteps=None):
        return self.model(x, t, y if args.class_cond else None, skip_layer=skip_layers[self.active_diffusion.timestep_map.index(t[0])])

    direct = lambda x, t, y=None: model(x, t, y)
    assert _find_modules(model_fn)[0] == [model] and _find_modules(cond_fn)[1] == [clf]
    assert _find_modules(direct)[0] == [model] and _find_modules(model)[0] == [model]
    assert _find_modules(ClassifierGuidance(clf, 2.0))[1] == [clf]
    assert _find_modules(lambda x, t: x) == ([], [])
    noise = torch.zeros(2, 3, 64, 64)
    assert try_fast_path(diffusion, model_fn, (2, 3, 64, 64), noise, True, cond_fn, {"y": torch.zeros(2, dtype=torch.long)}, "cpu") is None
