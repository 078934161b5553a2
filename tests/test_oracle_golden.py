"""CPU: the oracle restatement reproduces the reference's own outputs (committed fixtures made by
tests/golden/make_golden.py from the imported, unmodified reference) — bit-exact for the integer
maps and float64 tables, bit-exact for the fp32 UNet / sampler on the same torch build, and within
a few ulp otherwise (stated per test)."""
import numpy as np
import pytest
import torch

from oracle import diffusion_ref, fid_ref, unet_ref
from tests.util import SMALL_FLAGS, golden, oracle_weights, parse_skip_list

CANDS = {
    "cand10": [744, 137, 647, 856, 305, 441, 676, 572, 971, 85],
    "cand4": [153, 424, 926, 690],
    "cand6": [94, 834, 217, 944, 574, 354],
    "dedup": [5, 5, 900],
    "single": [500],
}


@pytest.mark.parametrize("name", list(CANDS))
def test_tables_bit_exact(name):
    g = golden("tables.npz")
    base = diffusion_ref.base_tables("cosine", 1000)
    assert np.array_equal(base["betas"], g["base/betas"])
    assert np.array_equal(base["alphas_cumprod"], g["base/alphas_cumprod"])
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], CANDS[name])
    assert tmap == g[f"{name}/timestep_map"].tolist()  # integer map: exact
    tb = diffusion_ref.diffusion_tables(nb)
    for k, v in tb.items():
        assert np.array_equal(v, g[f"{name}/{k}"]), k  # float64 tables: bit-exact


def test_dedup_and_sorted_map():
    g = golden("tables.npz")
    assert g["dedup/timestep_map"].tolist() == [5, 900]
    assert g["cand10/timestep_map"].tolist() == sorted(CANDS["cand10"])


def test_unet_small_matches_reference():
    g = golden("unet_small.npz")
    cfg, sd = oracle_weights(SMALL_FLAGS)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    assert int(g["layer_num"]) == unet_ref.build_arch(cfg).layer_num
    for i in range(3):
        t = torch.full((x.shape[0],), int(g[f"t{i}"]), dtype=torch.long)
        with torch.no_grad():
            out = unet_ref.unet_forward(sd, cfg, x, t, y, g[f"skip{i}"].tolist())
        # same torch build -> 0.0; allow fp32 reassociation noise of a different BLAS
        assert (out - torch.from_numpy(g[f"out{i}"])).abs().max().item() <= 2e-5


def test_ddim_small_matches_reference():
    g = golden("ddim_small.npz")
    cfg, sd = oracle_weights(SMALL_FLAGS)
    ccfg = unet_ref.classifier64_config(depth=1, width=64)
    from oracle import weights

    csd = weights.make_state_dict(unet_ref.param_shapes(ccfg, encoder_only=True), seed=1)
    base = diffusion_ref.base_tables("cosine", 1000)
    noise, y = torch.from_numpy(g["noise"]), torch.from_numpy(g["y"])
    for name in ["guided", "dedup", "noguide"]:
        ts = g[f"{name}/timesteps"].tolist()
        skips = parse_skip_list(g[f"{name}/skip_layers"])
        scale = float(g[f"{name}/scale"])
        tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], ts)
        tb = diffusion_ref.diffusion_tables(nb)
        seen = []

        def unet(x, t, yy, skip):
            seen.append((int(t[0]), list(skip)))
            return unet_ref.unet_forward(sd, cfg, x, t, yy, skip)

        outs = diffusion_ref.ddim_sample_loop(
            diffusion_ref.make_model_fn(unet, tmap), noise.shape, tb, tmap, noise, True,
            cond_fn=unet_ref.classifier_cond_fn(csd, ccfg, scale) if scale >= 0 else None,
            model_kwargs={"y": y, "skip_layers": skips}, return_all=True)
        # which (timestep, skip list) pairs the model saw: integer decisions, exact
        assert [s[0] for s in seen] == g[f"{name}/seen_t"].tolist()
        assert [s[1] for s in seen] == parse_skip_list(g[f"{name}/seen_skip"])
        assert (outs[1] - torch.from_numpy(g[f"{name}/step1"])).abs().max().item() <= 1e-4
        assert (outs[-1] - torch.from_numpy(g[f"{name}/final"])).abs().max().item() <= 1e-3
        u8 = diffusion_ref.pack_uint8(torch.from_numpy(g[f"{name}/final"]))
        assert np.array_equal(u8.numpy(), g[f"{name}/uint8"])


@pytest.mark.parametrize("name", ["full", "singular"])
def test_fid_matches_reference(name):
    g = golden("fid.npz")
    fid = fid_ref.frechet_distance(*fid_ref.compute_statistics(g[f"{name}/f1"]), *fid_ref.compute_statistics(g[f"{name}/f2"]))
    assert abs(fid - float(g[f"{name}/fid"])) <= 1e-6 * max(1.0, abs(float(g[f"{name}/fid"])))
