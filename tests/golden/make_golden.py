"""Generate the golden fixtures by running the UNMODIFIED reference (imported read-only from
/root/reference) on weights from oracle.weights, and check the oracle restatement against it.

Run in the authoring container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Outputs (committed): tests/golden/*.npz. Every array in them was produced by reference code;
`tests/test_oracle_golden.py` replays the oracle against them anywhere.
"""
import copy
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/examples/guided_diffusion"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.modules.setdefault("blobfile", types.ModuleType("blobfile"))  # dist_util imports it; unused here

from guided_diffusion import gaussian_diffusion as rgd  # noqa: E402
from guided_diffusion.respace import SpacedDiffusion  # noqa: E402
from guided_diffusion.script_util import (  # noqa: E402
    classifier_defaults, create_classifier, create_model_and_diffusion, model_and_diffusion_defaults)
from guided_diffusion.unet import EncoderUNetModel  # noqa: E402

from oracle import diffusion_ref, fid_ref, unet_ref, weights  # noqa: E402

torch.set_grad_enabled(True)
ADM_FLAGS = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=192, num_head_channels=64, num_res_blocks=3,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=False, use_scale_shift_norm=True,
                 use_dynamic_unet=True)
SMALL_FLAGS = dict(ADM_FLAGS, num_channels=64, num_res_blocks=1)

CANDIDATES = {
    # GD/sample_imagenet64_classifier_guidance_dynamic_subnet.sh:13-14
    "cand10": dict(timesteps=[744, 137, 647, 856, 305, 441, 676, 572, 971, 85],
                   skip_layers=[[], [], [], [], [], [], [30, 10, 39, 4, 15, 46, 49, 54, 8], [], [], []]),
    # GD/scripts/classifier_sample_generate_image.py:159-168
    "cand4": dict(timesteps=[153, 424, 926, 690], skip_layers=[[], [], [], []]),
    # GD/sample_imagenet64_classifier_guidance_subnet.sh:11
    "cand6": dict(timesteps=[94, 834, 217, 944, 574, 354], skip_layers=[[]] * 6),
    # duplicates collapse (set()), K' < K; single step exercises the K=1 special case
    "dedup": dict(timesteps=[5, 5, 900], skip_layers=[[1], [2], [3]]),
    "single": dict(timesteps=[500], skip_layers=[[]]),
}


def load_reset_diffusion():
    spec = importlib.util.spec_from_file_location("ref_sampler", os.path.join(REF, "scripts", "classifier_sample_prunedUNET.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.reset_diffusion


def build(flags):
    d = model_and_diffusion_defaults()
    d.update(flags)
    model, diffusion = create_model_and_diffusion(**d)
    return model.eval(), diffusion


def cfg_of(flags):
    return unet_ref.UNetConfig(model_channels=flags["num_channels"], num_res_blocks=flags["num_res_blocks"])


TABLE_KEYS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
              "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"]


def gen_tables():
    reset_diffusion = load_reset_diffusion()
    _, diffusion = build(SMALL_FLAGS)
    base = copy.deepcopy(diffusion)
    out = {}
    for name, cand in CANDIDATES.items():
        active = copy.deepcopy(base)
        reset_diffusion(cand["timesteps"], active, base)  # reference code, in place
        out[f"{name}/timestep_map"] = np.array(active.timestep_map, dtype=np.int64)
        for k in TABLE_KEYS:
            out[f"{name}/{k}"] = np.asarray(getattr(active, k), dtype=np.float64)
        if len(set(cand["timesteps"])) > 1:  # SpacedDiffusion.__init__ cannot build K'=1 (indexes [1])
            sd = SpacedDiffusion(use_timesteps=cand["timesteps"], betas=base.betas, model_mean_type=rgd.ModelMeanType.EPSILON,
                                 model_var_type=rgd.ModelVarType.LEARNED_RANGE, loss_type=rgd.LossType.MSE)
            assert sd.timestep_map == active.timestep_map
            for k in TABLE_KEYS:
                assert np.array_equal(getattr(sd, k), getattr(active, k)), k
        # oracle check
        tmap, nb = diffusion_ref.respace(base.alphas_cumprod, cand["timesteps"])
        tb = diffusion_ref.diffusion_tables(nb)
        assert tmap == active.timestep_map
        for k in TABLE_KEYS:
            assert np.array_equal(tb[k], out[f"{name}/{k}"]), (name, k)
    out["base/betas"] = base.betas
    out["base/alphas_cumprod"] = base.alphas_cumprod
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **out)
    print("tables.npz ok")


def gen_unet(tag, flags, batch, skips):
    model, _ = build(flags)
    cfg = cfg_of(flags)
    shapes = unet_ref.param_shapes(cfg)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == shapes
    assert list(model.state_dict().keys()) == list(shapes.keys())
    sd = weights.make_state_dict(shapes, seed=0)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, 3, 64, 64, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=torch.Generator().manual_seed(3))
    out = {"x": x.numpy(), "y": y.numpy(), "layer_num": np.int64(model.layer_num)}
    for i, (t, skip) in enumerate(skips):
        tt = torch.full((batch,), t, dtype=torch.long)
        with torch.no_grad():
            ref = model(x, tt, y, skip_layer=skip)
            mine = unet_ref.unet_forward(sd, cfg, x, tt, y, skip)
        d = (ref - mine).abs().max().item()
        print(f"{tag} t={t} skip={skip}: ref std {ref.std():.4f} oracle max-abs diff {d}")
        assert d == 0.0
        out[f"t{i}"] = np.int64(t)
        out[f"skip{i}"] = np.array(skip, dtype=np.int64)
        out[f"out{i}"] = ref.numpy()
    np.savez_compressed(os.path.join(HERE, f"unet_{tag}.npz"), **out)


def gen_ddim_small():
    """Full classifier-guided searched-DDIM loop of the reference on the small UNet, B=2."""
    model, diffusion = build(SMALL_FLAGS)
    cfg = cfg_of(SMALL_FLAGS)
    sd = weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0)
    model.load_state_dict(sd)
    cd = classifier_defaults()
    cd.update(classifier_depth=1, classifier_width=64)
    clf = create_classifier(**cd).eval()
    ccfg = unet_ref.classifier64_config(depth=1, width=64)
    cshapes = unet_ref.param_shapes(ccfg, encoder_only=True)
    assert {k: tuple(v.shape) for k, v in clf.state_dict().items()} == cshapes
    csd = weights.make_state_dict(cshapes, seed=1)
    clf.load_state_dict(csd)
    reset_diffusion = load_reset_diffusion()
    base = copy.deepcopy(diffusion)
    B = 2
    noise = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3))
    out = {"noise": noise.numpy(), "y": y.numpy()}
    import torch.nn.functional as F
    for name, cand, scale in [("guided", dict(timesteps=[690, 153, 926, 424], skip_layers=[[], [2, 9], [], [5, 12, 17]]), 1.0),
                              ("dedup", CANDIDATES["dedup"], 2.5),
                              ("noguide", dict(timesteps=[85, 971], skip_layers=[[0], []]), None)]:
        active = copy.deepcopy(base)
        reset_diffusion(cand["timesteps"], active, base)
        skip_layers = cand["skip_layers"]
        seen = []

        def cond_fn(x, t, y=None, skip_layers=None, timesteps=None):  # …progressive.py:383-390
            with torch.enable_grad():
                x_in = x.detach().requires_grad_(True)
                logits = clf(x_in, t)
                log_probs = F.log_softmax(logits, dim=-1)
                selected = log_probs[range(len(logits)), y.view(-1)]
                return torch.autograd.grad(selected.sum(), x_in)[0] * scale

        def model_fn(x, t, y=None, skip_layers=None, timesteps=None):  # …progressive.py:392-397
            t_index = active.timestep_map.index(t[0])
            seen.append((int(t[0]), list(skip_layers[t_index])))
            return model(x, t, y, skip_layer=skip_layers[t_index])

        imgs = active.ddim_sample_loop(model_fn, (B, 3, 64, 64), noise=noise, clip_denoised=True,
                                       model_kwargs={"y": y, "skip_layers": skip_layers},
                                       cond_fn=cond_fn if scale is not None else None, device="cpu", return_all_images=True)
        # oracle
        tmap, nb = diffusion_ref.respace(base.alphas_cumprod, cand["timesteps"])
        tb = diffusion_ref.diffusion_tables(nb)
        unet = lambda x, t, yy, skip: unet_ref.unet_forward(sd, cfg, x, t, yy, skip)
        o = diffusion_ref.ddim_sample_loop(
            diffusion_ref.make_model_fn(unet, tmap), (B, 3, 64, 64), tb, tmap, noise, True,
            cond_fn=unet_ref.classifier_cond_fn(csd, ccfg, scale) if scale is not None else None,
            model_kwargs={"y": y, "skip_layers": skip_layers}, return_all=True)
        assert len(o) == len(imgs)
        dmax = max((a - b).abs().max().item() for a, b in zip(o, imgs))
        print(f"ddim {name}: steps {len(imgs) - 1} final std {imgs[-1].std():.4f} oracle-vs-reference max diff {dmax}; (t, skip) seen {seen}")
        assert dmax == 0.0
        out[f"{name}/timesteps"] = np.array(cand["timesteps"], dtype=np.int64)
        out[f"{name}/skip_layers"] = np.array([",".join(map(str, s)) for s in skip_layers])
        out[f"{name}/scale"] = np.float64(-1.0 if scale is None else scale)
        out[f"{name}/final"] = imgs[-1].numpy()
        out[f"{name}/step1"] = imgs[1].numpy()
        out[f"{name}/seen_t"] = np.array([s[0] for s in seen], dtype=np.int64)
        out[f"{name}/seen_skip"] = np.array([",".join(map(str, s[1])) for s in seen])
        out[f"{name}/uint8"] = ((imgs[-1] + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "ddim_small.npz"), **out)


def gen_fid():
    """FIDStatistics.frechet_distance of the search script (…progressive.py:104-153). scipy>=1.16 dropped
    sqrtm(disp=...): the reference line is run with a shim that restores the old (value, errest) return."""
    from scipy import linalg
    real = linalg.sqrtm
    linalg.sqrtm = lambda a, disp=True: (real(a), 0.0) if disp is False else real(a)
    try:
        src = open(os.path.join(REF, "search_dynamic_unet_imagenet64_classifier_guidance_progressive.py")).read()
        start = src.index("class FIDStatistics:")
        end = src.index("class EvolutionSearcher")
        ns = {"np": np}
        exec(compile(src[start:end], "ref_fid", "exec"), ns)
        FIDStatistics = ns["FIDStatistics"]
        rng = np.random.RandomState(0)
        d = 64
        out = {}
        for name, n1, n2 in [("full", 500, 400), ("singular", 40, 50)]:  # N < d: singular covariance
            f1 = rng.randn(n1, d).astype(np.float32) * (1 + rng.rand(d)).astype(np.float32) + 0.3
            f2 = rng.randn(n2, d).astype(np.float32) @ (np.eye(d) + 0.1 * rng.randn(d, d)).astype(np.float32)
            m1, s1 = np.mean(f1, axis=0), np.cov(f1, rowvar=False)  # evaluator_v1.py:218-221
            m2, s2 = np.mean(f2, axis=0), np.cov(f2, rowvar=False)
            fid = FIDStatistics(m1, s1).frechet_distance(FIDStatistics(m2, s2))
            mine = fid_ref.frechet_distance(*fid_ref.compute_statistics(f1), *fid_ref.compute_statistics(f2))
            print(f"fid {name}: reference {fid} oracle {mine}")
            assert abs(fid - mine) <= 1e-9 * max(1.0, abs(fid))
            out[f"{name}/f1"], out[f"{name}/f2"], out[f"{name}/fid"] = f1, f2, np.float64(fid)
        np.savez_compressed(os.path.join(HERE, "fid.npz"), **out)
    finally:
        linalg.sqrtm = real


def gen_config1():
    """BASELINE config 1: full ADM-G 64 + depth-4 classifier, 4-step searched schedule [153,424,926,690],
    full architecture, classifier_scale 1.0, batch 8 - the reference's own ddim_sample_loop on CPU."""
    import torch.nn.functional as F

    model, diffusion = build(ADM_FLAGS)
    cfg = cfg_of(ADM_FLAGS)
    sd = weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0)
    model.load_state_dict(sd)
    cd = classifier_defaults()
    cd.update(classifier_depth=4)
    clf = create_classifier(**cd).eval()
    ccfg = unet_ref.classifier64_config(depth=4, width=128)
    csd = weights.make_state_dict(unet_ref.param_shapes(ccfg, encoder_only=True), seed=1)
    clf.load_state_dict(csd)
    reset_diffusion = load_reset_diffusion()
    base = copy.deepcopy(diffusion)
    B = 8
    noise = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3))
    cand = CANDIDATES["cand4"]
    active = copy.deepcopy(base)
    reset_diffusion(cand["timesteps"], active, base)

    def cond_fn(x, t, y=None, skip_layers=None, timesteps=None):
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = clf(x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return torch.autograd.grad(selected.sum(), x_in)[0] * 1.0

    def model_fn(x, t, y=None, skip_layers=None, timesteps=None):
        t_index = active.timestep_map.index(t[0])
        return model(x, t, y, skip_layer=skip_layers[t_index])

    import time
    t0 = time.time()
    imgs = active.ddim_sample_loop(model_fn, (B, 3, 64, 64), noise=noise, clip_denoised=True,
                                   model_kwargs={"y": y, "skip_layers": cand["skip_layers"]}, cond_fn=cond_fn,
                                   device="cpu", return_all_images=True)
    print(f"config1 reference: {time.time() - t0:.1f} s for {B} images; final std {imgs[-1].std():.4f}")
    np.savez_compressed(os.path.join(HERE, "config1_admg64_guided.npz"), noise=noise.numpy(), y=y.numpy(),
                        timesteps=np.array(cand["timesteps"], dtype=np.int64), final=imgs[-1].numpy(),
                        step1=imgs[1].numpy(),
                        uint8=((imgs[-1] + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy())


def gen_config2():
    """BASELINE config 2 (the benchmarked candidate) at batch 8: full ADM-G 64 + depth-4 classifier, the published
    10-step schedule with its block-skip mask (GD/sample_imagenet64_classifier_guidance_dynamic_subnet.sh:13-14),
    classifier_scale 1.0 - the reference's own ddim_sample_loop with the search script's closures
    (...progressive.py:383-397) on CPU. Also records which (timestep, skip list) pairs the UNet saw."""
    import torch.nn.functional as F

    model, diffusion = build(ADM_FLAGS)
    cfg = cfg_of(ADM_FLAGS)
    sd = weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0)
    model.load_state_dict(sd)
    cd = classifier_defaults()
    cd.update(classifier_depth=4)
    clf = create_classifier(**cd).eval()
    ccfg = unet_ref.classifier64_config(depth=4, width=128)
    csd = weights.make_state_dict(unet_ref.param_shapes(ccfg, encoder_only=True), seed=1)
    clf.load_state_dict(csd)
    reset_diffusion = load_reset_diffusion()
    base = copy.deepcopy(diffusion)
    B = 8
    noise = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(12))
    y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(13))
    cand = CANDIDATES["cand10"]
    active = copy.deepcopy(base)
    reset_diffusion(cand["timesteps"], active, base)
    seen = []

    def cond_fn(x, t, y=None, skip_layers=None, timesteps=None):
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = clf(x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return torch.autograd.grad(selected.sum(), x_in)[0] * 1.0

    def model_fn(x, t, y=None, skip_layers=None, timesteps=None):
        t_index = active.timestep_map.index(t[0])
        seen.append((int(t[0]), list(skip_layers[t_index])))
        return model(x, t, y, skip_layer=skip_layers[t_index])

    import time
    t0 = time.time()
    imgs = active.ddim_sample_loop(model_fn, (B, 3, 64, 64), noise=noise, clip_denoised=True,
                                   model_kwargs={"y": y, "skip_layers": cand["skip_layers"]}, cond_fn=cond_fn,
                                   device="cpu", return_all_images=True)
    print(f"config2 reference: {time.time() - t0:.1f} s for {B} images; final std {imgs[-1].std():.4f}; seen {seen}")
    np.savez_compressed(os.path.join(HERE, "config2_admg64_cand10_guided.npz"), noise=noise.numpy(), y=y.numpy(),
                        timesteps=np.array(cand["timesteps"], dtype=np.int64),
                        skip_layers=np.array([",".join(map(str, s)) for s in cand["skip_layers"]]),
                        seen_t=np.array([s[0] for s in seen], dtype=np.int64),
                        seen_skip=np.array([",".join(map(str, s[1])) for s in seen]),
                        final=imgs[-1].numpy(), step1=imgs[1].numpy(), step5=imgs[5].numpy(),
                        uint8=((imgs[-1] + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config1":
        gen_config1()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "config2":
        gen_config2()
        sys.exit(0)
    torch.manual_seed(0)
    gen_tables()
    gen_fid()
    gen_unet("small", SMALL_FLAGS, 2, [(676, []), (85, [0, 1, 3, 4, 5, 9, 11, 12, 13, 15]), (971, [2, 6, 7, 8, 10, 14, 16, 17])])
    gen_ddim_small()
    gen_unet("admg64", ADM_FLAGS, 1, [(153, []), (676, [30, 10, 39, 4, 15, 46, 49, 54, 8])])
    gen_config1()
    gen_config2()
    print("done")
