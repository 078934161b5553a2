"""Where does the largest sample error of a short searched-DDIM run come from?

`__graft_entry__.smoke()` reported max_abs = 0.461 at 38.3 dB for its 2-step candidate: an outlier ~19x
the RMS error. This script replays that run (and optionally others) step by step against the CPU oracle and
prints, per sampled step: the per-step coefficients A = sqrt(1/abar), Bm = sqrt(1/abar - 1), the error of the
UNet's eps prediction GIVEN THE ORACLE'S x_t (so errors do not compound), the error of x_{t-1} after the
update, error percentiles, and for the worst pixel whether pred_xstart sits on the clip boundary.

    python scripts/diag_outlier.py            # the smoke configuration
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults  # noqa: E402
from autodiffusion_b200.sampler import sample_candidate  # noqa: E402
from oracle import diffusion_ref, unet_ref, weights  # noqa: E402


def pct(err, qs=(50, 90, 99, 99.9, 100)):
    e = err.abs().flatten().double().numpy()
    return " ".join(f"p{q}={np.percentile(e, q):.4g}" for q in qs)


def main():
    flags = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=64, num_head_channels=64, num_res_blocks=1,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=True, use_scale_shift_norm=True,
                 use_dynamic_unet=True)
    d = model_and_diffusion_defaults()
    d.update(flags)
    model, diffusion = create_model_and_diffusion(**d)
    cfg = unet_ref.UNetConfig(model_channels=64, num_res_blocks=1)
    sd = weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0)
    model.load_state_dict(sd)
    model.to("cuda:0").eval()
    cand = {"timesteps": [85, 971], "skip_layers": [[0], [3, 7]]}
    B = 2
    noise = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(3))
    out = sample_candidate(model, diffusion, cand, (B, 3, 64, 64), noise.cuda(), y.cuda()).cpu()

    base = diffusion_ref.base_tables("cosine", 1000)
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], cand["timesteps"])
    tb = diffusion_ref.diffusion_tables(nb)
    unet = lambda x, t, yy, skip: unet_ref.unet_forward(sd, cfg, x, t, yy, skip)
    refs = diffusion_ref.ddim_sample_loop(diffusion_ref.make_model_fn(unet, tmap), noise.shape, tb, tmap, noise, True,
                                          model_kwargs={"y": y, "skip_layers": cand["skip_layers"]}, return_all=True)
    err = out - refs[-1]
    print(f"final: max_abs={err.abs().max().item():.4g} rms={err.pow(2).mean().sqrt().item():.4g} {pct(err)}")
    K = len(tmap)
    for n, i in enumerate(range(K)[::-1]):
        t_orig = tmap[i]
        A, Bm = float(np.float32(tb["sqrt_recip_alphas_cumprod"][i])), float(np.float32(tb["sqrt_recipm1_alphas_cumprod"][i]))
        x_t = refs[n]
        tt = torch.full((B,), t_orig, dtype=torch.long)
        with torch.no_grad():
            eps_ref = unet(x_t, tt, y, cand["skip_layers"][i])[:, :3]
        eps = model(x_t.cuda(), tt.cuda(), y.cuda(), skip_layer=cand["skip_layers"][i]).cpu()[:, :3]
        e_eps = eps - eps_ref
        x0_ref = A * x_t - Bm * eps_ref
        x0 = A * x_t - Bm * eps
        inside = (x0_ref.abs() < 1.0)
        e_x0 = (x0.clamp(-1, 1) - x0_ref.clamp(-1, 1))
        # one fused step from the oracle's x_t: isolates this step's contribution to x_{t-1}
        from autodiffusion_b200 import ops
        from autodiffusion_b200.gaussian_diffusion import ddim_coefficients
        coef = ddim_coefficients(tb, i)
        mo = model(x_t.cuda(), tt.cuda(), y.cuda(), skip_layer=cand["skip_layers"][i])
        x_prev = ops.ddim_step(x_t.cuda().contiguous(), mo.contiguous(), None, coef, True).cpu()
        e_prev = x_prev - refs[n + 1]
        w = e_prev.abs().flatten().argmax().item()
        print(f"step {n} (t={t_orig}): A={A:.4g} Bm={Bm:.4g}  eps err rms={e_eps.pow(2).mean().sqrt().item():.4g} "
              f"max={e_eps.abs().max().item():.4g} (eps std {eps_ref.std().item():.3g})")
        print(f"    x0 unclipped fraction {inside.float().mean().item():.3f}; clipped-x0 err {pct(e_x0)}")
        print(f"    x_(t-1) err from this step alone: {pct(e_prev)}")
        print(f"    worst pixel: x0_ref={x0_ref.flatten()[w].item():.4f} x0={x0.flatten()[w].item():.4f} "
              f"eps err there={e_eps.flatten()[w].item():.4g} -> Bm*err={Bm * e_eps.flatten()[w].item():.4g}")


if __name__ == "__main__":
    main()
