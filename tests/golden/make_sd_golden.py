"""Golden fixtures for the Stable-Diffusion-v1 family (BASELINE configs[4]) from the UNMODIFIED reference
(`/root/reference/examples/Stable Diffusion`, imported read-only), and the check that oracle/sd_unet_ref.py
restates it exactly.

Run in the authoring container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_sd_golden.py
Outputs (committed): tests/golden/sd_small.npz, tests/golden/sd_full.npz. Every array in them was produced by
reference code on weights from oracle.sd_unet_ref.make_weights.

Shims (none touch the arithmetic): `omegaconf.listconfig.ListConfig` stub (one isinstance,
openaimodel.py:476); `DDIMSampler.register_buffer` hard-codes `.to("cuda")` (ddim.py:19-23) and is replaced
by a plain setattr so the sampler runs on the CPU; LatentDiffusion (needs pytorch_lightning) is replaced by a
holder with the attributes the sampler reads (num_timesteps, betas, alphas_cumprod, alphas_cumprod_prev, device,
apply_model), its schedule built by the reference's own make_beta_schedule as DDPM.register_schedule does
(ldm/models/diffusion/ddpm.py:117-134).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/examples/Stable Diffusion"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
oc = types.ModuleType("omegaconf")
ocl = types.ModuleType("omegaconf.listconfig")
ocl.ListConfig = type("ListConfig", (), {})
oc.listconfig = ocl
sys.modules.setdefault("omegaconf", oc)
sys.modules.setdefault("omegaconf.listconfig", ocl)

from ldm.models.diffusion.ddim import DDIMSampler  # noqa: E402
from ldm.models.diffusion.plms import PLMSSampler  # noqa: E402
from ldm.modules.diffusionmodules.openaimodel import UNetModel  # noqa: E402
from ldm.modules.diffusionmodules.util import make_beta_schedule  # noqa: E402

from oracle import sd_unet_ref as R  # noqa: E402

DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
PLMSSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)

SMALL = R.SDConfig(model_channels=64, context_dim=128)
CAND10 = [981, 861, 741, 641, 501, 421, 301, 201, 121, 21]  # a searched-style 10-step subsequence (unsorted input below)


def ref_unet(cfg: R.SDConfig, sd):
    m = UNetModel(image_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                  model_channels=cfg.model_channels, attention_resolutions=list(cfg.attention_resolutions),
                  num_res_blocks=cfg.num_res_blocks, channel_mult=list(cfg.channel_mult), num_heads=cfg.num_heads,
                  use_spatial_transformer=True, transformer_depth=cfg.transformer_depth, context_dim=cfg.context_dim,
                  use_checkpoint=False, legacy=False)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_shapes == R.param_shapes(cfg), "oracle param_shapes differ from the reference state_dict"
    m.load_state_dict(sd)
    return m.eval()


class Holder:
    """What DDIMSampler reads from LatentDiffusion."""

    def __init__(self, unet):
        betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.0120)
        acp = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = 1000
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(acp, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, acp[:-1]), dtype=torch.float32)
        self.device = torch.device("cpu")
        self.unet = unet
        self.calls = []

    def apply_model(self, x, t, c):
        self.calls.append(int(t[0]))
        return self.unet(x, t, context=c)


def plms_golden():
    """PLMS (the sampler search_plms.sh uses): 6 searched steps so every Adams-Bashforth order occurs, CFG 7.5."""
    g = torch.Generator().manual_seed(23)
    sd = R.make_weights(SMALL, seed=0)
    m = ref_unet(SMALL, sd)
    holder = Holder(m)
    cand = [641, 981, 201, 21, 861, 421]
    x_T = torch.randn(1, 4, 64, 64, generator=g)  # batch 1: the CPU replay of this fixture runs in every test session
    ctx = torch.randn(1, 77, SMALL.context_dim, generator=g)
    uc = torch.randn(1, 77, SMALL.context_dim, generator=g)
    sampler = PLMSSampler(holder)
    samples, _ = sampler.sample(S=len(cand), conditioning=ctx, batch_size=1, shape=[4, 64, 64], verbose=False,
                                unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                                sampled_timestep=np.array(cand))
    mine = R.plms_sample(lambda xx, tt, cc: R.unet_forward(sd, SMALL, xx, tt, cc), x_T, ctx, uc, 7.5, cand, R.sd_alphas_cumprod())
    d = (samples - mine).abs().max().item()
    print("small CFG-PLMS: max|ref - oracle| =", d, "model calls at", holder.calls)
    assert d <= 1e-4 * samples.abs().max().item()
    # the multistep combinations on recorded eps tensors: the fused update's bit-exactness fixture
    es = [torch.randn(1, 4, 64, 64, generator=g) for _ in range(4)]

    class Fixed:
        num_timesteps = 1000
        betas, alphas_cumprod, alphas_cumprod_prev, device = holder.betas, holder.alphas_cumprod, holder.alphas_cumprod_prev, holder.device

        def __init__(self):
            self.k = 0

        def apply_model(self, xx, tt, cc):
            self.k += 1
            return es[0] if self.k == 1 else es[1]  # e_t, then (first step only) e_t_next

    out = {}
    for n_old in range(4):
        fx = Fixed()
        s2 = PLMSSampler(fx)
        s2.make_schedule(ddim_num_steps=len(cand), ddim_eta=0.0, verbose=False, sampled_timestep=sorted(cand))
        old = [es[3 - k] for k in range(n_old)][::-1]  # old_eps list, newest last: [.., es[3]]; here es[3], es[2], es[1]... newest = es[3]
        xp, x0, e_t = s2.p_sample_plms(x_T, ctx, torch.full((1,), 421), index=2, unconditional_guidance_scale=1.0,
                                       unconditional_conditioning=None, old_eps=list(old), t_next=torch.full((1,), 201))
        assert torch.equal(e_t, es[0])
        out[f"plms_x_prev_{n_old}"] = xp.numpy()
    np.savez_compressed(os.path.join(HERE, "sd_small_plms.npz"), cand=np.array(cand), x_T=x_T.numpy(), ctx=ctx.numpy(), uc=uc.numpy(),
                        samples=samples.numpy(), calls=np.array(holder.calls), es=torch.stack(es).numpy(), **out)


def dpm_golden():
    """DPM-Solver++(2M) with searched steps (search_dpm_solver.sh): 6 model evaluations (7 time points), CFG 7.5.
    Extra shim: the solver moves its time-step tensor with a hard-coded `.to('cuda')` (dpm_solver.py:1088,1091);
    during this run `Tensor.to('cuda')` is a no-op so the unmodified code runs on the CPU."""
    from ldm.models.diffusion.dpm_solver import DPMSolverSampler
    from ldm.models.diffusion.dpm_solver.dpm_solver import DPM_Solver, NoiseScheduleVP, model_wrapper

    DPMSolverSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(v, str) and v == "cuda") else v for v in a)
        return orig_to(self, *a, **k)

    g = torch.Generator().manual_seed(29)
    sd = R.make_weights(SMALL, seed=0)
    m = ref_unet(SMALL, sd)

    class FloatHolder(Holder):
        def apply_model(self, x, t, c):
            self.calls.append(float(t[0]))
            return self.unet(x, t, context=c)

    holder = FloatHolder(m)
    cand = [981, 861, 641, 421, 201, 61, 0]  # descending, as the SD search keeps them (time_step + 1 entries)
    x_T = torch.randn(1, 4, 64, 64, generator=g)
    ctx = torch.randn(1, 77, SMALL.context_dim, generator=g)
    uc = torch.randn(1, 77, SMALL.context_dim, generator=g)
    torch.Tensor.to = to_cpu
    try:
        sampler = DPMSolverSampler(holder)
        samples, _ = sampler.sample(S=len(cand) - 1, conditioning=ctx, batch_size=1, shape=[4, 64, 64], verbose=False,
                                    unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                                    sampled_timestep=cand)
        # schedule scalars of the reference at the candidate's time points
        ns = NoiseScheduleVP("discrete", alphas_cumprod=holder.alphas_cumprod)
        tt = R.dpm_timesteps(cand)
        ref_lam, ref_alpha, ref_std = ns.marginal_lambda(tt), ns.marginal_alpha(tt), ns.marginal_std(tt)
    finally:
        torch.Tensor.to = orig_to
    rec = []
    mine = R.dpm_solver_sample(lambda xx, t, cc: R.unet_forward(sd, SMALL, xx, t, cc), x_T, ctx, uc, 7.5, cand,
                               R.sd_alphas_cumprod(), record=rec)
    d = (samples - mine).abs().max().item()
    print("small CFG DPM-Solver++(2M): max|ref - oracle| =", d, "model times", holder.calls)
    assert d <= 1e-4 * samples.abs().max().item()
    vp = R.DiscreteVP(R.sd_alphas_cumprod())
    assert torch.equal(vp.lam(tt), ref_lam) and torch.equal(vp.alpha(tt), ref_alpha) and torch.equal(vp.std(tt), ref_std)
    assert np.allclose(rec, holder.calls, rtol=0, atol=0)
    np.savez_compressed(os.path.join(HERE, "sd_small_dpm.npz"), cand=np.array(cand), x_T=x_T.numpy(), ctx=ctx.numpy(), uc=uc.numpy(),
                        samples=samples.numpy(), calls=np.array(holder.calls), times=tt.numpy(), lam=ref_lam.numpy(),
                        alpha=ref_alpha.numpy(), std=ref_std.numpy())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "plms":
        return plms_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "dpm":
        return dpm_golden()
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(11)
    # ---- small config: forward + CFG DDIM ----
    sd = R.make_weights(SMALL, seed=0)
    m = ref_unet(SMALL, sd)
    x = torch.randn(2, 4, 64, 64, generator=g)
    t = torch.tensor([981, 21])
    ctx = torch.randn(2, 77, SMALL.context_dim, generator=g)
    with torch.no_grad():
        out = m(x, t, context=ctx)
    mine = R.unet_forward(sd, SMALL, x, t, ctx)
    d_small = (out - mine).abs().max().item()
    print("small forward: max|ref - oracle| =", d_small, " out std", out.std().item())
    assert d_small <= 1e-5 * out.abs().max().item()

    holder = Holder(m)
    assert torch.equal(holder.alphas_cumprod, R.sd_alphas_cumprod())
    sampler = DDIMSampler(holder)
    cand = [CAND10[i] for i in (3, 0, 7, 9, 1, 5, 2, 8, 4, 6)][:4]  # 4 unsorted searched steps
    x_T = torch.randn(2, 4, 64, 64, generator=g)
    uc = torch.randn(1, 77, SMALL.context_dim, generator=g).repeat(2, 1, 1)
    samples, _ = sampler.sample(S=len(cand), conditioning=ctx, batch_size=2, shape=[4, 64, 64], verbose=False,
                                unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=0.0, x_T=x_T,
                                sampled_timestep=np.array(cand))
    mine_s = R.ddim_sample(lambda xx, tt, cc: R.unet_forward(sd, SMALL, xx, tt, cc), x_T, ctx, uc, 7.5, cand,
                           R.sd_alphas_cumprod())
    d_s = (samples - mine_s).abs().max().item()
    print("small CFG-DDIM: max|ref - oracle| =", d_s, "steps seen", holder.calls)
    assert d_s <= 1e-4 * samples.abs().max().item()
    steps, alphas, alphas_prev, s1m = R.ddim_tables(R.sd_alphas_cumprod(), cand)
    assert np.array_equal(np.asarray(sampler.ddim_alphas), alphas.numpy())
    assert np.array_equal(np.asarray(sampler.ddim_alphas_prev, dtype=np.float32), alphas_prev.numpy())
    assert np.array_equal(np.asarray(sampler.ddim_sqrt_one_minus_alphas), s1m.numpy())
    # one exact step on recorded eps: the fused update's bit-exactness fixture
    e_u, e_c = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 4, 64, 64, generator=g)

    class Fixed:
        num_timesteps = 1000
        betas, alphas_cumprod, alphas_cumprod_prev, device = holder.betas, holder.alphas_cumprod, holder.alphas_cumprod_prev, holder.device

        def apply_model(self, xx, tt, cc):
            return torch.cat([e_u, e_c])

    s2 = DDIMSampler(Fixed())
    s2.make_schedule(ddim_num_steps=len(cand), ddim_eta=0.0, verbose=False, sampled_timestep=sorted(cand))
    step_out = {}
    for index in range(len(cand)):
        xp, x0 = s2.p_sample_ddim(x_T, ctx, torch.full((2,), steps[index]), index=index,
                                  unconditional_guidance_scale=7.5, unconditional_conditioning=uc)
        mx, m0 = R.ddim_step(x_T, e_u + 7.5 * (e_c - e_u), alphas[index], alphas_prev[index], s1m[index])
        assert torch.equal(xp, mx) and torch.equal(x0, m0), "ddim_step is not bit-exact vs p_sample_ddim"
        step_out[f"step_x_prev_{index}"] = xp.numpy()
    print("p_sample_ddim on fixed eps: oracle bit-exact for", len(cand), "indices")
    np.savez_compressed(os.path.join(HERE, "sd_small.npz"), x=x.numpy(), t=t.numpy(), ctx=ctx.numpy(), out=out.numpy(),
                        cand=np.array(cand), x_T=x_T.numpy(), uc=uc.numpy(), samples=samples.numpy(),
                        steps_seen=np.array(holder.calls), e_u=e_u.numpy(), e_c=e_c.numpy(),
                        ddim_alphas=alphas.numpy(), ddim_alphas_prev=alphas_prev.numpy(), ddim_s1m=s1m.numpy(), **step_out)

    # ---- full SD-v1 UNet (859.5 M parameters): one forward ----
    full = R.sd_v1_config()
    sdf = R.make_weights(full, seed=0)
    nparam = sum(v.numel() for v in sdf.values())
    mf = ref_unet(full, sdf)
    xf = torch.randn(1, 4, 64, 64, generator=g)
    tf_ = torch.tensor([501])
    cf = torch.randn(1, 77, 768, generator=g)
    with torch.no_grad():
        of = mf(xf, tf_, context=cf)
    del mf
    minef = R.unet_forward(sdf, full, xf, tf_, cf)
    d_full = (of - minef).abs().max().item()
    print(f"full forward ({nparam / 1e6:.1f} M params): max|ref - oracle| = {d_full}, out std {of.std().item():.4f}")
    assert d_full <= 1e-5 * of.abs().max().item()
    plms_golden()
    dpm_golden()
    np.savez_compressed(os.path.join(HERE, "sd_full.npz"), x=xf.numpy(), t=tf_.numpy(), ctx=cf.numpy(), out=of.numpy(),
                        nparam=np.array(nparam))


if __name__ == "__main__":
    main()
