"""CPU: oracle/sd_unet_ref.py (the SD-v1 UNet + searched-timestep CFG DDIM restatement) reproduces the outputs of the
unmodified reference (tests/golden/sd_small.npz, sd_full.npz, made by tests/golden/make_sd_golden.py).
Tolerances: the schedule tables and the update step on recorded eps are bit-exact; the fp32 network outputs agree to
1e-5 of the output range (same torch build: measured 0.0)."""
import numpy as np
import torch

from oracle import sd_unet_ref as R
from tests.util import golden

SMALL = R.SDConfig(model_channels=64, context_dim=128)


def test_sd_param_inventory():
    shapes = R.param_shapes(R.sd_v1_config())
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == int(golden("sd_full.npz")["nparam"]) == 859520964  # SD-v1 UNet of v1-inference_coco.yaml:29-44
    arch = R.build_arch(R.sd_v1_config())
    n_st = sum(b.kind == "st" for layers in arch.input_blocks + [arch.middle] + arch.output_blocks for b in layers)
    assert n_st == 16 and len(arch.input_blocks) == 12 and len(arch.output_blocks) == 12


def test_sd_small_forward_matches_reference():
    g = golden("sd_small.npz")
    sd = R.make_weights(SMALL, seed=0)
    out = R.unet_forward(sd, SMALL, torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["ctx"]))
    ref = torch.tensor(g["out"])
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_sd_schedule_tables_bit_exact():
    g = golden("sd_small.npz")
    steps, a, ap, s1m = R.ddim_tables(R.sd_alphas_cumprod(), g["cand"].tolist())
    assert steps == sorted(g["cand"].tolist()) == sorted(set(g["steps_seen"].tolist()))
    assert g["steps_seen"].tolist() == sorted(g["cand"].tolist(), reverse=True)  # the sampler walks them descending
    assert np.array_equal(a.numpy(), g["ddim_alphas"])
    assert np.array_equal(ap.numpy(), g["ddim_alphas_prev"])
    assert np.array_equal(s1m.numpy(), g["ddim_s1m"])


def test_sd_cfg_ddim_step_bit_exact_on_recorded_eps():
    g = golden("sd_small.npz")
    steps, a, ap, s1m = R.ddim_tables(R.sd_alphas_cumprod(), g["cand"].tolist())
    e_u, e_c, x = torch.tensor(g["e_u"]), torch.tensor(g["e_c"]), torch.tensor(g["x_T"])
    for index in range(len(steps)):
        xp, _ = R.ddim_step(x, e_u + 7.5 * (e_c - e_u), a[index], ap[index], s1m[index])
        assert np.array_equal(xp.numpy(), g[f"step_x_prev_{index}"])


def test_sd_cfg_ddim_sampling_matches_reference():
    g = golden("sd_small.npz")
    sd = R.make_weights(SMALL, seed=0)
    out = R.ddim_sample(lambda x, t, c: R.unet_forward(sd, SMALL, x, t, c), torch.tensor(g["x_T"]), torch.tensor(g["ctx"]),
                        torch.tensor(g["uc"]), 7.5, g["cand"].tolist(), R.sd_alphas_cumprod())
    ref = torch.tensor(g["samples"])
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_sd_plms_sampling_matches_reference():
    g = golden("sd_small_plms.npz")
    sd = R.make_weights(SMALL, seed=0)
    calls = []

    def model(x, t, c):
        calls.append(int(t[0]))
        return R.unet_forward(sd, SMALL, x, t, c)

    out = R.plms_sample(model, torch.tensor(g["x_T"]), torch.tensor(g["ctx"]), torch.tensor(g["uc"]), 7.5, g["cand"].tolist(),
                        R.sd_alphas_cumprod())
    ref = torch.tensor(g["samples"])
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert calls == g["calls"].tolist()  # 6 steps + the pseudo-improved-Euler extra call at the second timestep


def test_sd_dpm_solver_sampling_matches_reference():
    g = golden("sd_small_dpm.npz")
    sd = R.make_weights(SMALL, seed=0)
    rec = []
    out = R.dpm_solver_sample(lambda x, t, c: R.unet_forward(sd, SMALL, x, t, c), torch.tensor(g["x_T"]), torch.tensor(g["ctx"]),
                              torch.tensor(g["uc"]), 7.5, g["cand"].tolist(), R.sd_alphas_cumprod(), record=rec)
    ref = torch.tensor(g["samples"])
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert np.array_equal(np.float32(rec), g["calls"])
