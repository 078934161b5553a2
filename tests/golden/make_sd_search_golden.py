"""Golden trace of the Stable-Diffusion search script's evolutionary operators (authoring container only).

Imports the UNMODIFIED `scripts/search_ea.py` from /root/reference/examples/"Stable Diffusion" (modules absent offline -
pytorch_lightning, omegaconf, pytorch_fid, the dataloader builder - are stubbed: none is touched by the operators),
builds its EvolutionSearcher without running __init__ (which loads reference statistics), stubs `get_cand_fid` with a
deterministic function of the candidate, runs `search()` under fixed seeds and records every individual visited, the
final top list and the log. tests/test_sd_search_cpu.py replays the same seeds through autodiffusion_b200.sd_search.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_sd_search_golden.py
"""
import importlib.util
import json
import logging
import os
import random
import sys
import types
import zlib

import numpy as np

REF = "/root/reference/examples/Stable Diffusion"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.dont_write_bytecode = True


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return m


_stub("pytorch_lightning", seed_everything=lambda s: None)
_stub("omegaconf", OmegaConf=type("OmegaConf", (), {}))
_stub("omegaconf.listconfig", ListConfig=type("ListConfig", (), {}))
_stub("pytorch_fid")
_stub("pytorch_fid.inception", InceptionV3=type("InceptionV3", (), {}))
_stub("ldm.data.build_dataloader", build_dataloader=lambda *a, **k: None)
_stub("ldm.util", instantiate_from_config=lambda *a, **k: None)


def stub_fid(cand) -> float:
    return (zlib.crc32(str(cand).encode()) % 100000) / 1000.0


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_sd_search", os.path.join(REF, "scripts", "search_ea.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(mod, cfg, dpm_params=None):
    s = object.__new__(mod.EvolutionSearcher)
    s.opt = types.SimpleNamespace(**cfg)
    s.sampler = types.SimpleNamespace(ddpm_num_timesteps=cfg["ddpm_num_timesteps"])
    s.time_step = cfg["time_step"]
    s.max_epochs, s.select_num, s.population_num = cfg["max_epochs"], cfg["select_num"], cfg["population_num"]
    s.m_prob, s.crossover_num, s.mutation_num = cfg["m_prob"], cfg["crossover_num"], cfg["mutation_num"]
    s.ddim_discretize = "uniform"
    s.keep_top_k = {s.select_num: [], 50: []}
    s.epoch, s.candidates, s.vis_dict = 0, [], {}
    s.use_ddim_init_x = cfg["use_ddim_init_x"]
    s.dpm_params = dpm_params
    s.get_cand_fid = lambda cand=None, opt=None, device="cuda": stub_fid(cand)
    lines = []
    handler = logging.Handler()
    handler.emit = lambda rec: lines.append(rec.getMessage())
    root = logging.getLogger()
    root.addHandler(handler)
    root.setLevel(logging.INFO)
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    try:
        s.search()
    finally:
        root.removeHandler(handler)
    return {"config": cfg, "dpm_params": dpm_params, "visited": list(s.vis_dict.keys()),
            "fids": [s.vis_dict[k]["fid"] for k in s.vis_dict], "top": s.keep_top_k[50], "epoch": s.epoch, "log": lines}


if __name__ == "__main__":
    import torch

    mod = load_ref()
    base = dict(ddpm_num_timesteps=1000, time_step=4, max_epochs=6, select_num=4, population_num=12, m_prob=0.25,
                crossover_num=4, mutation_num=5, seed=0, use_ddim_init_x=False, dpm_solver=False)
    # the dpm_params main() builds (search_ea.py:889-902): 1001 / time_step + 1 uniform time points between 1 and 1/1000
    full = [v.item() for v in list(torch.linspace(1.0, 0.001, 1001))]
    init = [v.item() for v in list(torch.linspace(1.0, 0.001, 4 + 1))]
    out = {
        "random_init": run(mod, base),
        "ddim_init": run(mod, dict(base, use_ddim_init_x=True, seed=3, time_step=5)),
        "dpm": run(mod, dict(base, use_ddim_init_x=True, seed=5, dpm_solver=True, max_epochs=4),
                   dpm_params={"full_timesteps": full, "init_timesteps": init}),
    }
    for k, v in out.items():
        print(k, "visited", len(v["visited"]), "log lines", len(v["log"]), "top0", v["top"][0])
    json.dump(out, open(os.path.join(HERE, "sd_search_trace.json"), "w"))
