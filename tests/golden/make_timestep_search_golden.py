"""Golden traces of the reference's timestep-only evolutionary searches (run in the authoring container only).

Imports the UNMODIFIED scripts GD/search_imagenet64_classifier_guidance.py and GD/search_uncondition_model.py from
/root/reference, builds their EvolutionSearcher without running __init__ (TensorFlow, pickled Inception statistics), stubs
`get_cand_fid` with a deterministic function of the candidate, runs `search()` under fixed seeds and records every
individual the operators produced, the top list and the log. tests/test_timestep_search_cpu.py replays the same seeds
through autodiffusion_b200.timestep_search.TimestepSearcher and must reproduce the traces exactly.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_timestep_search_golden.py
"""
import importlib.util
import json
import os
import random
import sys
import types
import zlib

import numpy as np

REF = "/root/reference/examples/guided_diffusion"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.modules.setdefault("blobfile", types.ModuleType("blobfile"))
sys.dont_write_bytecode = True


def stub_fid(cand) -> float:
    return (zlib.crc32(str(cand).encode()) % 100000) / 1000.0


def load_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def search_space_of(mod_src_cfg, n_steps):
    """The `__main__` block of search_imagenet64_classifier_guidance.py:645-668, executed on the given core steps."""
    core, use_init, time_step = mod_src_cfg
    from guided_diffusion.respace import space_timesteps

    search_space_core = sorted(core)
    if use_init:
        search_space_core += list(space_timesteps(n_steps, "ddim" + str(time_step)))
    R = int(n_steps / 100)
    search_space = []
    for s in search_space_core:
        left = max(s - R, 0)
        right = min(s + R, n_steps)
        search_space += [i for i in range(left, right)]
    return sorted(list(set(search_space)))


def run(mod, cfg, variant):
    S = mod.EvolutionSearcher
    s = object.__new__(S)
    args = types.SimpleNamespace(**{k: v for k, v in cfg.items() if k != "search_space_core"})
    mod.args = args  # search() reads the module-level `args`
    s.args = args

    class _Diff:
        original_num_steps = cfg["original_num_steps"]

    s.model, s.base_diffusion, s.classifier = object(), _Diff(), None
    s.time_step = cfg["time_step"]
    s.max_epochs, s.select_num, s.population_num = cfg["max_epochs"], cfg["select_num"], cfg["population_num"]
    s.m_prob, s.crossover_num, s.mutation_num = cfg["m_prob"], cfg["crossover_num"], cfg["mutation_num"]
    s.keep_top_k = {s.select_num: [], 50: []}
    s.epoch, s.candidates, s.vis_dict = 0, [], {}
    s.max_fid, s.thres = 48.0, 0.2
    s.rf_features, s.rf_lebal = [], []
    space = None
    if variant == "imagenet64":
        if cfg.get("search_space_core"):
            space = search_space_of((cfg["search_space_core"], cfg["use_ddim_init_x"], cfg["time_step"]), cfg["original_num_steps"])
        s.search_space = space
    else:
        s.x0 = cfg.get("init_x", "")
    s.get_cand_fid = lambda cand=None, args=None: stub_fid(cand)
    lines = []
    mod.logger.log = lambda *a, **k: lines.append(" ".join(str(x) for x in a))
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    space0 = None if space is None else list(space)
    s.search()
    return {"config": cfg, "variant": variant, "search_space": space0, "visited": list(s.vis_dict.keys()),
            "fids": [s.vis_dict[k]["fid"] for k in s.vis_dict], "top": s.keep_top_k[50], "epoch": s.epoch,
            "log_head": lines[:12], "n_log": len(lines)}


if __name__ == "__main__":
    img = load_ref("search_imagenet64_classifier_guidance")
    unc = load_ref("search_uncondition_model")
    base = dict(original_num_steps=1000, time_step=4, max_epochs=5, select_num=4, population_num=12, m_prob=0.25,
                crossover_num=4, mutation_num=5, seed=0, use_ddim_init_x=False, use_ddim=True)
    out = {
        "imagenet64_random": run(img, base, "imagenet64"),
        "imagenet64_ddim_init": run(img, dict(base, use_ddim_init_x=True, seed=3, time_step=5), "imagenet64"),
        "imagenet64_search_space": run(img, dict(base, seed=5, use_ddim_init_x=True, search_space_core=[926, 153, 424, 690, 5]),
                                       "imagenet64"),
        "uncondition_random": run(unc, dict(base, seed=1, init_x=""), "uncondition"),
        "uncondition_ddim_init": run(unc, dict(base, seed=2, use_ddim_init_x=True, init_x="", time_step=6), "uncondition"),
        "uncondition_init_x": run(unc, dict(base, seed=4, init_x="[644, 737, 67, 804]"), "uncondition"),
    }
    with open(os.path.join(HERE, "timestep_search_trace.json"), "w") as f:
        json.dump(out, f)
    for k, v in out.items():
        print(k, "visited", len(v["visited"]), "epochs", v["epoch"], "n_log", v["n_log"], "top1", v["top"][0][:60])
