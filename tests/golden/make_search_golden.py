"""Golden trace of the reference's evolutionary operators (run in the authoring container only).

Imports the UNMODIFIED reference search script from /root/reference, builds its EvolutionSearcher without
running __init__ (which needs TensorFlow and a pickled Inception reference), stubs `get_cand_fid` with a
deterministic function of the candidate, runs `search()` under fixed seeds and records, in order, every
individual the operators produced and the final top list. tests/test_search_cpu.py replays the same
seeds through autodiffusion_b200.search.EvolutionSearcher and must reproduce the trace exactly.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_search_golden.py
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


def load_ref():
    spec = importlib.util.spec_from_file_location(
        "ref_search", os.path.join(REF, "search_dynamic_unet_imagenet64_classifier_guidance_progressive.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(mod, cfg):
    S = mod.EvolutionSearcher
    s = object.__new__(S)
    args = types.SimpleNamespace(**cfg)
    mod.args = args  # search() reads the module-level `args` (:642-648)
    s.args = args

    class _Model:
        layer_num = cfg["layer_num"]

    class _Diff:
        original_num_steps = cfg["original_num_steps"]

    s.model, s.base_diffusion, s.classifier = _Model(), _Diff(), None
    s.init_time_step = cfg["time_step"]
    s.max_index_number = cfg["time_step"] * cfg["layer_num"]
    s.max_epochs, s.select_num, s.population_num = cfg["max_epochs"], cfg["select_num"], cfg["population_num"]
    s.m_prob, s.crossover_num, s.mutation_num = cfg["m_prob"], cfg["crossover_num"], cfg["mutation_num"]
    s.keep_top_k = {s.select_num: [], 50: []}
    s.epoch, s.candidates, s.vis_dict = 0, [], {}
    s.max_fid, s.max_prun, s.min_prun = 48.0, cfg["max_prun"], cfg["min_prun"]
    s.model_layers = cfg["layer_num"]
    s.skip_layer_range = [0, 0]
    s.last_best_cand = None
    s.get_cand_fid = lambda cand=None, args=None: stub_fid(cand)
    lines = []
    mod.logger.log = lambda *a, **k: lines.append(" ".join(str(x) for x in a))
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    s.search()
    return {"config": cfg, "visited": list(s.vis_dict.keys()), "fids": [s.vis_dict[k]["fid"] for k in s.vis_dict],
            "top": s.keep_top_k[50], "skip_layer_range": s.skip_layer_range, "epoch": s.epoch,
            "log_head": lines[:12], "n_log": len(lines)}


if __name__ == "__main__":
    mod = load_ref()
    base = dict(layer_num=58, original_num_steps=1000, time_step=4, max_epochs=9, select_num=4, population_num=12,
                m_prob=0.25, crossover_num=4, mutation_num=5, max_prun=0.2, min_prun=0.05, seed=0,
                use_ddim_init_x=False, use_ddim=True)
    out = {"random_init": run(mod, base),
           "ddim_init": run(mod, dict(base, use_ddim_init_x=True, seed=3, max_epochs=4, time_step=5))}
    with open(os.path.join(HERE, "search_trace.json"), "w") as f:
        json.dump(out, f)
    for k, v in out.items():
        print(k, "visited", len(v["visited"]), "epochs", v["epoch"], "range", v["skip_layer_range"], "top1", v["top"][0][:80])
