"""CPU: the timestep-only search drivers (autodiffusion_b200.timestep_search.TimestepSearcher) against golden traces of the
unmodified reference scripts GD/search_imagenet64_classifier_guidance.py and GD/search_uncondition_model.py
(tests/golden/timestep_search_trace.json, made by make_timestep_search_golden.py with a stubbed FID): same seeds ->
the same individuals in the same order, the same top list, the same log. Integer / index work: exact."""
import json
import os
import random
import types

import numpy as np
import pytest

from tests.test_search_cpu import StubEvaluator
from tests.util import GOLDEN

TRACE = json.load(open(os.path.join(GOLDEN, "timestep_search_trace.json")))


def build(g, evaluator, log):
    from autodiffusion_b200.timestep_search import TimestepSearcher, build_search_space

    cfg = g["config"]
    args = types.SimpleNamespace(**{k: v for k, v in cfg.items() if k != "search_space_core"}, batch_size=4, num_samples=8,
                                 image_size=64)
    model = types.SimpleNamespace(layer_num=58)
    diffusion = types.SimpleNamespace(original_num_steps=cfg["original_num_steps"])
    space = None
    if cfg.get("search_space_core"):
        from autodiffusion_b200.respace import space_timesteps

        init = list(space_timesteps(cfg["original_num_steps"], "ddim" + str(cfg["time_step"]))) if cfg["use_ddim_init_x"] else None
        space = build_search_space(cfg["search_space_core"], cfg["original_num_steps"], init)
        assert space == g["search_space"]  # the window construction of the script's __main__ (:645-668)
    return TimestepSearcher(args, model, diffusion, cfg["time_step"], classifier=None, search_space=space,
                            variant=g["variant"], evaluator=evaluator, log=log)


@pytest.mark.parametrize("name", sorted(TRACE))
def test_timestep_search_reproduces_the_reference_trace(name):
    g = TRACE[name]
    lines = []
    ev = StubEvaluator(False)
    s = build(g, ev, lines.append)
    random.seed(g["config"]["seed"])
    np.random.seed(g["config"]["seed"])
    top = s.search()
    assert list(s.vis_dict.keys()) == g["visited"]
    assert [s.vis_dict[k]["fid"] for k in s.vis_dict] == g["fids"]
    assert top == g["top"] and s.epoch == g["epoch"]
    assert ev.calls == g["visited"]
    assert lines[:12] == g["log_head"] and len(lines) == g["n_log"]
    assert all(isinstance(eval(c), list) for c in s.vis_dict)  # individuals are bare timestep lists


def test_timestep_search_resumes_from_a_saved_state(tmp_path):
    g = TRACE["imagenet64_search_space"]
    path = str(tmp_path / "ts.pkl")
    cfg = g["config"]
    s1 = build(dict(g, config=dict(cfg, max_epochs=2)), StubEvaluator(False), lambda l: None)
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    try:  # die in the middle of epoch 1's mutation phase, after its state was written
        orig = s1.get_cross
        s1.get_cross = lambda *a, **k: (_ for _ in ()).throw(KeyboardInterrupt()) if s1.epoch == 1 else orig(*a, **k)
        s1.search(state_path=path)
    except KeyboardInterrupt:
        pass
    ev2 = StubEvaluator(False)
    s2 = build(g, ev2, lambda l: None)
    s2.load_state(path)
    top = s2.search(state_path=path)
    assert list(s2.vis_dict.keys()) == g["visited"] and top == g["top"] and s2.epoch == g["epoch"]
