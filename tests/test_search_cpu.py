"""CPU: the evolutionary search driver (autodiffusion_b200.search.EvolutionSearcher) against a golden trace of
the reference's own operators (tests/golden/search_trace.json, made by make_search_golden.py from the
unmodified reference script with a stubbed FID): under the same seeds it must visit exactly the same
individuals in the same order and end with the same top list and prune range. Integer / index work:
bit-exact. Also: deferred FIDs, log-line format, save/resume."""
import json
import os
import random
import types
import zlib
from concurrent.futures import Future

import numpy as np

from tests.util import GOLDEN


def stub_fid(cand) -> float:
    return (zlib.crc32(str(cand).encode()) % 100000) / 1000.0


class StubEvaluator:
    """Stands in for CandidateEvaluator: FID = a deterministic function of the candidate, optionally deferred."""

    def __init__(self, deferred):
        self.deferred = deferred
        self.calls = []
        self.unresolved = []

    def get_cand_fid(self, cand=None, args=None):
        self.calls.append(str(cand))
        return stub_fid(cand)


class DeferredStub(StubEvaluator):
    def __init__(self):
        super().__init__(True)

    def submit_cand_fid(self, cand=None, args=None):
        self.calls.append(str(cand))
        f = Future()
        self.unresolved.append((f, stub_fid(cand)))
        if len(self.unresolved) > 3:  # resolve late and out of step with submission
            g, v = self.unresolved.pop(0)
            g.set_result(v)
        return f

    def flush(self):
        for g, v in self.unresolved:
            g.set_result(v)
        self.unresolved = []


def build(cfg, evaluator, log):
    from autodiffusion_b200.search import EvolutionSearcher

    args = types.SimpleNamespace(**cfg, batch_size=4, num_samples=8, image_size=64)
    model = types.SimpleNamespace(layer_num=cfg["layer_num"])
    diffusion = types.SimpleNamespace(original_num_steps=cfg["original_num_steps"])
    return EvolutionSearcher(args, model, diffusion, cfg["time_step"], classifier=None, evaluator=evaluator, log=log)


def test_search_reproduces_the_reference_trace():
    trace = json.load(open(os.path.join(GOLDEN, "search_trace.json")))
    for name, g in trace.items():
        cfg = g["config"]
        lines = []
        ev = StubEvaluator(False)
        s = build(cfg, ev, lines.append)
        random.seed(cfg["seed"])
        np.random.seed(cfg["seed"])
        top = s.search()
        assert list(s.vis_dict.keys()) == g["visited"], name          # same individuals, same order
        assert [s.vis_dict[k]["fid"] for k in s.vis_dict] == g["fids"]
        assert top == g["top"] and s.skip_layer_range == g["skip_layer_range"] and s.epoch == g["epoch"]
        assert ev.calls == g["visited"]                                 # each individual evaluated exactly once
        assert lines[:12] == g["log_head"] and len(lines) == g["n_log"]  # the log users grep is line-identical


def test_deferred_fids_change_nothing_but_when_the_fid_lines_appear():
    g = json.load(open(os.path.join(GOLDEN, "search_trace.json")))["random_init"]
    cfg = g["config"]
    lines = []
    ev = DeferredStub()
    s = build(cfg, ev, lines.append)
    orig_join = s.join
    s.join = lambda: (ev.flush(), orig_join())[1]
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    top = s.search()
    assert list(s.vis_dict.keys()) == g["visited"] and top == g["top"]
    fid_lines = [l for l in lines if l.startswith("cand: ") and ", fid: " in l]
    assert [l.split(", fid: ")[0][len("cand: "):] for l in fid_lines] == g["visited"]  # emitted in submission order
    assert len(lines) == g["n_log"]


class PopulationStub(StubEvaluator):
    """An evaluator on more than one rank: whole candidates are dealt to ranks at join() (evaluate_population)."""
    world_size = 2

    def __init__(self):
        super().__init__(True)
        self.generations = []

    def evaluate_population(self, cands, args=None):
        self.generations.append(len(cands))
        self.calls += [str(c) for c in cands]
        return [stub_fid(c) for c in cands]

    def submit_cand_fid(self, cand=None, args=None):
        raise AssertionError("with population sharding nothing is sampled before join()")


def test_population_sharding_visits_the_same_individuals():
    """More than one rank: the search queues a generation's candidates and evaluates them as one sharded population;
    the individuals visited, the top-k and the log are those of the serial reference run."""
    g = json.load(open(os.path.join(GOLDEN, "search_trace.json")))["random_init"]
    cfg = g["config"]
    lines = []
    ev = PopulationStub()
    s = build(cfg, ev, lines.append)
    assert s.shard_population
    random.seed(cfg["seed"])
    np.random.seed(cfg["seed"])
    top = s.search()
    assert list(s.vis_dict.keys()) == g["visited"] and top == g["top"]
    assert ev.calls == g["visited"] and len(ev.generations) >= 2 and max(ev.generations) > 1
    fid_lines = [l for l in lines if l.startswith("cand: ") and ", fid: " in l]
    assert [l.split(", fid: ")[0][len("cand: "):] for l in fid_lines] == g["visited"]
    assert len(lines) == g["n_log"]


def test_save_and_resume(tmp_path):
    """Resume exactly as scripts/search_candidates.py does - `load_state(path)` then `search(path)` - from a state
    written inside the loop (after epoch e's selection and prune-range widening). The resumed run must not repeat
    that selection: same individuals, same top list and prune range as the uninterrupted golden run, no duplicate
    top-k entries, nothing evaluated twice."""
    g = json.load(open(os.path.join(GOLDEN, "search_trace.json")))["random_init"]
    cfg = dict(g["config"])
    for stop_after in (1, 4):
        path = str(tmp_path / f"state{stop_after}.pkl")
        # run 1: dies after `stop_after` epochs (the state file is the one written at that epoch's selection)
        s1 = build(dict(cfg, max_epochs=stop_after), StubEvaluator(False), lambda l: None)
        random.seed(cfg["seed"])
        np.random.seed(cfg["seed"])
        s1.search(state_path=path)
        # run 2: a fresh process constructs the searcher again and resumes
        ev2 = StubEvaluator(False)
        lines = []
        s2 = build(cfg, ev2, lines.append)
        random.seed(12345)  # whatever the new process seeded: load_state restores both generators
        s2.load_state(path)
        assert s2.epoch == stop_after - 1 and len(s2.vis_dict) == len(s1.vis_dict)
        top = s2.search(state_path=path)
        assert list(s2.vis_dict.keys()) == g["visited"] and top == g["top"]
        assert s2.skip_layer_range == g["skip_layer_range"] and s2.epoch == g["epoch"]
        assert len(set(top)) == len(top) and len(set(s2.keep_top_k[s2.select_num])) == len(s2.keep_top_k[s2.select_num])
        assert not set(ev2.calls) & set(s1.vis_dict.keys())
        assert not any(l.startswith("epoch = {}".format(stop_after - 1)) for l in lines)  # that epoch was already logged
