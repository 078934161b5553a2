"""CPU: the Stable-Diffusion search driver reproduces the reference script's operator trace
(tests/golden/sd_search_trace.json, recorded from the unmodified scripts/search_ea.py by
tests/golden/make_sd_search_golden.py with a stubbed FID): same individuals in the same order, same top list, and -
when scoring is not deferred - the same log, line for line."""
import json
import os
import random
import types
import zlib

import numpy as np
import pytest

from tests.util import GOLDEN


def stub_fid(cand) -> float:
    return (zlib.crc32(str(cand).encode()) % 100000) / 1000.0


def build(g, defer, lines, batches=None):
    from autodiffusion_b200.sd_search import EvolutionSearcher

    cfg = g["config"]

    def evaluate(cands):
        if batches is not None:
            batches.append(len(cands))
        return [stub_fid(c) for c in cands]

    opt = types.SimpleNamespace(**cfg)
    return EvolutionSearcher(opt, cfg["time_step"], evaluate, ddpm_num_timesteps=cfg["ddpm_num_timesteps"],
                             dpm_params=g["dpm_params"], log=lines.append, defer=defer)


@pytest.mark.parametrize("mode", ["random_init", "ddim_init", "dpm"])
def test_sd_search_reproduces_the_reference_trace(mode):
    g = json.load(open(os.path.join(GOLDEN, "sd_search_trace.json")))[mode]
    lines = []
    s = build(g, False, lines)
    random.seed(g["config"]["seed"])
    np.random.seed(g["config"]["seed"])
    top = s.search()
    assert list(s.vis_dict.keys()) == g["visited"]
    assert [s.vis_dict[k]["fid"] for k in s.vis_dict] == g["fids"]
    assert top == g["top"] and s.epoch == g["epoch"]
    assert lines == g["log"]


@pytest.mark.parametrize("mode", ["random_init", "dpm"])
def test_deferred_population_scoring_visits_the_same_individuals(mode):
    """defer=True: individuals are scored in batches (one per generation) - what a multi-GPU evaluator shards."""
    g = json.load(open(os.path.join(GOLDEN, "sd_search_trace.json")))[mode]
    lines, batches = [], []
    s = build(g, True, lines, batches)
    random.seed(g["config"]["seed"])
    np.random.seed(g["config"]["seed"])
    top = s.search()
    assert list(s.vis_dict.keys()) == g["visited"] and top == g["top"]
    assert sum(batches) == len(g["visited"]) and max(batches) >= g["config"]["population_num"] - 1
    fid_lines = [l for l in lines if l.startswith("cand: ") and ", fid: " in l]
    assert [l.split(", fid: ")[0][len("cand: "):] for l in fid_lines] == g["visited"]
    assert sorted(lines) == sorted(g["log"])  # the same lines; only the position of the fid lines moves


def _sd_eval_worker(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from autodiffusion_b200.sd_evaluator import SDCandidateEvaluator

    class Stub(SDCandidateEvaluator):  # the sharding / exchange logic only: no sampler, no device
        def __init__(self):
            self.rank, self.world_size, self.group = rank, world, None
            self.sampler = types.SimpleNamespace(model=types.SimpleNamespace(device="cpu"))
            self.mine = []

        def get_cand_fid(self, cand):
            self.mine.append(list(cand))
            return stub_fid(cand)

    ev = Stub()
    cands = [[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12], [5, 9, 13]]
    fids = ev.evaluate(cands)
    np.save(os.path.join(tmp, f"f{rank}.npy"), np.array(fids))
    json.dump(ev.mine, open(os.path.join(tmp, f"m{rank}.json"), "w"))
    dist.destroy_process_group()


def test_sd_population_is_dealt_to_ranks_world2(tmp_path):
    """gloo, world size 2: candidate i is sampled and scored by rank i % 2 only; every rank ends with all FIDs."""
    import torch.multiprocessing as mp

    world, port = 2, 30700 + (os.getpid() % 500)
    mp.spawn(_sd_eval_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    cands = [[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12], [5, 9, 13]]
    want = [stub_fid(c) for c in cands]
    for r in range(world):
        assert np.allclose(np.load(tmp_path / f"f{r}.npy"), want)
        assert json.load(open(tmp_path / f"m{r}.json")) == cands[r::world]
