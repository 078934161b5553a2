"""GPU: population evaluation (BASELINE configs[2]) on one rank - candidates drawn by the search's own
`sample_active_subnet`, launch plans of candidate i+1 recorded and captured on a side stream while candidate i samples
(no device synchronisation, no validation run), FIDs deferred to the host worker. The pipelined path must give exactly
the FIDs of evaluating the same candidates one by one (`get_cand_fid`): same seeds -> same images."""
import copy

import numpy as np
import pytest
import torch

from oracle import fid_ref
from tests.util import SMALL_FLAGS, build_ours, oracle_weights

pytestmark = pytest.mark.gpu


def test_evaluate_population_matches_one_by_one_evaluation():
    from autodiffusion_b200.classifier import ClassifierGuidance
    from autodiffusion_b200.evaluator import CandidateEvaluator, FIDStatistics
    from autodiffusion_b200.search import draw_population
    from tests.test_classifier_gpu import build_classifier

    cfg, sd = oracle_weights(SMALL_FLAGS)
    model, diffusion = build_ours(SMALL_FLAGS, sd)
    clf, _, _ = build_classifier(1, 64)
    proj = torch.randn(3 * 64 * 64, 48, generator=torch.Generator().manual_seed(13)).cuda() / 255.0
    feature_fn = lambda u8: u8.reshape(u8.shape[0], -1).float() @ proj
    rng = np.random.RandomState(0)
    ref_stats = FIDStatistics(*fid_ref.compute_statistics(rng.randn(200, 48) * 2 + 5))
    pop = draw_population(5, 4, model.layer_num, 0.2, seed=3)
    assert len({str(c) for c in pop}) == 5 and all(len(c["timesteps"]) >= 4 for c in pop)
    assert any(len(s) > 0 for c in pop for s in c["skip_layers"])  # the prune range is open: masks are not all empty

    kw = dict(batch_size=4, num_samples=10, image_size=64, seed=5, cond_fn=ClassifierGuidance(clf, 1.0), max_cached_plans=2)
    ev = CandidateEvaluator(model, diffusion, feature_fn, ref_stats, **kw)
    fids = ev.evaluate_population(pop)
    info = ev.last_population
    assert info["whole_per_rank"] == [5] and info["shared"] == 0
    print(f"population of 5 on one rank: fids {np.round(fids, 4).tolist()}; first plan {info['plan_build_first_s']:.3f} s, "
          f"the other four built under sampling in {info['plan_build_overlapped_s']:.3f} s")
    # one by one, on a fresh evaluator and a fresh model (no shared plan cache)
    model2, diffusion2 = build_ours(SMALL_FLAGS, sd)
    ev2 = CandidateEvaluator(model2, diffusion2, feature_fn, ref_stats, **dict(kw, cond_fn=ClassifierGuidance(clf, 1.0)))
    want = [ev2.get_cand_fid(c) for c in pop]
    assert np.allclose(fids, want, rtol=1e-6, atol=1e-6), (fids, want)
    # the plan cache of the model stays bounded however many skip sets a population brings
    model.max_cached_plans = 4
    ev.evaluate_population(draw_population(4, 4, model.layer_num, 0.2, seed=9))
    assert len(model._plans) <= 5
