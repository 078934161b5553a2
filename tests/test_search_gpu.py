"""GPU: the search driver end to end on the CUDA path - small UNet + native depth-1 classifier, a 3-epoch
search of 2-step candidates; deferred host-side FIDs must equal the blocking ones bit for bit (same moments,
same float64 host arithmetic), and every visited individual must carry a finite FID."""
import math
import random
import types

import numpy as np
import pytest
import torch

from tests.test_classifier_gpu import build_classifier
from tests.util import SMALL_FLAGS, build_ours, oracle_weights

pytestmark = pytest.mark.gpu


def _run(defer):
    from autodiffusion_b200.evaluator import FIDStatistics
    from autodiffusion_b200.search import EvolutionSearcher

    cfg, sd = oracle_weights(SMALL_FLAGS)
    model, diffusion = build_ours(SMALL_FLAGS, sd)
    clf, _, _ = build_classifier(1, 64)
    d = 16
    proj = (torch.randn(3 * 64 * 64, d, generator=torch.Generator().manual_seed(7)) * (3.0 / (3 * 64 * 64) ** 0.5)).cuda()
    feature_fn = lambda u8: (u8.reshape(u8.shape[0], -1).float() / 255.0 - 0.5) @ proj
    rs = np.random.RandomState(11)
    a = rs.randn(d, d) / d ** 0.5
    ref = FIDStatistics(0.05 * rs.randn(d), a @ a.T * 0.05 + 0.02 * np.eye(d))
    args = types.SimpleNamespace(max_epochs=3, select_num=2, population_num=4, m_prob=0.3, crossover_num=1, mutation_num=2,
                                 max_prun=0.2, min_prun=0.0, batch_size=4, num_samples=8, image_size=64, class_cond=True,
                                 clip_denoised=True, classifier_scale=1.0, use_ddim=True, use_ddim_init_x=False, time_step=2,
                                 seed=0)
    lines = []
    s = EvolutionSearcher(args, model, diffusion, 2, classifier=clf, feature_fn=feature_fn, ref_stats=ref,
                          log=lines.append, defer_fid=defer)
    random.seed(5)
    np.random.seed(5)
    top = s.search()
    torch.cuda.synchronize()
    return s, top, lines


def test_search_runs_on_the_cuda_path_and_deferred_fids_match():
    s1, top1, lines1 = _run(defer=False)
    s2, top2, lines2 = _run(defer=True)
    assert list(s1.vis_dict) == list(s2.vis_dict) and len(s1.vis_dict) >= 8
    for k in s1.vis_dict:
        f1, f2 = s1.vis_dict[k]["fid"], s2.vis_dict[k]["fid"]
        assert math.isfinite(f1) and f1 == f2, (k, f1, f2)
    assert top1 == top2
    assert sum(l.startswith("No.1 ") for l in lines1) == 3  # one ranking per epoch, as users grep it
    print(f"search: {len(s1.vis_dict)} individuals, best fid {s1.vis_dict[top1[0]]['fid']:.4f}, "
          f"classifier+UNet kernels launched: {s1.model.gpu_launches}")
