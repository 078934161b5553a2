"""CPU: host-side logic of the candidate evaluator — candidate resolution (sorted-rank skip indexing,
dedup), batch sharding, world-size-independent seeds, and the moment all-reduce + FID finalisation over
gloo with world size 2 (the N>1 path; the CUDA accumulation kernel itself is covered by -m gpu tests)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fid_ref
from tests.util import golden


def _diffusion():
    from autodiffusion_b200 import create_gaussian_diffusion

    return create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="cosine")


def test_resolve_candidate_sorted_rank_and_dedup():
    from autodiffusion_b200.sampler import resolve_candidate

    base = _diffusion()
    cand = {"timesteps": [690, 153, 926, 424], "skip_layers": [[], [9, 2], [], [17, 5, 12, 5]]}
    active, per_step = resolve_candidate(cand, base)
    assert active.timestep_map == [153, 424, 690, 926]
    # skip_layers[j] belongs to the j-th SMALLEST timestep (…progressive.py:394-396), not to timesteps[j]
    assert per_step == [[], [2, 9], [], [5, 12, 17]]
    g = golden("ddim_small.npz")
    seen_t = g["guided/seen_t"].tolist()
    assert [active.timestep_map[i] for i in range(active.num_timesteps)][::-1] == seen_t

    active, per_step = resolve_candidate({"timesteps": [5, 5, 900], "skip_layers": [[1], [2], [3]]}, base)
    assert active.timestep_map == [5, 900] and per_step == [[1], [2]]  # trailing entry unused after dedup
    active, per_step = resolve_candidate([94, 834, 217], base)  # timestep-only candidate (list form)
    assert active.num_timesteps == 3 and per_step == [[], [], []]
    with pytest.raises(IndexError):
        resolve_candidate({"timesteps": [1, 2, 3], "skip_layers": [[]]}, base)
    assert base.num_timesteps == 1000  # base untouched


def test_shards_cover_every_batch_once_and_seeds_ignore_world_size():
    from autodiffusion_b200.evaluator import batch_seed, shard_batches

    for nb in (1, 4, 7, 10):
        for world in (1, 2, 3, 8):
            got = sorted(b for r in range(world) for b in shard_batches(nb, r, world))
            assert got == list(range(nb))
    s = {batch_seed(0, "cand", b) for b in range(100)}
    assert len(s) == 100
    assert batch_seed(0, "cand", 3) != batch_seed(1, "cand", 3) != batch_seed(0, "cand2", 3)
    assert batch_seed(5, "c", 2) == batch_seed(5, "c", 2)


def _reduce_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from autodiffusion_b200.evaluator import FIDStatistics, MomentAccumulator

    g = golden("fid.npz")
    f1, f2 = g["full/f1"].astype(np.float64), g["full/f2"]
    mine = f1[rank::world]  # this rank's share of the candidate's features
    acc = MomentAccumulator(f1.shape[1], "cpu")
    acc.load_partial(mine.shape[0], mine.sum(0), mine.T @ mine)
    acc.all_reduce()
    mu, sigma = acc.statistics()
    fid = FIDStatistics(mu, sigma).frechet_distance(FIDStatistics(*fid_ref.compute_statistics(f2)))
    np.save(os.path.join(tmp, f"r{rank}.npy"), np.array([fid, float(acc.n.item())]))
    dist.destroy_process_group()


def test_moment_allreduce_world2_matches_reference_fid(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 500)
    mp.spawn(_reduce_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = golden("fid.npz")
    want = float(g["full/fid"])  # produced by the reference's FIDStatistics.frechet_distance
    for r in range(world):
        fid, n = np.load(tmp_path / f"r{r}.npy")
        assert n == g["full/f1"].shape[0]
        assert abs(fid - want) <= 1e-6 * max(1.0, abs(want)), (fid, want)


def test_moment_statistics_match_numpy_cov():
    from autodiffusion_b200.evaluator import MomentAccumulator

    rng = np.random.RandomState(1)
    f = rng.randn(300, 48) * 3 + 10  # large mean: checks the fp64 cancellation margin
    acc = MomentAccumulator(48, "cpu")
    acc.load_partial(300, f.sum(0), f.T @ f)
    mu, sigma = acc.statistics()
    np.testing.assert_allclose(mu, f.mean(0), rtol=1e-12)
    np.testing.assert_allclose(sigma, np.cov(f, rowvar=False), rtol=1e-9, atol=1e-10)
    acc.reset()
    with pytest.raises(ValueError):
        acc.statistics()


def test_fid_statistics_matches_reference_fixture():
    from autodiffusion_b200.evaluator import FIDStatistics

    g = golden("fid.npz")
    for name in ("full", "singular"):
        a = FIDStatistics(*fid_ref.compute_statistics(g[f"{name}/f1"]))
        b = FIDStatistics(*fid_ref.compute_statistics(g[f"{name}/f2"]))
        assert abs(a.frechet_distance(b) - float(g[f"{name}/fid"])) <= 1e-6 * float(g[f"{name}/fid"])
        # the evaluator's default (symmetric-eigenproblem form) against the same reference value
        assert abs(a.frechet_distance_eigh(b) - float(g[f"{name}/fid"])) <= 1e-6 * float(g[f"{name}/fid"])


def test_eigh_form_of_the_frechet_distance_matches_sqrtm():
    """The symmetric-eigenproblem form (fid_method="eigh") against the reference's sqrtm arithmetic."""
    from autodiffusion_b200.evaluator import FIDStatistics

    rs = np.random.RandomState(0)
    d = 96
    a, b = rs.randn(d, d) / d ** 0.5, rs.randn(d, d) / d ** 0.5
    f1 = FIDStatistics(0.1 * rs.randn(d), a @ a.T + 0.1 * np.eye(d))
    f2 = FIDStatistics(0.1 * rs.randn(d), b @ b.T + 0.1 * np.eye(d))
    x, y = f1.frechet_distance(f2), f1.frechet_distance_eigh(f2)
    assert abs(x - y) <= 1e-9 * abs(x)
    assert abs(x - fid_ref.frechet_distance(f1.mu, f1.sigma, f2.mu, f2.sigma)) <= 1e-9 * abs(x)


def _resolve_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from concurrent.futures import Future

    from autodiffusion_b200.evaluator import CandidateEvaluator, _RemoteFid

    ev = object.__new__(CandidateEvaluator)  # only the exchange logic: no model, no device
    ev.shard_fid, ev.world_size, ev.rank, ev.group = True, world, rank, None
    futs = []
    for seq in range(5):  # candidate `seq` is finished by rank seq % world (submit_cand_fid's rule)
        if seq % world == rank:
            f = Future()
            f.set_result(seq + 0.25)
            futs.append(f)
        else:
            futs.append(_RemoteFid(seq % world))
    np.save(os.path.join(tmp, f"v{rank}.npy"), np.array(ev.resolve(futs)))
    dist.destroy_process_group()


def test_sharded_host_fid_values_reach_every_rank(tmp_path):
    """world size 2, gloo: each rank computes the host FID of every other candidate; resolve() exchanges them."""
    world, port = 2, 30100 + (os.getpid() % 500)
    mp.spawn(_resolve_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"v{r}.npy").tolist() == [0.25, 1.25, 2.25, 3.25, 4.25]


def _disagree_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from concurrent.futures import Future

    from autodiffusion_b200.evaluator import CandidateEvaluator

    ev = object.__new__(CandidateEvaluator)
    ev.shard_fid, ev.world_size, ev.rank, ev.group = True, world, rank, None
    f = Future()
    f.set_result(1.5)  # BOTH ranks claim the candidate: the SUM would silently read 3.0
    try:
        ev.resolve([f])
        out = "no error"
    except RuntimeError as e:
        out = str(e)
    open(os.path.join(tmp, f"d{rank}.txt"), "w").write(out)
    dist.destroy_process_group()


def test_fid_exchange_detects_ownership_disagreement(tmp_path):
    """if ranks disagree on who finishes a candidate the SUM all-reduce returns 2x the FID or 0.0 (a
    'perfect' candidate). resolve() all-reduces an owner count beside the values and raises unless it is exactly 1."""
    world, port = 2, 30700 + (os.getpid() % 500)
    mp.spawn(_disagree_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert "instead of exactly one" in open(tmp_path / f"d{r}.txt").read()


def test_fid_owner_depends_on_the_candidate_only():
    from autodiffusion_b200.evaluator import fid_owner

    keys = [str({"timesteps": [i, 2 * i + 1], "skip_layers": [[], [i % 5]]}) for i in range(64)]
    for world in (1, 2, 8):
        owners = [fid_owner(k, world) for k in keys]
        assert all(0 <= o < world for o in owners)
        assert owners == [fid_owner(k, world) for k in keys]  # no hidden per-rank state (the old `_seq` counter)
        if world > 1:
            assert len(set(owners)) == world  # spread over ranks


def test_population_schedule_longest_first_with_a_batch_sharded_tail():
    """BASELINE configs[2]: 50 candidates x 4 batches on 8 ranks. Whole candidates longest-first, the 50 mod 8 tail split
    by batches: every (candidate, batch) exactly once, 25 batches on every rank (dealing i % world caps at 7 vs 6 = 89 %)."""
    from autodiffusion_b200.evaluator import schedule_population

    rs = np.random.RandomState(3)
    for n, nb, world in [(50, 4, 8), (8, 4, 4), (3, 4, 8), (7, 3, 2), (5, 1, 1), (13, 5, 4)]:
        costs = (1.0 + 0.1 * rs.rand(n)).tolist()
        whole, shared = schedule_population(costs, nb, world)
        assert schedule_population(costs, nb, world) == (whole, shared)  # deterministic: every rank computes the same
        seen = [(i, b) for r in whole for i in whole[r] for b in range(nb)]
        seen += [(i, b) for i, per in shared for bs in per.values() for b in bs]
        assert sorted(seen) == [(i, b) for i in range(n) for b in range(nb)]
        assert len(shared) == (n % world if world > 1 else 0)
        load = [sum(costs[i] for i in whole[r]) + sum(costs[i] / nb * len(per.get(r, [])) for i, per in shared)
                for r in range(world)]
        ideal = sum(costs) / world
        if n >= world:
            assert max(load) <= ideal * 1.12, (n, nb, world, load)
        for r in whole:  # longest first inside a rank
            assert [costs[i] for i in whole[r]] == sorted((costs[i] for i in whole[r]), reverse=True)
    whole, shared = schedule_population([1.0] * 50, 4, 8)
    assert [len(whole[r]) for r in range(8)] == [6] * 8 and len(shared) == 2
    assert all(sum(len(per.get(r, [])) for _, per in shared) == 1 for r in range(8))


def test_eigh_and_sqrtm_agree_when_samples_are_fewer_than_dimensions():
    """searches use 1000 samples at d = 2048, i.e. rank-deficient covariances; the opt-in eigh form must stay
    within the north star's +-0.1 of the reference's sqrtm arithmetic there (measured here: ~1e-6 relative)."""
    from autodiffusion_b200.evaluator import FIDStatistics

    rs = np.random.RandomState(5)
    d, n = 160, 60
    f1 = rs.randn(n, d) * (0.5 + rs.rand(d)) + 0.2
    f2 = rs.randn(4 * d, d) @ (np.eye(d) + 0.05 * rs.randn(d, d))
    a = FIDStatistics(*fid_ref.compute_statistics(f1))  # singular: rank <= n - 1
    b = FIDStatistics(*fid_ref.compute_statistics(f2))
    x, y = a.frechet_distance(b), a.frechet_distance_eigh(b)
    assert np.linalg.matrix_rank(a.sigma) < d
    assert abs(x - y) <= 1e-4 * abs(x) and abs(x - y) <= 0.1


def test_default_fid_method_is_the_reference_arithmetic():
    import inspect

    from autodiffusion_b200.evaluator import CandidateEvaluator

    assert inspect.signature(CandidateEvaluator.__init__).parameters["fid_method"].default == "sqrtm"


def test_remote_fid_placeholder_refuses_a_direct_result():
    from autodiffusion_b200.evaluator import _RemoteFid

    with pytest.raises(RuntimeError):
        _RemoteFid(1).result()


def test_inception_pool3_feature_extractor_contract():
    """The native extractor: torchvision's parameter names (pt_inception-2015-12-05 loads unchanged), deterministic for a
    seed, rejects anything that is not uint8 NHWC RGB, and has no CPU path."""
    from autodiffusion_b200.inception import InceptionPool3
    from oracle.inception_ref import InceptionPool3Ref

    m = InceptionPool3(seed=0)
    ref = InceptionPool3Ref(seed=0)
    want = {k: tuple(v.shape) for k, v in ref.net.state_dict().items()}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == want
    assert torch.equal(m.Mixed_6c.branch7x7dbl_3.conv.weight, InceptionPool3(seed=0).Mixed_6c.branch7x7dbl_3.conv.weight)
    assert not torch.equal(m.Conv2d_1a_3x3.conv.weight, InceptionPool3(seed=1).Conv2d_1a_3x3.conv.weight)
    u8 = torch.randint(0, 256, (2, 64, 64, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    with pytest.raises(ValueError):
        m(u8.float())
    with pytest.raises(ValueError):
        m(u8.permute(0, 3, 1, 2))
    with pytest.raises(RuntimeError):
        m(u8)  # CPU tensor: there is no fallback
    f = ref(u8)  # the oracle graph itself: uint8 NHWC in, fp32 [B, 2048] out
    assert f.shape == (2, 2048) and f.dtype == torch.float32 and torch.isfinite(f).all()


def test_inception_fid_variant_pooling():
    """The three pooling details that turn torchvision's Inception into the FID graph (pytorch-fid's FIDInceptionA/C/E):
    padding-excluding 3x3 average pools (a constant input stays constant up to the border) and a max pool in Mixed_7c.
    (Oracle-side check; the native kernels are compared with this graph in tests/test_inception_gpu.py.)"""
    from oracle.inception_ref import InceptionPool3Ref

    fid, tv = InceptionPool3Ref(seed=0).net, InceptionPool3Ref(seed=0, fid_variant=False).net
    with torch.no_grad():
        x = torch.ones(1, 192, 9, 9)
        pf, pt = fid.Mixed_5b._forward(x)[3], tv.Mixed_5b._forward(x)[3]
        assert (pf - pf[:, :, 4:5, 4:5]).abs().max() < 1e-5           # constant everywhere, border included
        assert (pt[:, :, 0, 0] - pt[:, :, 4, 4]).abs().max() > 1e-3   # torchvision divides the border sums by 9 as well
        assert torch.allclose(pf[:, :, 4, 4], pt[:, :, 4, 4], atol=1e-5)
        x7 = torch.randn(1, 2048, 5, 5, generator=torch.Generator().manual_seed(1))
        want = fid.Mixed_7c.branch_pool(torch.nn.functional.max_pool2d(x7, 3, 1, 1))
        assert torch.allclose(fid.Mixed_7c._forward(x7)[3], want) and not torch.allclose(tv.Mixed_7c._forward(x7)[3], want)
