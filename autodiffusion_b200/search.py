"""Evolutionary search over (timestep subsequence, per-step block-skip lists), evaluated on the B200 path.

Drop-in for `EvolutionSearcher` of
GD/search_dynamic_unet_imagenet64_classifier_guidance_progressive.py (:155-715): the same constructor
arguments, the same individuals (`str(dict)` keys of `vis_dict`, :343-363), the same operators — random
individuals under the index budget (`sample_active_subnet`, :284-338), per-position crossover
(:472-520), timestep / skip-list mutation (:522-590), the DDIM-initialised population
(:650-669), top-k bookkeeping and the progressive widening of the prune range (:684-697) — and the
same log lines (`cand: {…}, fid: X`, `No.i {…} fid = X`, …) that users grep.

What is different underneath:
  * every individual is evaluated by `evaluator.CandidateEvaluator`: one CUDA graph per candidate
    (UNet + native classifier guidance + DDIM updates), batches sharded over the ranks of the box, one
    all-reduce of the FID moments - instead of a Python sampling loop and an image all_gather;
  * the host-side part of an evaluation (mu / Sigma / sqrtm, the reference's `fid_time`) is deferred to a
    worker thread and overlaps the sampling of the following individuals. Operators only ever need the
    *legality* of a new individual (has it been visited?), never its FID, until the generation is
    complete; FIDs are joined before `update_top_k`. Log lines are emitted in the reference's order;
  * no `pdb.set_trace()` (the reference drops into the debugger when its rejection loops run long,
    :306-308, 331-333): a `RuntimeError` is raised instead;
  * `save_state` / `load_state`: `vis_dict`, top-k lists, epoch, prune range and both RNG states, so an
    interrupted search resumes without re-evaluating visited individuals.
All ranks run the same driver with the same seeds (the reference seeds `random` / `numpy.random`
globally, :762-765), so they draw the same individuals without communication.

The reference's operator quirks are kept, because they change which individuals a seed produces:
skip-list "mutation" of a non-empty list compares instead of assigning (`cand[...][i][j] == new_c`,
:575, :628) and is therefore a no-op that still consumes random numbers; the crossover tail uses
Python's lexicographic list comparison (:499-505).
"""
from __future__ import annotations

import ast
import copy
import pickle
import random
from concurrent.futures import Future
from typing import Callable, Dict, List, Optional

import numpy as np

from .classifier import ClassifierGuidance, EncoderUNetModel
from .evaluator import CandidateEvaluator, FIDStatistics
from .respace import space_timesteps

__all__ = ["EvolutionSearcher", "sample_active_subnet", "draw_population"]

_MAX_REJECTIONS = int(1e6)


def _choice(seq):
    """The reference's module-level `choice` (:46-47): numpy's global RNG, not `random.choice`."""
    seq = tuple(seq)
    return seq[np.random.randint(len(seq))]


def sample_active_subnet(n_steps: int, L: int, max_index_number: int, skip_layer_range):
    """A random individual under the (step, block) index budget (…progressive.py:284-338): timesteps from a shuffled
    order of the `n_steps` base steps; per step a skip list of int(u * L) randomly chosen block ids, u uniform in
    `skip_layer_range`; steps are added while the budget `max_index_number` of executed (step, block) slots allows.
    Draws from the global `random` / `numpy.random` generators in the reference's order."""
    order = list(range(n_steps))
    random.shuffle(order)
    lo, hi = skip_layer_range
    used, t_idx = 0, 0
    skip_lists, timesteps = [], []
    for _ in range(100000):
        n_skip = -10000
        tries = 0
        while used + L - n_skip > max_index_number:
            tries += 1
            n_skip = int((np.random.random_sample() * (hi - lo) + lo) * L)
            if tries > _MAX_REJECTIONS:
                raise RuntimeError("sample_active_subnet: no skip count fits the remaining index budget "
                                   f"(used {used} of {max_index_number}, range {list(skip_layer_range)})")
        layers = list(range(L))
        random.shuffle(layers)
        skip_lists.append(layers[:n_skip])
        timesteps.append(order[t_idx])
        t_idx += 1
        used += L - n_skip
        room = used + L - int(L * hi)
        if room > max_index_number:
            break
        if room == max_index_number:
            layers = list(range(L))
            random.shuffle(layers)
            skip_lists.append(layers[:int(L * hi)])
            timesteps.append(order[t_idx])
            break
    else:
        raise RuntimeError("sample_active_subnet did not terminate")
    return {"timesteps": timesteps, "skip_layers": skip_lists}


def draw_population(n: int, time_step: int, layer_num: int, max_prun: float, seed: int = 0, n_steps: int = 1000):
    """`n` random individuals as the search draws its population once the prune range is fully open
    (`skip_layer_range = [0, max_prun]`, budget `time_step * layer_num`): BASELINE configs[2]'s workload. The global
    generators are seeded for the draw and restored afterwards."""
    st_py, st_np = random.getstate(), np.random.get_state()
    random.seed(seed)
    np.random.seed(seed)
    try:
        return [sample_active_subnet(n_steps, layer_num, time_step * layer_num, [0, max_prun]) for _ in range(n)]
    finally:
        random.setstate(st_py)
        np.random.set_state(st_np)


class EvolutionSearcher:
    def __init__(self, args, model, base_diffusion, time_step, classifier=None, index_step=None, *,
                 feature_fn: Optional[Callable] = None, ref_stats: Optional[FIDStatistics] = None,
                 evaluator: Optional[CandidateEvaluator] = None, log: Callable[[str], None] = print,
                 defer_fid: bool = True, shard_population: Optional[bool] = None):
        """`args` carries the reference's flags (:722-752): max_epochs, select_num, population_num, m_prob,
        crossover_num, mutation_num, max_prun, min_prun, batch_size, num_samples, image_size, class_cond,
        clip_denoised, classifier_scale, use_ddim, use_ddim_init_x, time_step.
        `classifier`: an `EncoderUNetModel` (guidance runs on the CUDA path), any cond_fn-style callable, or None.
        `feature_fn` / `ref_stats` replace the reference's TensorFlow Inception evaluator and pickled reference
        statistics (:190-203); or pass a ready `evaluator` (anything with submit_cand_fid / get_cand_fid)."""
        self.args = args
        self.model = model
        self.base_diffusion = base_diffusion
        self.classifier = classifier
        self.active_diffusion = copy.deepcopy(base_diffusion)
        self.init_time_step = time_step
        self.model_layers = model.layer_num
        self.max_index_number = time_step * self.model_layers
        if index_step is not None:
            self.max_index_number = int(ast.literal_eval(str(index_step)))
        self.max_epochs = args.max_epochs
        self.select_num = args.select_num
        self.population_num = args.population_num
        self.m_prob = args.m_prob
        self.crossover_num = args.crossover_num
        self.mutation_num = args.mutation_num
        self.keep_top_k: Dict[int, List[str]] = {self.select_num: [], 50: []}
        self.epoch = 0
        self.candidates: List[str] = []
        self.vis_dict: Dict[str, dict] = {}
        self.max_fid = getattr(args, "max_fid", 48.0)
        self.max_prun = getattr(args, "max_prun", 0.0)
        self.min_prun = getattr(args, "min_prun", 0.0)
        self.skip_layer_range = [0, 0]
        self.last_best_cand = None
        self.log = log
        self.defer_fid = defer_fid
        # more than one rank: whole candidates (not batches of one candidate) are dealt to ranks, SURVEY §8(e)
        self.shard_population = shard_population
        self._pending: Dict[str, Future] = {}
        self._queued: Dict[str, object] = {}
        self._selection_done = False  # selection / widening of `self.epoch` already applied (see load_state)
        if not getattr(args, "use_ddim", True):
            raise NotImplementedError("the evaluator path covers DDIM sampling (use_ddim=True), as every search script sets")
        if evaluator is None:
            if feature_fn is None or ref_stats is None:
                raise ValueError("EvolutionSearcher needs feature_fn and ref_stats (or an evaluator)")
            if isinstance(classifier, EncoderUNetModel):
                cond_fn = ClassifierGuidance(classifier, getattr(args, "classifier_scale", 1.0))
            else:
                cond_fn = classifier  # a cond_fn-style callable or None
            evaluator = CandidateEvaluator(
                model, base_diffusion, feature_fn, ref_stats, batch_size=args.batch_size, num_samples=args.num_samples,
                image_size=args.image_size, class_cond=getattr(args, "class_cond", True),
                clip_denoised=getattr(args, "clip_denoised", True), cond_fn=cond_fn, seed=getattr(args, "seed", 0))
        self.evaluator = evaluator
        if self.shard_population is None:
            self.shard_population = getattr(evaluator, "world_size", 1) > 1

    # ---- genome <-> flat index list (kept for parity with :206-217; used by predictor-based variants) ----
    def cand2gen(self, cand):
        ret = []
        for i, t in enumerate(cand["timesteps"]):
            kept = [k for k in range(self.model_layers) if k not in cand["skip_layers"][i]]
            ret += [k + self.model_layers * t for k in kept]
        if len(ret) < self.max_index_number:
            ret += [0] * (self.max_index_number - len(ret))
        return ret

    # ---- evaluation ----
    def get_cand_fid(self, cand=None, args=None) -> float:
        return self.evaluator.get_cand_fid(cand=cand, args=args)

    def _visit(self, cand: str) -> bool:
        """:355-367 (is_legal / is_legal_before_search are the same function in the reference)."""
        info = self.vis_dict.setdefault(cand, {})
        if "visited" in info:
            self.log("cand: {} has visited!".format(cand))
            return False
        parsed = ast.literal_eval(cand)
        if self.defer_fid and self.shard_population and callable(getattr(self.evaluator, "evaluate_population", None)):
            self._queued[cand] = parsed  # sampled at join(): the generation's candidates are dealt to the ranks
        elif self.defer_fid and callable(getattr(self.evaluator, "submit_cand_fid", None)):
            self._pending[cand] = self.evaluator.submit_cand_fid(cand=parsed, args=self.args)
        else:
            info["fid"] = self.evaluator.get_cand_fid(cand=parsed, args=self.args)
            self.log("cand: {}, fid: {}".format(cand, info["fid"]))
        info["visited"] = True
        return True

    is_legal = _visit
    is_legal_before_search = _visit

    def join(self):
        """Resolve every deferred FID (in submission order) and emit its log line."""
        if self._queued:
            fids = self.evaluator.evaluate_population(list(self._queued.values()), self.args)
            for cand, fid in zip(list(self._queued), fids):
                self.vis_dict[cand]["fid"] = fid
                self.log("cand: {}, fid: {}".format(cand, fid))
            self._queued.clear()
        resolve = getattr(self.evaluator, "resolve", None)
        fids = resolve(list(self._pending.values())) if callable(resolve) else [f.result() for f in self._pending.values()]
        for (cand, fut), fid in zip(list(self._pending.items()), fids):
            self.vis_dict[cand]["fid"] = fid
            self.log("cand: {}, fid: {}".format(cand, self.vis_dict[cand]["fid"]))
            del self._pending[cand]

    def update_top_k(self, candidates, *, k, key, reverse=False):
        assert k in self.keep_top_k
        self.join()
        self.log("select ......")
        t = self.keep_top_k[k]
        have = set(t)  # guard: an uninterrupted run never re-adds an individual (every new one passed is_legal)
        t += [c for c in candidates if not (c in have or have.add(c))]
        t.sort(key=key, reverse=reverse)
        self.keep_top_k[k] = t[:k]

    # ---- individuals ----
    def sample_active_subnet(self):
        """A random individual under the (step, block) index budget `max_index_number` (:284-338)."""
        return sample_active_subnet(self.base_diffusion.original_num_steps, self.model_layers, self.max_index_number,
                                    self.skip_layer_range)

    def _fill(self, num, tag):
        self.log("random select ........")
        while len(self.candidates) < num:
            cand = str(self.sample_active_subnet())
            if not self._visit(cand):
                continue
            self.candidates.append(cand)
            self.log("random {}/{}".format(len(self.candidates), num))
        self.log("random_num = {}".format(len(self.candidates)))

    def get_random(self, num):
        self._fill(num, "random")

    def get_random_before_search(self, num):
        self._fill(num, "random")

    def _cross_pair(self, k):
        """:477-507."""
        c1 = ast.literal_eval(_choice(self.keep_top_k[k]))
        c2 = ast.literal_eval(_choice(self.keep_top_k[k]))
        new = {"timesteps": [], "skip_layers": []}
        for i in range(min(len(c1["timesteps"]), len(c2["timesteps"]))):
            src = c1 if np.random.random_sample() < 0.5 else c2
            new["timesteps"].append(src["timesteps"][i])
            new["skip_layers"].append(src["skip_layers"][i])
        for parent in (c1, c2):  # lexicographic list comparison, as the reference writes it
            if new["timesteps"] < parent["timesteps"]:
                new["timesteps"] += parent["timesteps"][len(new["timesteps"]):]
                new["skip_layers"] += parent["skip_layers"][len(new["skip_layers"]):]
        return new

    def get_cross(self, k, cross_num):
        assert k in self.keep_top_k
        self.log("cross ......")
        res = []
        max_iters = cross_num * 10
        while len(res) < cross_num and max_iters > 0:
            max_iters -= 1
            cand = str(self._cross_pair(k))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("cross {}/{}".format(len(res), cross_num))
        self.log("cross_num = {}".format(len(res)))
        return res

    def _mutate(self, cand: dict, m_prob: float, grow_empty: bool) -> dict:
        """:534-583 (`grow_empty`) and :601-634 (mutate_init_x: empty skip lists stay empty)."""
        n_steps = self.base_diffusion.original_num_steps
        free = [i for i in range(n_steps) if i not in cand["timesteps"]]
        for i in range(len(cand["timesteps"])):
            if np.random.random_sample() < m_prob:
                new_t = random.choice(free)
                free.remove(new_t)
                cand["timesteps"][i] = new_t
                if not free:
                    break
        lo, hi = self.skip_layer_range
        if hi == 0:
            return cand
        L = self.model_layers
        for i in range(len(cand["skip_layers"])):
            free = [j for j in range(L) if j not in cand["skip_layers"][i]]
            if len(cand["skip_layers"][i]) == 0:
                if grow_empty and np.random.random_sample() < m_prob:
                    layers = list(range(L))
                    n_skip = int((np.random.random_sample() * (hi - lo) + lo) * L)
                    random.shuffle(layers)
                    cand["skip_layers"][i] = layers[:n_skip]
            else:
                for _j in range(len(cand["skip_layers"][i])):
                    if np.random.random_sample() < m_prob:
                        new_c = random.choice(free)
                        free.remove(new_c)
                        # the reference compares here (`== new_c`) instead of assigning: the list is unchanged
                        if not free:
                            break
        return cand

    def get_mutation(self, k, mutation_num, m_prob):
        assert k in self.keep_top_k
        self.log("mutation ......")
        res = []
        max_iters = mutation_num * 10
        while len(res) < mutation_num and max_iters > 0:
            max_iters -= 1
            cand = str(self._mutate(ast.literal_eval(_choice(self.keep_top_k[k])), m_prob, grow_empty=True))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("mutation {}/{}".format(len(res), mutation_num))
        self.log("mutation_num = {}".format(len(res)))
        return res

    def mutate_init_x(self, x0, mutation_num, m_prob):
        self.log("mutation x0 ......")
        res = []
        max_iters = mutation_num * 10
        while len(res) < mutation_num and max_iters > 0:
            max_iters -= 1
            cand = str(self._mutate(ast.literal_eval(x0), m_prob, grow_empty=False))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("mutation x0 {}/{}".format(len(res), mutation_num))
        self.log("mutation_num = {}".format(len(res)))
        return res

    # ---- persistence ----
    def save_state(self, path: str):
        self.join()
        with open(path, "wb") as f:
            pickle.dump(dict(vis_dict=self.vis_dict, keep_top_k=self.keep_top_k, epoch=self.epoch,
                             candidates=self.candidates, skip_layer_range=self.skip_layer_range,
                             last_best_cand=self.last_best_cand, py_random=random.getstate(),
                             np_random=np.random.get_state(), selection_done=self._selection_done), f)

    def load_state(self, path: str):
        with open(path, "rb") as f:
            st = pickle.load(f)
        self.vis_dict, self.keep_top_k, self.epoch = st["vis_dict"], st["keep_top_k"], st["epoch"]
        self.candidates, self.skip_layer_range = st["candidates"], st["skip_layer_range"]
        self.last_best_cand = st["last_best_cand"]
        # written from inside the loop, i.e. AFTER selection / prune-range widening of `epoch`: `search()` must not
        # repeat them for that epoch (it would re-add the generation to the top-k lists and widen the range twice)
        self._selection_done = bool(st.get("selection_done", False))
        random.setstate(st["py_random"])
        np.random.set_state(st["np_random"])

    # ---- the search loop (:636-715) ----
    def search(self, state_path: Optional[str] = None):
        args = self.args
        self.log("population_num = {} select_num = {} mutation_num = {} crossover_num = {} random_num = {} max_epochs = {}".format(
            self.population_num, self.select_num, self.mutation_num, self.crossover_num,
            self.population_num - self.mutation_num - self.crossover_num, self.max_epochs))
        if self.epoch == 0 and not self.candidates:
            if not getattr(args, "use_ddim_init_x", False):
                self.get_random_before_search(self.population_num)
            else:
                steps = self.base_diffusion.original_num_steps
                init_x = list(space_timesteps(steps, "ddim" + str(args.time_step)))
                init_cand = str({"timesteps": init_x, "skip_layers": [[]] * len(init_x)})
                self._visit(init_cand)
                self.candidates.append(init_cand)
                self.get_random_before_search(self.population_num // 2 + 1)
                self.candidates += self.mutate_init_x(x0=init_cand, m_prob=0.1,
                                                      mutation_num=self.population_num - self.population_num // 2 - 1)
        while self.epoch < self.max_epochs:
            if not self._selection_done:  # a state loaded from `state_path` resumes right after this block
                self.log("epoch = {}".format(self.epoch))
                fid_of = lambda x: self.vis_dict[x]["fid"]
                self.update_top_k(self.candidates, k=self.select_num, key=fid_of)
                self.update_top_k(self.candidates, k=50, key=fid_of)
                self.log("epoch = {} : top {} result".format(self.epoch, len(self.keep_top_k[50])))
                for i, cand in enumerate(self.keep_top_k[50]):
                    self.log("No.{} {} fid = {}".format(i + 1, cand, self.vis_dict[cand]["fid"]))
                # progressive widening of the prune range (:684-693)
                if self.skip_layer_range[1] == 0 and (self.last_best_cand == self.keep_top_k[50][0] or self.epoch > 4):
                    self.skip_layer_range[1] = self.max_prun / 5
                elif 0 < self.skip_layer_range[1] < self.max_prun:
                    self.skip_layer_range[1] += self.max_prun / 5
                if self.skip_layer_range[0] == 0 and self.epoch > 5:
                    self.skip_layer_range[0] = self.min_prun
                self.last_best_cand = self.keep_top_k[50][0]
                self.log("skip_layer_range_left = {} , skip_layer_range_right {}".format(*self.skip_layer_range))
                self._selection_done = True
                if state_path:
                    self.save_state(state_path)
            if self.epoch + 1 == self.max_epochs:
                break
            mutation = self.get_mutation(self.select_num, self.mutation_num, self.m_prob)
            self.candidates = mutation
            self.candidates += self.get_cross(self.select_num, self.crossover_num)
            self.get_random(self.population_num)
            self.epoch += 1
            self._selection_done = False
        self.join()
        return self.keep_top_k[50]
