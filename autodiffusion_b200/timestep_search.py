"""Evolutionary search over timestep subsequences only (full architecture), evaluated on the B200 path.

Drop-in for the `EvolutionSearcher` of the reference's two timestep-only drivers:

  * GD/search_imagenet64_classifier_guidance.py (:154-575) - class-conditional ADM-G with classifier guidance, optional
    `search_space` (a window of +-original_num_steps/100 around given core steps, built in its `__main__`, :645-668);
    BASELINE configs[0]'s caller. Its loop mutates / crosses after the LAST selection as well (no early break);
  * GD/search_uncondition_model.py (:150-575) - unconditional models (the LSUN-bedroom config, configs[3]), no
    classifier, optional explicit start individual `init_x` (half the first population are its mutations, m_prob 0.05),
    early break after the last selection.

An individual is `str(list_of_timesteps)`. Operators (same RNG consumption as the reference, so a seed yields the same
individuals; pinned by tests/golden/timestep_search_trace.json recorded from the unmodified scripts):
  random      `random.shuffle(space)` IN PLACE, first `time_step` entries (:265-275) - the shuffle persists in
              `search_space`, which changes the order of every later mutation's free list;
  crossover   per position, parent 1 or 2 with probability 1/2 (:395-411);
  mutation    per position with probability m_prob: a `random.choice` of the unused steps, without replacement (:432-464).
Evaluation, deferred FIDs, population sharding over ranks, save / resume: inherited from `search.EvolutionSearcher`.
"""
from __future__ import annotations

import ast
import random
from typing import Callable, List, Optional, Sequence

import numpy as np

from .respace import space_timesteps
from .search import EvolutionSearcher, _choice

__all__ = ["TimestepSearcher", "build_search_space"]


def build_search_space(core: Sequence[int], original_num_steps: int, init_x: Optional[Sequence[int]] = None) -> List[int]:
    """search_imagenet64_classifier_guidance.py:645-668: every step within R = original_num_steps / 100 below and R - 1
    above a core step (`range(max(s - R, 0), min(s + R, N))`), core = sorted(given) + optional DDIM initial steps."""
    core = sorted(core) + (list(init_x) if init_x is not None else [])
    r = int(original_num_steps / 100)
    space: List[int] = []
    for s in core:
        space += list(range(max(s - r, 0), min(s + r, original_num_steps)))
    return sorted(set(space))


class TimestepSearcher(EvolutionSearcher):
    def __init__(self, args, model, base_diffusion, time_step, classifier=None, search_space: Optional[List[int]] = None, *,
                 variant: str = "imagenet64", **kw):
        """`variant`: "imagenet64" (search_imagenet64_classifier_guidance.py) or "uncondition"
        (search_uncondition_model.py; `args.init_x` may hold a start individual as a string). Other keyword arguments
        as `search.EvolutionSearcher` (feature_fn / ref_stats or evaluator, log, defer_fid, shard_population)."""
        assert variant in ("imagenet64", "uncondition")
        if not hasattr(model, "layer_num"):
            model.layer_num = 0  # plain UNetModel checkpoints: the genome has no architecture part
        super().__init__(args, model, base_diffusion, time_step, classifier, **kw)
        self.variant = variant
        self.time_step = time_step
        self.search_space = search_space  # shuffled in place by sample_active_subnet, as the reference does
        self.x0 = getattr(args, "init_x", "") if variant == "uncondition" else ""

    # ---- operators ----
    def _space(self) -> List[int]:
        if self.search_space is not None:
            return self.search_space
        return list(range(self.base_diffusion.original_num_steps))

    def sample_active_subnet(self):
        space = self._space()
        random.shuffle(space)
        return space[:self.time_step]

    def _cross_pair(self, k):
        c1 = ast.literal_eval(_choice(self.keep_top_k[k]))
        c2 = ast.literal_eval(_choice(self.keep_top_k[k]))
        return [c1[i] if np.random.random_sample() < 0.5 else c2[i] for i in range(len(c1))]

    def _mutate(self, cand: list, m_prob: float, grow_empty: bool = False) -> list:
        free = [i for i in self._space() if i not in cand]
        for i in range(len(cand)):
            if np.random.random_sample() < m_prob:
                new_t = random.choice(free)
                free.remove(new_t)
                cand[i] = new_t
                if not free:
                    break
        return cand

    # ---- the search loop ----
    def search(self, state_path: Optional[str] = None):
        args = self.args
        self.log("population_num = {} select_num = {} mutation_num = {} crossover_num = {} random_num = {} max_epochs = {}".format(
            self.population_num, self.select_num, self.mutation_num, self.crossover_num,
            self.population_num - self.mutation_num - self.crossover_num, self.max_epochs))
        if self.epoch == 0 and not self.candidates and not self._selection_done:
            half = self.population_num // 2
            if self.x0 != "":  # search_uncondition_model.py:509-512
                self.get_random_before_search(half)
                self.candidates += self.mutate_init_x(x0=self.x0, mutation_num=self.population_num - half, m_prob=0.05)
            elif getattr(args, "use_ddim_init_x", False):
                steps = self.base_diffusion.original_num_steps
                respacing = ("ddim" if getattr(args, "use_ddim", True) else "") + str(args.time_step)
                init_x = str(list(space_timesteps(steps, respacing)))
                self._visit(init_x)
                self.candidates.append(init_x)
                # :390 asks for population // 2 + 1 individuals in total, the unconditional script (:524) for population // 2
                self.get_random_before_search(half + (1 if self.variant == "imagenet64" else 0))
                self.candidates += self.mutate_init_x(x0=init_x, mutation_num=self.population_num - half - 1, m_prob=0.1)
            else:
                self.get_random_before_search(self.population_num)
        while self.epoch < self.max_epochs:
            if not self._selection_done:
                self.log("epoch = {}".format(self.epoch))
                fid_of = lambda x: self.vis_dict[x]["fid"]
                self.update_top_k(self.candidates, k=self.select_num, key=fid_of)
                self.update_top_k(self.candidates, k=50, key=fid_of)
                self.log("epoch = {} : top {} result".format(self.epoch, len(self.keep_top_k[50])))
                for i, cand in enumerate(self.keep_top_k[50]):
                    self.log("No.{} {} fid = {}".format(i + 1, cand, self.vis_dict[cand]["fid"]))
                self._selection_done = True
                if state_path:
                    self.save_state(state_path)
            if self.variant == "uncondition" and self.epoch + 1 == self.max_epochs:
                break  # search_uncondition_model.py:553-554; the ImageNet script keeps going (:556-571)
            self.candidates = self.get_mutation(self.select_num, self.mutation_num, self.m_prob)
            self.candidates += self.get_cross(self.select_num, self.crossover_num)
            self.get_random(self.population_num)
            self.epoch += 1
            self._selection_done = False
        self.join()
        return self.keep_top_k[50]

    # persistence: the in-place shuffled search space is part of the state
    def save_state(self, path: str):
        super().save_state(path)
        import pickle

        with open(path, "rb") as f:
            st = pickle.load(f)
        st["search_space"] = self.search_space
        with open(path, "wb") as f:
            pickle.dump(st, f)

    def load_state(self, path: str):
        super().load_state(path)
        import pickle

        with open(path, "rb") as f:
            st = pickle.load(f)
        if st.get("search_space") is not None:
            self.search_space = st["search_space"]
