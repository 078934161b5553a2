"""Evolutionary search over sampling time steps for the Stable-Diffusion path.

Drop-in for `EvolutionSearcher` of /root/reference/examples/"Stable Diffusion"/scripts/search_ea.py:184-633: the same
operators (random individuals, uniform-crossover, per-gene mutation against the unused time steps, the DDIM-initialised
population, top-k selection), the same consumption of Python's and numpy's global RNGs - so a seed visits the same
individuals in the same order as the reference script - and the same log lines. A candidate is a sorted list of
`time_step` integer DDPM steps (DDIM / PLMS) or of `time_step + 1` continuous times out of `dpm_params['full_timesteps']`
(DPM-Solver, search_ea.py:889-902); `vis_dict` is keyed by `str(sorted(cand))` as in the reference.

What differs is how candidates are scored: the reference samples and scores each individual serially inside `is_legal`
(:247-264). Here `evaluate(list_of_candidates) -> list_of_fids` is a callable that receives a whole batch of individuals:
with `defer=True` (default) an operator only registers its individuals and the batch is evaluated when their FIDs are
first needed (`join()`, before every top-k update) - which is what lets a multi-GPU evaluator deal whole candidates to
ranks. The operators never look at FID values between those points, so the individuals visited are unchanged.
"""
from __future__ import annotations

import copy
import random
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def _choice(x):
    """search_ea.py:47-48."""
    x = tuple(x)
    return x[np.random.randint(len(x))]


def _parse(cand: str):
    return eval(cand, {"np": np, "__builtins__": {}})  # keys may hold numpy scalars' reprs (init population), as the reference's do


def make_ddim_timesteps_uniform(num_ddim_timesteps: int, num_ddpm_timesteps: int) -> np.ndarray:
    """ldm/modules/diffusionmodules/util.py:46-61, 'uniform'."""
    c = round(num_ddpm_timesteps / num_ddim_timesteps)
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class EvolutionSearcher:
    def __init__(self, opt, time_step: int, evaluate: Callable[[List[list]], Sequence[float]], ddpm_num_timesteps: int = 1000,
                 dpm_params: Optional[dict] = None, log: Callable[[str], None] = print, defer: bool = True):
        """`opt` carries the reference's flags (search_ea.py:640-870): max_epochs, select_num, population_num, m_prob,
        crossover_num, mutation_num, use_ddim_init_x, dpm_solver. `evaluate` scores a list of candidates."""
        self.opt = opt
        self.time_step = time_step
        self.evaluate = evaluate
        self.ddpm_num_timesteps = ddpm_num_timesteps
        self.dpm_params = dpm_params
        self.log = log
        self.defer = defer
        self.max_epochs = opt.max_epochs
        self.select_num = opt.select_num
        self.population_num = opt.population_num
        self.m_prob = opt.m_prob
        self.crossover_num = opt.crossover_num
        self.mutation_num = opt.mutation_num
        self.use_ddim_init_x = opt.use_ddim_init_x
        self.dpm = bool(getattr(opt, "dpm_solver", False))
        if self.dpm and dpm_params is None:
            raise ValueError("dpm_solver search needs dpm_params {'full_timesteps', 'init_timesteps'} (search_ea.py:889-902)")
        self.keep_top_k: Dict[int, List[str]] = {self.select_num: [], 50: []}
        self.epoch = 0
        self.candidates: List[str] = []
        self.vis_dict: Dict[str, dict] = {}
        self._queued: Dict[str, list] = {}

    # ---- scoring ----
    def _visit(self, cand: str) -> bool:
        """is_legal / is_legal_before_search (:231-264; the two are the same function)."""
        key = str(sorted(_parse(cand)))
        info = self.vis_dict.setdefault(key, {})
        if "visited" in info:
            self.log("cand: {} has visited!".format(key))
            return False
        parsed = _parse(key)
        if self.defer:
            self._queued[key] = parsed
        else:
            info["fid"] = float(self.evaluate([parsed])[0])
            self.log("cand: {}, fid: {}".format(key, info["fid"]))
        info["visited"] = True
        return True

    is_legal = _visit
    is_legal_before_search = _visit

    def join(self):
        """Score every registered individual (in registration order) and emit its log line."""
        if not self._queued:
            return
        fids = self.evaluate(list(self._queued.values()))
        for key, fid in zip(list(self._queued), fids):
            self.vis_dict[key]["fid"] = float(fid)
            self.log("cand: {}, fid: {}".format(key, self.vis_dict[key]["fid"]))
        self._queued.clear()

    def update_top_k(self, candidates, *, k, key, reverse=False):
        assert k in self.keep_top_k
        self.join()
        self.log("select ......")
        t = self.keep_top_k[k]
        t += candidates
        t.sort(key=key, reverse=reverse)
        self.keep_top_k[k] = t[:k]

    # ---- individuals ----
    def sample_active_subnet(self):
        """:489-495."""
        use_timestep = [i for i in range(self.ddpm_num_timesteps)]
        random.shuffle(use_timestep)
        return use_timestep[:self.time_step]

    def sample_active_subnet_dpm(self):
        """:497-502."""
        use_timestep = copy.deepcopy(self.dpm_params["full_timesteps"])
        random.shuffle(use_timestep)
        return use_timestep[:self.time_step + 1]

    def _fill_random(self):
        """get_random_before_search / get_random (:266-294)."""
        num = self._fill_target
        self.log("random select ........")
        while len(self.candidates) < num:
            cand = self.sample_active_subnet_dpm() if self.dpm else self.sample_active_subnet()
            cand = str(sorted(cand))
            if not self._visit(cand):
                continue
            self.candidates.append(cand)
            self.log("random {}/{}".format(len(self.candidates), num))
        self.log("random_num = {}".format(len(self.candidates)))

    def get_random(self, num):
        self._fill_target = num
        self._fill_random()

    get_random_before_search = get_random

    def get_cross(self, k, cross_num):
        """:296-329."""
        assert k in self.keep_top_k
        self.log("cross ......")
        res = []
        max_iters = cross_num * 10
        while len(res) < cross_num and max_iters > 0:
            max_iters -= 1
            cand1 = _parse(_choice(self.keep_top_k[k]))
            cand2 = _parse(_choice(self.keep_top_k[k]))
            new_cand = [cand1[i] if np.random.random_sample() < 0.5 else cand2[i] for i in range(len(cand1))]
            cand = str(sorted(new_cand))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("cross {}/{}".format(len(res), cross_num))
        self.log("cross_num = {}".format(len(res)))
        return res

    def _mutate(self, cand: list, m_prob: float) -> list:
        """random_func of get_mutation(_dpm) / mutate_init_x(_dpm) (:338-356, 378-396, 417-435, 456-474)."""
        pool = self.dpm_params["full_timesteps"] if self.dpm else range(self.ddpm_num_timesteps)
        candidates = [i for i in pool if i not in cand]
        for i in range(len(cand)):
            if np.random.random_sample() < m_prob:
                new_c = random.choice(candidates)
                del candidates[candidates.index(new_c)]
                cand[i] = new_c
                if len(candidates) == 0:
                    break
        return cand

    def get_mutation(self, k, mutation_num, m_prob):
        """:331-369 / 371-409."""
        assert k in self.keep_top_k
        self.log("mutation ......")
        res = []
        max_iters = mutation_num * 10
        while len(res) < mutation_num and max_iters > 0:
            max_iters -= 1
            cand = self._mutate(_parse(_choice(self.keep_top_k[k])), m_prob)
            cand = str(sorted(cand))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("mutation {}/{}".format(len(res), mutation_num))
        self.log("mutation_num = {}".format(len(res)))
        return res

    get_mutation_dpm = get_mutation

    def mutate_init_x(self, x0, mutation_num, m_prob):
        """:411-448 / 450-487."""
        self.log("mutation x0 ......")
        res = []
        max_iters = mutation_num * 10
        while len(res) < mutation_num and max_iters > 0:
            max_iters -= 1
            cand = self._mutate(_parse(x0), m_prob)
            cand = str(sorted(cand))
            if not self._visit(cand):
                continue
            res.append(cand)
            self.log("mutation x0 {}/{}".format(len(res), mutation_num))
        self.log("mutation_num = {}".format(len(res)))
        return res

    mutate_init_x_dpm = mutate_init_x

    # ---- the loop (:568-633) ----
    def search(self):
        self.log("population_num = {} select_num = {} mutation_num = {} crossover_num = {} random_num = {} max_epochs = {}".format(
            self.population_num, self.select_num, self.mutation_num, self.crossover_num,
            self.population_num - self.mutation_num - self.crossover_num, self.max_epochs))
        if self.use_ddim_init_x is False:
            self.get_random_before_search(self.population_num)
        else:
            if self.dpm:
                init_x = self.dpm_params["init_timesteps"]
            else:
                init_x = make_ddim_timesteps_uniform(self.time_step, self.ddpm_num_timesteps)
            init_x = sorted(list(init_x))
            self.is_legal_before_search(str(init_x))
            self.candidates.append(str(init_x))
            self.get_random_before_search(self.population_num // 2)
            self.candidates += self.mutate_init_x(x0=str(init_x), mutation_num=self.population_num - self.population_num // 2 - 1,
                                                  m_prob=0.1)
        while self.epoch < self.max_epochs:
            self.log("epoch = {}".format(self.epoch))
            fid_of = lambda x: self.vis_dict[x]["fid"]
            self.update_top_k(self.candidates, k=self.select_num, key=fid_of)
            self.update_top_k(self.candidates, k=50, key=fid_of)
            self.log("epoch = {} : top {} result".format(self.epoch, len(self.keep_top_k[50])))
            for i, cand in enumerate(self.keep_top_k[50]):
                self.log("No.{} {} fid = {}".format(i + 1, cand, self.vis_dict[cand]["fid"]))
            if self.epoch + 1 == self.max_epochs:
                break
            mutation = self.get_mutation(self.select_num, self.mutation_num, self.m_prob)
            self.candidates = mutation
            self.candidates += self.get_cross(self.select_num, self.crossover_num)
            self.get_random(self.population_num)
            self.epoch += 1
        return self.keep_top_k[50]
