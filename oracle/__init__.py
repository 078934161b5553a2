"""CPU oracle for the AutoDiffusion candidate-evaluator hot path.

TEST INFRASTRUCTURE ONLY. Nothing under `autodiffusion_b200/` imports this package; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker or as the timed CPU baseline — never as the
thing shipped.

What it is: a plain-PyTorch (fp32, CPU) functional restatement of the reference's algorithm
for the path SURVEY.md §8 scopes, each function citing the reference file:line it follows
(paths relative to /root/reference/examples/guided_diffusion/).

How it is pinned: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), so the pin is the reference itself, imported in the authoring
container: `tests/golden/make_golden.py` builds the reference modules
(`guided_diffusion.dynamic_unet.Dynamic_UNetModel`, `unet.EncoderUNetModel`,
`respace.SpacedDiffusion`, the search script's `reset_diffusion`/`model_fn`/`cond_fn` closures)
on weights from `oracle.weights`, checks this restatement against them, and commits the
reference's outputs as fixtures under `tests/golden/`. `tests/test_oracle_golden.py` re-checks
the oracle against those fixtures wherever the tests run (the GPU box has no /root/reference).
"""
