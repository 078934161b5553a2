"""Fast-path recognition: the reference's stock sampling call routed to the fused whole-candidate plan.

SURVEY.md §8(b) row 3. A user who swaps the import and runs the reference's `get_cand_fid` body unchanged
(search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-420) calls

    diffusion.ddim_sample_loop(model_fn, shape, clip_denoised=..., model_kwargs={"y": classes, "skip_layers": ...},
                               cond_fn=cond_fn, device=...)

with two opaque Python closures. The generic loop must call them once per step - one GPU->CPU sync per step for
`timestep_map.index(t[0])`, an autograd tape through the classifier - and cannot fuse anything. The closures are
nevertheless almost always *transparent*: `model_fn` forwards `x`, `t`, `y` to our UNet with a skip list chosen from
`t`, and returns the UNet's output untouched; `cond_fn` is `autograd.grad(log_softmax(classifier(x, t))[y].sum(), x) *
scale`. Whether a given pair is transparent is decided here by *tracing*, not by guessing from source text:

  * the UNet and the classifier are put in trace mode (no kernels run): `forward` records its arguments and returns a
    sentinel;
  * `model_fn(x, t_i, **model_kwargs)` is called for every step's original timestep t_i (a CPU tensor, so the closure's
    `index(t[0])` costs no device sync). It is transparent iff it made exactly one UNet call with that very `x`, that
    very `t`, `y` = `model_kwargs["y"]` (or None) and returned the sentinel object itself. The recorded `skip_layer`
    arguments are the per-step skip lists;
  * `cond_fn(x, t_i, **model_kwargs)` is called the same way. The classifier's traced forward returns zero logits through
    an autograd node whose backward records the incoming d/dlogits and returns a tensor of ones. It is transparent iff it
    made exactly one classifier call on a tensor that shares `x`'s storage, the recorded d/dlogits equals
    onehot(y) - softmax(0) (i.e. the closure differentiates sum_n log_softmax(logits)[n, y_n]) and its return value is
    a constant tensor: that constant is the guidance scale.

Only if every step passes is the call routed to `sampler.SchedulePlan` (UNet forwards with skipped blocks elided, native
classifier forward + input-gradient, fused DDIM updates - one CUDA graph, cached per candidate). Anything else - a
closure that post-processes the output, another classifier, eta != 0, `denoised_fn`, `progress`,
`return_all_images` - takes the generic per-step loop (same kernels, no graph), never a CPU path.
`ADB_NO_FAST_PATH=1` disables the recognition.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable, List, Optional

import torch as th

_MAX_PLANS = 8


class _Trace:
    """Put a module in trace mode for the duration of a `with` block (`module._trace` is the call log)."""

    def __init__(self, *modules):
        self.modules = [m for m in modules if m is not None]

    def __enter__(self):
        for m in self.modules:
            m.__dict__["_trace"] = []
        return self

    def __exit__(self, *exc):
        for m in self.modules:
            m.__dict__.pop("_trace", None)


class _TracedLogits(th.autograd.Function):
    """Zero logits that remember what is back-propagated into them."""

    @staticmethod
    def forward(ctx, x, n_out, rec):
        ctx.rec = rec
        ctx.shape = x.shape
        return th.zeros((x.shape[0], n_out), dtype=th.float32, device=x.device)

    @staticmethod
    def backward(ctx, dlogits):
        ctx.rec["dlogits"] = dlogits.detach()
        return th.ones(ctx.shape, dtype=th.float32, device=dlogits.device), None, None


def _find_modules(fn, seen=None, depth=0):
    """Our UNet / classifier modules reachable from a callable: the callable itself, its bound `self`, its closure cells
    (and one level of attributes of `self`-like objects in them - the reference's closures capture `self`)."""
    from .classifier import ClassifierGuidance, EncoderUNetModel
    from .dynamic_unet import Dynamic_UNetModel

    unets, clfs = [], []
    seen = seen if seen is not None else set()

    def visit(obj, d):
        if id(obj) in seen or d > 2:
            return
        seen.add(id(obj))
        if isinstance(obj, Dynamic_UNetModel):
            unets.append(obj)
        elif isinstance(obj, EncoderUNetModel):
            clfs.append(obj)
        elif isinstance(obj, ClassifierGuidance):
            clfs.append(obj.classifier)
        elif callable(obj) and hasattr(obj, "__closure__") and obj.__closure__:
            for cell in obj.__closure__:
                try:
                    visit(cell.cell_contents, d + 1)
                except ValueError:
                    pass
        elif hasattr(obj, "__dict__") and not isinstance(obj, (th.Tensor, th.nn.Module)) and d < 2:
            for v in list(vars(obj).values()):
                if isinstance(v, (Dynamic_UNetModel, EncoderUNetModel, ClassifierGuidance)):
                    visit(v, d + 1)
        if hasattr(obj, "__self__"):
            visit(obj.__self__, d + 1)

    visit(fn, depth)
    return unets, clfs


def try_fast_path(diffusion, model, shape, noise, clip_denoised, cond_fn, model_kwargs, device) -> Optional[Callable]:
    """-> a zero-argument callable that runs the fused plan and returns x_0 (fp32 [B,C,H,W], a fresh tensor), or None when
    the call is not recognised. `noise` must already be the initial x_T on the device."""
    if os.environ.get("ADB_NO_FAST_PATH", "0") == "1":
        return None
    from .classifier import ClassifierGuidance
    from .dynamic_unet import Dynamic_UNetModel
    from .gaussian_diffusion import ModelMeanType, ddim_coefficients
    from .sampler import SchedulePlan

    if diffusion.model_mean_type != ModelMeanType.EPSILON or diffusion.rescale_timesteps:
        return None
    kwargs = dict(model_kwargs or {})
    unets, _ = _find_modules(model)
    if len(unets) != 1:
        return None
    unet: Dynamic_UNetModel = unets[0]
    B = shape[0]
    if tuple(shape) != (B, unet.in_channels, shape[2], shape[3]) or not noise.is_cuda or unet._device() != noise.device:
        return None
    y = kwargs.get("y")
    tmap = list(getattr(diffusion, "timestep_map", range(diffusion.num_timesteps)))
    K = diffusion.num_timesteps
    if len(tmap) != K:
        return None
    clf = None
    native_guidance = isinstance(cond_fn, ClassifierGuidance)
    if cond_fn is not None:
        _, clfs = _find_modules(cond_fn)
        if len(clfs) != 1:
            return None
        clf = clfs[0]
        if clf._device() != noise.device:
            return None

    # ---- trace the closures over every step's original timestep ----
    per_step: List[List[int]] = []
    checks, scales = [], []
    sentinel = unet.io_buffers(B, shape[2], shape[3]).out
    try:
        with _Trace(unet, clf):
            for i in range(K):
                t = th.full((B,), int(tmap[i]), dtype=th.long)  # CPU: the closure's `.index(t[0])` must not sync the device
                unet._trace.clear()
                out = model(noise, t, **kwargs)
                calls = unet._trace
                if len(calls) != 1 or out is not sentinel:
                    return None
                cx, ct, cy, cskip = calls[0]
                if cx is not noise or ct is not t or (cy is not None and cy is not y):
                    return None
                if (cy is None) != (unet.num_classes is None):
                    return None
                per_step.append(sorted(set(int(s) for s in cskip)))
                if cond_fn is not None and not native_guidance:
                    clf._trace.clear()
                    g = cond_fn(noise, t, **kwargs)
                    cc = clf._trace
                    if len(cc) != 1 or y is None or not isinstance(g, th.Tensor) or tuple(g.shape) != tuple(noise.shape):
                        return None
                    gx, gt, rec = cc[0]
                    if gx.data_ptr() != noise.data_ptr() or gt is not t or "dlogits" not in rec:
                        return None
                    checks.append((g, rec["dlogits"]))
    except Exception:
        return None  # a closure that cannot be traced (e.g. does arithmetic on the model output) is not transparent
    scale = None
    if cond_fn is not None:
        if native_guidance:
            scale = cond_fn.classifier_scale
        else:
            # one device read for all steps: the returned tensors are constants (= scale x ones) and the gradient seeds are
            # onehot(y) - softmax(0)
            n_out = checks[0][1].shape[1]
            want = th.nn.functional.one_hot(y.view(-1), n_out).float() - 1.0 / n_out
            ok = th.stack([((g == g.flatten()[0]).all() & th.allclose(dl, want, atol=1e-6)).float() for g, dl in checks])
            vals = th.stack([g.flatten()[0].float() for g, _ in checks])
            ok, vals = ok.cpu(), vals.cpu()
            if not bool(ok.all()) or not bool((vals == vals[0]).all()):
                return None
            scale = float(vals[0])

    # ---- cached fused plan ----
    coefs = tuple(tuple(ddim_coefficients(diffusion, i)) for i in range(K))
    key = (tuple(tmap), coefs, tuple(tuple(s) for s in per_step), B, shape[2], bool(clip_denoised), unet._generation,
           None if clf is None else (id(clf), clf._generation), scale)
    cache: "OrderedDict[tuple, SchedulePlan]" = unet.__dict__.setdefault("_fast_plans", OrderedDict())
    plan = cache.get(key)
    if plan is None:
        guide = None if clf is None else (cond_fn if native_guidance else ClassifierGuidance(clf, scale))
        plan = SchedulePlan(unet, diffusion, per_step, B, image_size=shape[2], clip_denoised=clip_denoised, cond_fn=guide,
                            pack_uint8=False)
        # the plan read the diffusion's tables at build time: it does not keep the (mutable) diffusion object's state
        cache[key] = plan
        while len(cache) > _MAX_PLANS:
            cache.popitem(last=False)
    else:
        cache.move_to_end(key)

    def run():
        return plan.run(noise, y).clone()

    return run
