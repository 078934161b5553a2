#!/usr/bin/env python
"""bench.py — ADM-G 64x64 searched-DDIM candidate sampling throughput (images/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): ADM-G ImageNet-64 class-conditional UNet (295.9 M params) guided
by the noisy classifier (EncoderUNetModel width 128 depth 4, 65.4 M params, classifier_scale 1.0) —
random-init by the oracle-independent recipe below — on the published 10-step searched candidate
(timesteps + block-skip mask of GD/sample_imagenet64_classifier_guidance_dynamic_subnet.sh:13-14),
batch 256 per GPU. This is what the reference's evaluator runs per candidate
(search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-420): every DDIM step = UNet
forward + classifier forward + classifier input-gradient + guided update.
One bench "step" = one full K'=10-step sampling pass over one batch: 10 UNet forwards (9 full +
1 with 9 blocks skipped) + 10 classifier forward/backward passes + 10 fused DDIM updates + uint8 pack,
all in one CUDA graph. The UNet-only variant (no cond_fn; SURVEY §8d reports both) is timed too and
reported under "unet_only".

  value : images/s, inputs (x_T, y) already resident in HBM, whole schedule as one CUDA graph.
  e2e   : same through the public API with HOST buffers: pinned x_T/y -> H2D, sampling, uint8
          NHWC images -> D2H, every step inside the timed region.
  roofline : the dominant kernel (tcgen05 implicit-GEMM conv / k=1 GEMM): algorithmic FLOPs of its
          launches / their CUDA-event durations measured here, vs MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle (torch fp32 restatement of the reference) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAND10 = {
    "timesteps": [744, 137, 647, 856, 305, 441, 676, 572, 971, 85],
    "skip_layers": [[], [], [], [], [], [], [30, 10, 39, 4, 15, 46, 49, 54, 8], [], [], []],
}
ADM_FLAGS = dict(attention_resolutions="32,16,8", class_cond=True, diffusion_steps=1000, dropout=0.1, image_size=64,
                 learn_sigma=True, noise_schedule="cosine", num_channels=192, num_head_channels=64, num_res_blocks=3,
                 resblock_updown=True, use_new_attention_order=True, use_fp16=True, use_scale_shift_norm=True,
                 use_dynamic_unet=True)
# GFLOP per image of the published 10-step candidate (9 x 219.356 + 182.431)
GFLOP_PER_IMAGE = 2156.64
METRIC = "ADM-G 64x64 images/s, 10-step searched DDIM + block-skip mask, classifier-guided"
CLASSIFIER = dict(classifier_depth=4, classifier_width=128)  # ADM-G 64x64 noisy classifier (GD/README flags)
# BASELINE configs[0]'s 4-step searched schedule, full architecture (GD/scripts/classifier_sample_generate_image.py:159-168)
CAND4 = {"timesteps": [153, 424, 926, 690], "skip_layers": [[], [], [], []]}
GFLOP_PER_IMAGE_4STEP = 877.42  # 4 x 219.356


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--dump-ops", default=None, help="write every recorded op's kind/flops/bytes/ms of one step to this CSV")
    ap.add_argument("--unet-only", action="store_true", help="headline = the UNet-only variant (no classifier cond_fn)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary measurements of the default line: 4-step schedule, stock-API call, same-box "
                         "torch-eager arm (N=1), population evaluation (N>1)")
    ap.add_argument("--pop-candidates", type=int, default=0,
                    help="population_eval (N>1): number of candidates (default 6 x N + 1; BASELINE configs[2]'s 50 at N=8)")
    ap.add_argument("--sd-sampler", default="ddim", choices=["ddim", "plms", "dpm"],
                    help="sdv1 workload: searched-timestep DDIM (BASELINE configs[4]), PLMS, or DPM-Solver++(2M)")
    ap.add_argument("--workload", default="admg64", choices=["admg64", "lsun256", "sdv1"],
                    help="admg64 = BASELINE configs[1] (default, the metric's config); lsun256 = configs[3]")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bench_weights(shapes, seed=0):
    """Deterministic random-init (no checkpoint offline): fan-in scaled normals so activations stay O(1).
    zero_module'd tensors of the reference init are re-drawn too, otherwise the net outputs exactly 0."""
    import torch
    import zlib

    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        norm = any(k in name for k in (".in_layers.0.", ".out_layers.0.", ".norm.")) or name.startswith("out.0.")
        if name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name == "label_emb.weight":
            t = 0.3 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / fan_in ** 0.5
        sd[name] = t
    return sd


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of the reference on host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(batch, n_ddim_steps, warm=True, guided=True, device="cpu"):
    """images/s of the oracle on cand10, measured on `n_ddim_steps` consecutive schedule
    positions (starting at the first sampled step) at batch `batch`, scaled to the 10-step schedule.
    device="cuda" runs the same torch code through PyTorch's own CUDA kernels (the `--impl torch-eager` arm)."""
    import contextlib

    import torch
    from oracle import diffusion_ref, unet_ref, weights

    torch.set_num_threads(os.cpu_count() or 1)
    on_gpu = str(device).startswith("cuda")
    cfg = unet_ref.adm_g64_config()
    sd = {k: v.to(device) for k, v in weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0).items()}
    cond_fn = None
    if guided:
        ccfg = unet_ref.classifier64_config(depth=CLASSIFIER["classifier_depth"], width=CLASSIFIER["classifier_width"])
        csd = {k: v.to(device) for k, v in weights.make_state_dict(unet_ref.param_shapes(ccfg, encoder_only=True), seed=1).items()}
        cond_fn = unet_ref.classifier_cond_fn(csd, ccfg, 1.0)
    base = diffusion_ref.base_tables("cosine", 1000)
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], CAND10["timesteps"])
    tb = diffusion_ref.diffusion_tables(nb)
    noise = torch.randn(batch, 3, 64, 64, generator=torch.Generator().manual_seed(2)).to(device)
    y = torch.randint(0, 1000, (batch,), generator=torch.Generator().manual_seed(3)).to(device)
    unet = lambda x, t, yy, skip: unet_ref.unet_forward(sd, cfg, x, t, yy, skip)
    model_fn = diffusion_ref.make_model_fn(unet, tmap)
    K = len(tmap)

    def run_steps(positions):
        x = noise
        # on the GPU: tensors the oracle creates land on the device; convolutions / matmuls in fp16 as the reference's
        # use_fp16=True torso (autocast), GroupNorm and the sampler arithmetic in fp32
        ctx = contextlib.ExitStack()
        if on_gpu:
            ctx.enter_context(torch.device(device))
            ctx.enter_context(torch.autocast("cuda", dtype=torch.float16))
            torch.cuda.synchronize()
        with ctx:
            t0 = time.perf_counter()
            for i in positions:
                one = {k: v[i:i + 1] for k, v in tb.items()}
                # a 1-step schedule = step i of the 10-step one in isolation (same ops, same tables)
                x = diffusion_ref.ddim_sample_loop(
                    lambda xx, ts, **kw: model_fn(xx, ts, **{**kw}), x.shape, one, [tmap[i]], x, True, cond_fn=cond_fn,
                    model_kwargs={"y": y, "skip_layers": CAND10["skip_layers"]})
            if on_gpu:
                torch.cuda.synchronize()
            return time.perf_counter() - t0

    order = list(range(K))[::-1]
    if warm:
        run_steps(order[:1])
    return run_steps, order, K


def run_torch_eager(args):
    """Not the reference arm and not our product: the oracle's torch code run through PyTorch's own CUDA kernels
    (cuDNN / cuBLAS, fp16 autocast) on one B200 - what a user of the reference gets on this hardware (SURVEY §8d:
    "the real bar"). Full 10-step candidate at --batch images, wall clock with synchronisation on both sides."""
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.backends.cudnn.benchmark = True
    batch = args.batch
    run_steps, order, K = cpu_reference_rate(batch, 1, guided=not args.unet_only, device="cuda")
    for _ in range(max(1, args.warmup)):
        run_steps(order)
    times = [run_steps(order) for _ in range(args.steps)]
    mean = sum(times) / len(times)
    print(json.dumps({
        "impl": "torch-eager", "metric": METRIC if not args.unet_only else METRIC.replace("classifier-guided", "UNet-only"),
        "value": batch / mean, "unit": "images/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "dtype": "fp16 autocast", "data": "synthetic",
        "config": {"workload": "ADM-G 64x64 cand10, " + ("UNet-only" if args.unet_only else "classifier-guided") +
                               ", the oracle's PyTorch code on cuda:0 (cuDNN/cuBLAS eager, fp16 autocast, cudnn.benchmark)",
                   "batch": batch, "torch": torch.__version__},
    }), flush=True)


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    run_steps, order, K = cpu_reference_rate(batch, 1, guided=not args.unet_only)
    # every bench step = ONE schedule position (rotating through the 10), so the run stays bounded
    for w in range(args.warmup):
        run_steps([order[w % K]])
    times = []
    for s in range(args.steps):
        times.append(run_steps([order[s % K]]))
    mean_pos = sum(times) / len(times)
    value = batch / (mean_pos * K)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_pos * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ADM-G 64x64 cand10 (10 searched steps + skip mask), " +
                               ("UNet-only" if args.unet_only else "classifier-guided (depth-4 width-128 noisy classifier, autograd input gradient)") +
                               ", CPU oracle port of the reference", "batch": batch},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"one DDIM step (UNet fwd + classifier fwd/bwd + update) per bench step at batch {batch}, rotating through the "
                                   f"10 schedule positions; images/s = batch / (10 x mean step time)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from autodiffusion_b200 import (classifier_defaults, create_classifier, create_model_and_diffusion,
                                    model_and_diffusion_defaults)
    from autodiffusion_b200.classifier import ClassifierGuidance
    from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate

    d = model_and_diffusion_defaults()
    d.update(ADM_FLAGS)
    model, diffusion = create_model_and_diffusion(**d)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(bench_weights(shapes))
    model.to(dev).eval()
    model.convert_to_fp16()
    cd = classifier_defaults()
    cd.update(CLASSIFIER)
    clf = create_classifier(**cd)
    clf.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in clf.state_dict().items()}, seed=1))
    clf.to(dev).eval()
    guidance = ClassifierGuidance(clf, 1.0)

    B = args.batch
    active, per_step = resolve_candidate(CAND10, diffusion)
    K = active.num_timesteps

    def build(cond_fn):
        t0 = time.time()
        pl = SchedulePlan(model, active, per_step, B, clip_denoised=True, cond_fn=cond_fn, pack_uint8=True)
        torch.cuda.synchronize()
        return pl, time.time() - t0

    plan_u, t_build_u = build(None)
    plan_g, t_build_g = build(guidance)
    _, t_rebuild = build(None if args.unet_only else guidance)  # every per-mask UNet plan / classifier operand is cached now
    head = plan_u if args.unet_only else plan_g
    other = plan_g if args.unet_only else plan_u

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise_dev = torch.randn(head.shape, generator=g, device=dev)
    y_dev = torch.randint(0, 1000, (B,), generator=g, device=dev)
    noise_host = noise_dev.cpu().pin_memory()
    y_host = y_dev.cpu().pin_memory()
    u8_host = torch.empty((B, 64, 64, 3), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        cs = ClockSampler(local) if sample_clocks else None
        if cs:
            cs.start()
            time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        w1 = time.time()
        ms = e0.elapsed_time(e1)
        clocks = cs.stop(w0, w1) if cs else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    def step_e2e():
        head.run(noise_host, y_host)  # H2D of x_T and y inside
        u8_host.copy_(head.u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs the images (…progressive.py:427)

    ms, clocks = timed(lambda: head.run(noise_dev, y_dev), args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    ms_other, _ = timed(lambda: other.run(noise_dev, y_dev), args.steps, max(1, args.warmup // 2))

    extras = {}
    if not args.no_extras:
        # (1) the metric's other schedule: 4-step searched DDIM, full architecture (BASELINE "4/10-step"), same batch
        act4, per4 = resolve_candidate(CAND4, diffusion)
        plan4 = SchedulePlan(model, act4, per4, B, clip_denoised=True, cond_fn=None if args.unet_only else guidance, pack_uint8=True)
        ms4, _ = timed(lambda: plan4.run(noise_dev, y_dev), args.steps, 3)
        v4 = B * world * args.steps / (ms4 * 1e-3)
        gf4 = GFLOP_PER_IMAGE_4STEP + (0 if args.unet_only else act4.num_timesteps * sum(fl for (_, fl, _) in plan_g.guidance.op_info()) / B / 1e9)
        extras["four_step"] = {"value": v4, "unit": "images/s", "ms_per_step": ms4 / args.steps, "timesteps": CAND4["timesteps"],
                               "gflop_per_image": round(gf4, 2), "tflops_effective": v4 / world * gf4 / 1e3,
                               "gpu_launches_per_step": plan4.launches}
        del plan4
        # (2) the same candidate through the reference's own call: `diffusion.ddim_sample_loop(model_fn, shape, ...,
        # cond_fn=cond_fn)` with the search script's closures (…progressive.py:383-420), recognised by tracing and fused
        # (fastpath.py); per batch: labels drawn on the device, uint8 NHWC conversion and `.cpu()` as the reference does
        extras["stock_api"] = stock_api_rate(model, clf, diffusion, B, dev, args, timed, world)
        if world > 1:
            from autodiffusion_b200.population import run_population

            # 8 GPUs: BASELINE configs[2]'s 50 candidates (48 whole + 2 batch-sharded); fewer ranks: 6 per rank + 1, so that
            # the batch-sharded tail and its NCCL moment all-reduce are exercised at every N > 1 within the same ~30 s
            n_c = args.pop_candidates or (50 if world == 8 else 6 * world + 1)
            # batches that divide the 1000 samples (250 at the default 256): no partly used last batch
            pb = max(b for b in range(1, B + 1) if 1000 % b == 0)
            pop = run_population(model, diffusion, None if args.unet_only else guidance, n_c, num_samples=1000, batch_size=pb,
                                 fid_method="eigh")
            pop.pop("fids")
            pop["vs_sampling_rate"] = pop["images_per_s"] / (B * world * args.steps / (ms * 1e-3))
            extras["population_eval"] = pop
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)
    e2e_value = imgs / (ms_e2e * 1e-3)
    other_value = imgs / (ms_other * 1e-3)

    # algorithmic GFLOP per image: UNet from SURVEY §8(d) (hook-measured on the reference, cross-checked against
    # the recorded plan below); classifier = the recorded forward + data-gradient products of one guidance pass
    unet_rec = sum(fl for up in plan_u.steps for (_, fl, _) in up.plan.op_info()) / B / 1e9
    clf_per_step = sum(fl for (_, fl, _) in plan_g.guidance.op_info()) / B / 1e9
    gflop_u, gflop_g = GFLOP_PER_IMAGE, GFLOP_PER_IMAGE + K * clf_per_step
    gflop_head = gflop_u if args.unet_only else gflop_g
    variant = lambda v, gf, msv, pl: {"value": v, "unit": "images/s", "ms_per_step": msv / args.steps,
                                      "gflop_per_image": round(gf, 2), "tflops_effective": v / world * gf / 1e3,
                                      "gpu_launches_per_step": pl.launches}

    line = {
        "metric": METRIC if not args.unet_only else METRIC.replace(", classifier-guided", ", UNet-only"),
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "ADM-G 64x64 UNet (295.9M params) + noisy classifier (65.4M params, depth 4 width 128, scale 1.0), "
                               "random-init; published 10-step searched candidate (timesteps + block-skip mask); one step = full "
                               "10-step sampling of one batch: per DDIM step UNet forward" +
                               (" (UNet-only variant: no cond_fn)" if args.unet_only else
                                " + classifier forward + classifier input-gradient + guided update") + ", one CUDA graph",
                   "batch_per_gpu": B, "ddim_steps": K,
                   "l2": "activations per launch (>=400 MB at batch 256) exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world} (independent batches per rank, no data-path collective)"},
        "ms_per_unet_fwd": ms_other / args.steps / K if not args.unet_only else ms / args.steps / K,
        "gflop_per_image": round(gflop_head, 2),
        "tflops_effective": value / world * gflop_head / 1e3,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": noise_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": u8_host.numel()},
        "gpu_launches": head.launches * args.steps,
        ("guided" if args.unet_only else "unet_only"): variant(other_value, gflop_g if args.unet_only else gflop_u, ms_other, other),
        "flop_accounting": {"unet_gflop_per_image_survey": gflop_u, "unet_gflop_per_image_recorded_plan": round(unet_rec, 2),
                            "classifier_fwd_bwd_gflop_per_image_per_step": round(clf_per_step, 3)},
        "plan_build_s": {"first_candidate_unet": round(t_build_u, 3), "first_candidate_classifier_added": round(t_build_g, 3),
                         "next_candidate_same_masks": round(t_rebuild, 4)},
        "clocks": clocks,
    }
    line.update(extras)

    if rank == 0 and not args.no_roofline:
        # per-kernel times: the same recorded ops, eager on the current stream with an event pair around each
        pk = peaks()
        info, ms_ops = [], []
        for up in head.steps:
            up.plan.run_profiled()  # warm
        if head.guidance is not None:
            head.guidance.run_profiled()
        for n, up in enumerate(head.steps):
            head.t_in.fill_(head.t_values[n])
            info += up.plan.op_info()
            ms_ops += up.plan.run_profiled()
            if head.guidance is not None:
                info += [("clf:" + k if k not in ("conv_igemm",) else k, fl, by) for (k, fl, by) in head.guidance.op_info()]
                ms_ops += head.guidance.run_profiled()
            # the fused guidance + DDIM update, timed like every other kernel: recorded into a one-op plan and run with a
            # CUDA-event pair around the launch itself (adb_plan_run_profiled), not around a Python call
            from autodiffusion_b200 import ops as _ops

            upd = _ops.Plan()
            _ops.ddim_step(head.x, head.model_out, head.grad, head.coefs[n], head.clip_denoised, x_prev=head.x, plan=upd)
            upd.run_profiled()
            info += upd.op_info()
            ms_ops += upd.run_profiled()
        if head.u8 is not None:
            from autodiffusion_b200 import ops as _ops

            pk_plan = _ops.Plan()
            _ops.pack_uint8(head.final, out=head.u8, plan=pk_plan)
            pk_plan.run_profiled()
            info += pk_plan.op_info()
            ms_ops += pk_plan.run_profiled()
        agg = {}
        for (kind, fl, by), t in zip(info, ms_ops):
            a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += t
            a[2] += fl
            a[3] += by
        total_ms = sum(ms_ops)
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                f.write("idx,kind,flops,bytes,ms\n")
                for i, ((kind, fl, by), t) in enumerate(zip(info, ms_ops)):
                    f.write(f"{i},{kind},{fl:.0f},{by:.0f},{t:.5f}\n")
        conv = agg["conv_igemm"]
        achieved = conv[2] / (conv[1] * 1e-3) / 1e12
        traffic, traffic_detail = None, None
        tp = os.path.join(ROOT, "profiles", "conv_igemm_traffic.json")
        if os.path.exists(tp):  # dram bytes per launch of this kernel from the committed `ncu --set full` capture
            traffic_detail = json.load(open(tp))
            traffic = traffic_detail["dram_bytes_per_launch_avg"]
        line["roofline"] = {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM: 3x3/1x1 conv, qkv/proj, and their data gradients)",
                            "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                            "frac": achieved / pk["tf_sustained"], "traffic": traffic, "traffic_detail": traffic_detail,
                            "peak_source": pk["source"] + " (sustained bf16: kernel timed inside a long step)",
                            "launches": conv[0], "share_of_step": conv[1] / total_ms,
                            "flops_per_launch_avg": conv[2] / conv[0], "ms_per_launch_avg": conv[1] / conv[0]}
        line["kernel_breakdown"] = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / total_ms, 4),
                                        "tflops": (v[2] / (v[1] * 1e-3) / 1e12) if v[2] and v[1] else None,
                                        "gbs": (v[3] / (v[1] * 1e-3) / 1e9) if v[3] and v[1] else None,
                                        "frac_of_hbm_peak": (v[3] / (v[1] * 1e-3) / 1e9 / pk["hbm"]) if v[3] and v[1] else None}
                                    for k, v in agg.items()}
    if rank == 0 and not args.no_cpu_baseline:
        import torch as _t

        run_steps, order, Kc = cpu_reference_rate(args.ref_batch, 1, guided=not args.unet_only)
        tt = [run_steps([order[i]]) for i in (0, 3, 6)]  # includes position 6 (t=676), the masked step
        mean_pos = sum(tt) / len(tt)
        line["cpu_baseline"] = {"value": args.ref_batch / (mean_pos * Kc), "unit": "images/s", "cores": _t.get_num_threads(),
                                "kind": "port",
                                "sample": f"3 of the 10 DDIM steps of the same candidate (incl. the masked one) at batch "
                                          f"{args.ref_batch}, fp32 torch CPU oracle (UNet forward" +
                                          ("" if args.unet_only else " + classifier forward + autograd input gradient") +
                                          "); images/s = batch / (10 x mean step time)"}
    if rank == 0 and world == 1 and not args.no_extras and not args.unet_only:
        # BASELINE configs[3] and configs[4] in the driver-visible line (their own batch sizes, 3 timed steps each): the
        # same code paths as `--workload lsun256` / `--workload sdv1`
        try:
            del plan_u, plan_g, head, other
            torch.cuda.empty_cache()
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup, sub.batch, sub.dump_ops, sub.sd_sampler, sub.impl = 3, 2, 256, None, "ddim", "ours"
            keep = ("metric", "value", "unit", "ms_per_step", "ms_per_unet_fwd", "tflops_effective", "e2e", "config", "roofline")
            other_cfg = {}
            for name, fn in (("lsun256", run_lsun), ("sdv1", run_sdv1)):
                r = fn(sub, inner=True)
                other_cfg[name] = {k: r[k] for k in keep if k in r}
                torch.cuda.empty_cache()
            line["other_configs"] = other_cfg
        except Exception as e:
            line["other_configs"] = {"unavailable": repr(e)[:300]}
        # the hardware-matched bar: the reference's torch code (oracle restatement) through PyTorch's own CUDA
        # kernels on this same GPU - one full 10-step candidate at the same batch, after one warm-up pass
        try:
            torch.cuda.empty_cache()
            torch.backends.cudnn.benchmark = True
            run_steps, order, _ = cpu_reference_rate(B, 1, guided=True, device="cuda")
            run_steps(order)
            dt = run_steps(order)
            line["torch_eager_same_box"] = {"value": B / dt, "unit": "images/s", "ms_per_step": dt * 1e3, "batch": B,
                                            "what": "the oracle's PyTorch code on cuda:0 (cuDNN/cuBLAS eager, fp16 autocast, "
                                                    "autograd classifier gradient), same candidate, 1 timed pass",
                                            "ours_over_it": value / (B / dt)}
        except Exception as e:  # never lose the headline to the comparison arm
            line["torch_eager_same_box"] = {"unavailable": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stock_api_rate(model, clf, diffusion, B, dev, args, timed, world):
    """images/s of the reference's sampling block run UNMODIFIED on our objects (closures as in
    search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:383-420)."""
    import copy
    import types

    import torch as th
    import torch.nn.functional as F

    from autodiffusion_b200.respace import reset_diffusion

    self = types.SimpleNamespace(model=model, classifier=clf, active_diffusion=copy.deepcopy(diffusion))
    reset_diffusion(CAND10["timesteps"], self.active_diffusion, diffusion)
    a = types.SimpleNamespace(classifier_scale=1.0, class_cond=True, clip_denoised=True, image_size=64, batch_size=B)
    skip_layers = CAND10["skip_layers"]

    def cond_fn(x, t, y=None, skip_layers=None, timesteps=None):
        assert y is not None
        with th.enable_grad():
            x_in = x.detach().requires_grad_(True)
            logits = self.classifier(x_in, t)
            log_probs = F.log_softmax(logits, dim=-1)
            selected = log_probs[range(len(logits)), y.view(-1)]
            return th.autograd.grad(selected.sum(), x_in)[0] * a.classifier_scale

    def model_fn(x, t, y=None, skip_layers=None, timesteps=None):
        assert y is not None
        t_index = self.active_diffusion.timestep_map.index(t[0])
        skip_layer = skip_layers[t_index]
        return self.model(x, t, y if a.class_cond else None, skip_layer=skip_layer)

    def one_batch():
        model_kwargs = {}
        classes = th.randint(low=0, high=1000, size=(a.batch_size,), device=dev)
        model_kwargs["y"] = classes
        model_kwargs["skip_layers"] = skip_layers
        sample = self.active_diffusion.ddim_sample_loop(
            model_fn, (a.batch_size, 3, a.image_size, a.image_size), clip_denoised=a.clip_denoised,
            model_kwargs=model_kwargs, cond_fn=None if args.unet_only else cond_fn, device=dev)
        sample = ((sample + 1) * 127.5).clamp(0, 255).to(th.uint8)
        sample = sample.permute(0, 2, 3, 1).contiguous()
        return sample.cpu().numpy(), classes.cpu().numpy()

    n0 = len(model.__dict__.get("_fast_plans", {}))
    ms, _ = timed(one_batch, args.steps, 2)
    fused = len(model.__dict__.get("_fast_plans", {})) > n0
    return {"value": B * world * args.steps / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / args.steps,
            "recognised_and_fused": fused,
            "what": "diffusion.ddim_sample_loop(model_fn, shape, model_kwargs, cond_fn) with the search script's closures, labels "
                    "drawn per batch, uint8 NHWC conversion and .cpu() of images + labels inside the timed region"}


# ------------------------------------------------------------------------------------------
# secondary workload: BASELINE configs[3], ADM unconditional LSUN-bedroom 256x256
# ------------------------------------------------------------------------------------------
LSUN_FLAGS = dict(attention_resolutions="32,16,8", class_cond=False, diffusion_steps=1000, dropout=0.1, image_size=256,
                  learn_sigma=True, noise_schedule="linear", num_channels=256, num_head_channels=64, num_res_blocks=2,
                  resblock_updown=True, use_fp16=True, use_scale_shift_norm=True)
# the first 10 of the 15 published searched steps (GD/sample_LSUN_bedroom_subnet.sh:9)
LSUN_CAND = {"timesteps": [644, 737, 67, 804, 134, 871, 6, 639, 268, 335], "skip_layers": [[] for _ in range(10)]}
LSUN_GFLOP_PER_FWD = 2239.67  # hook-measured on the reference module


def run_lsun(args, inner=False):
    """images/s of the unconditional LSUN-256 model (552.8 M params) on a 10-step searched schedule, batch 64 per
    GPU (the reference's LSUN search has no classifier: search_uncondition_model.py). inner=True: called from the
    default bench line (process group already set up); returns the result dict instead of printing it."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not inner:
        dist.init_process_group("nccl", device_id=dev)
    from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults
    from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate

    d = model_and_diffusion_defaults()
    d.update(LSUN_FLAGS)
    model, diffusion = create_model_and_diffusion(**d)
    model.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in model.state_dict().items()}))
    model.to(dev).eval()
    model.convert_to_fp16()
    B = args.batch if args.batch != 256 else 64
    active, per_step = resolve_candidate(LSUN_CAND, diffusion)
    K = active.num_timesteps
    plan = SchedulePlan(model, active, per_step, B, clip_denoised=True, cond_fn=None, pack_uint8=True)
    noise = torch.randn(plan.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    noise_host = noise.cpu().pin_memory()
    u8_host = torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def e2e():
        plan.run(noise_host, None)
        u8_host.copy_(plan.u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms = timed(lambda: plan.run(noise, None), args.steps, args.warmup)
    ms_e2e = timed(e2e, args.steps, 1)
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)
    info, ms_ops = [], []
    for up in plan.steps[:1]:
        up.plan.run_profiled()
        info = up.plan.op_info()
        ms_ops = up.plan.run_profiled()
    agg = {}
    for (kind, fl, by), t in zip(info, ms_ops):
        a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += fl
        a[3] += by
    if rank == 0 or inner:
        res_line = ({
            "metric": "ADM LSUN-bedroom 256x256 images/s, 10-step searched DDIM (unconditional)", "value": value,
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ADM unconditional LSUN-256 UNet (552.8M params, random-init), first 10 of the 15 published "
                                   "searched timesteps, full architecture, one step = 10-step sampling of one batch, one CUDA graph",
                       "batch_per_gpu": B, "ddim_steps": K},
            "ms_per_unet_fwd": ms / args.steps / K,
            "tflops_effective": value / world * K * LSUN_GFLOP_PER_FWD / 1e3,
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": noise_host.numel() * 4,
                    "d2h_bytes_per_step": u8_host.numel()},
            "gpu_launches": plan.launches * args.steps,
            "kernel_breakdown_one_forward": {k: {"launches": v[0], "ms": round(v[1], 3),
                                                 "tflops": (v[2] / (v[1] * 1e-3) / 1e12) if v[2] and v[1] else None,
                                                 "gbs": (v[3] / (v[1] * 1e-3) / 1e9) if v[3] and v[1] else None}
                                             for k, v in agg.items()},
        })
        if inner:
            del plan
            return res_line
        print(json.dumps(res_line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# secondary workload: BASELINE configs[4], Stable Diffusion v1 UNet on 64x64x4 latents, CFG 7.5
# ------------------------------------------------------------------------------------------
SD_CAND = [981, 861, 741, 641, 501, 421, 301, 201, 121, 21]  # a searched-style 10-step subsequence of the 1000 DDPM steps
SD_GFLOP_PER_FWD = 803.27  # per image per UNet forward (x2 per step under CFG)


def sd_cpu_rate(n_steps_timed=1):
    """The oracle port of the reference's CFG DDIM step on the host cores: batch 1 (2 UNet forwards per step)."""
    import torch

    from oracle import sd_unet_ref as R

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = R.sd_v1_config()
    sd = R.make_weights(cfg, seed=0)
    g = torch.Generator().manual_seed(0)
    x, c, uc = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    t0 = time.time()
    R.ddim_sample(lambda xx, tt, cc: R.unet_forward(sd, cfg, xx, tt, cc), x, c, uc, 7.5, SD_CAND[:n_steps_timed], R.sd_alphas_cumprod())
    dt = time.time() - t0
    return 1.0 / (dt / n_steps_timed * len(SD_CAND)), dt


def run_sdv1(args, inner=False):
    """latent images/s of the SD-v1 UNet (859.5 M params) under the searched 10-step DDIM with CFG 7.5, batch 32 per
    GPU (scripts/search_ea.py:504-538 without the text encoder / VAE, which are outside the searched path)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "torch-eager":  # the oracle's torch code through PyTorch's own CUDA kernels (fp16 autocast): the B200 bar
        if rank == 0:
            import torch

            from oracle import sd_unet_ref as R

            torch.backends.cudnn.benchmark = True
            cfg = R.sd_v1_config()
            sd = {k: v.cuda() for k, v in R.make_weights(cfg, seed=0).items()}
            B = args.batch if args.batch != 256 else 32
            g = torch.Generator().manual_seed(0)
            x = torch.randn(B, 4, 64, 64, generator=g).cuda()
            c, uc = torch.randn(B, 77, 768, generator=g).cuda(), torch.randn(B, 77, 768, generator=g).cuda()

            acp = R.sd_alphas_cumprod()  # host-side schedule, as the reference builds it

            def once():
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.device("cuda"), torch.autocast("cuda", dtype=torch.float16):
                    R.ddim_sample(lambda xx, tt, cc: R.unet_forward(sd, cfg, xx, tt, cc).float(), x, c, uc, 7.5, SD_CAND, acp)
                torch.cuda.synchronize()
                return time.perf_counter() - t0

            for _ in range(max(1, args.warmup)):
                once()
            dt = sum(once() for _ in range(args.steps)) / args.steps
            print(json.dumps({"impl": "torch-eager", "metric": "SD-v1 latent images/s, 10-step searched DDIM, CFG 7.5", "value": B / dt,
                              "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                              "higher_is_better": True, "dtype": "fp16 autocast", "data": "synthetic",
                              "config": {"workload": "SD-v1 UNet, CFG 7.5, 10 searched DDIM steps, the oracle's PyTorch code on cuda:0 "
                                                     "(cuDNN/cuBLAS eager, fp16 autocast, unfused attention as the reference's einsum path)",
                                         "batch": B, "torch": torch.__version__}}), flush=True)
        return
    if args.impl == "reference":
        if rank == 0:
            rate, dt = sd_cpu_rate(1)
            print(json.dumps({"impl": "reference", "metric": "SD-v1 latent images/s, 10-step searched DDIM, CFG 7.5", "value": rate,
                              "unit": "images/s", "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                              "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": "SD-v1 UNet, CFG 7.5, 10 searched steps; sample: 1 of the 10 steps at batch 1"},
                              "cpu_baseline": {"value": rate, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                               "sample": "one CFG DDIM step (2 UNet forwards) at batch 1, extrapolated to 10 steps"},
                              "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not inner:
        dist.init_process_group("nccl", device_id=dev)
    from autodiffusion_b200.sd_ddim import CandidatePlan, DPMCandidatePlan, LatentDiffusionUNet
    from autodiffusion_b200.sd_unet import UNetModel

    unet = UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                     num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                     context_dim=768, use_checkpoint=True, legacy=False)
    unet.load_state_dict(bench_weights({k: tuple(v.shape) for k, v in unet.state_dict().items()}))
    unet.to(dev).eval()
    ld = LatentDiffusionUNet(unet)
    B = args.batch if args.batch != 256 else 32
    K = len(SD_CAND)
    t0 = time.time()
    if args.sd_sampler == "dpm":  # 10 model evaluations = 11 searched time points (search_ea.py keeps time_step + 1)
        plan = DPMCandidatePlan(unet, ld.alphas_cumprod, SD_CAND + [0], B, (4, 64, 64), 7.5, True)
    else:
        plan = CandidatePlan(unet, ld.alphas_cumprod, SD_CAND, B, (4, 64, 64), 7.5, True, method=args.sd_sampler)
    fwd_per_image = (K + (1 if args.sd_sampler == "plms" else 0)) * 2  # PLMS: one extra evaluation in its first step
    torch.cuda.synchronize()
    build_s = time.time() - t0
    t0 = time.time()  # a second candidate of the same geometry: only its own chain is captured
    other = CandidatePlan(unet, ld.alphas_cumprod, [t - 10 for t in SD_CAND], B, (4, 64, 64), 7.5, True)
    torch.cuda.synchronize()
    build_next_s = time.time() - t0
    del other
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x_T = torch.randn((B, 4, 64, 64), device=dev, generator=g)
    cond = torch.randn((B, 77, 768), device=dev, generator=g)
    uncond = torch.randn((1, 77, 768), device=dev, generator=g).repeat(B, 1, 1).contiguous()
    hx, hc, hu = x_T.cpu().pin_memory(), cond.cpu().pin_memory(), uncond.cpu().pin_memory()
    hout = torch.empty((B, 4, 64, 64), dtype=torch.float32).pin_memory()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def e2e():
        plan.run(hx, hc, hu)
        hout.copy_(plan.x, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    w0 = time.time()
    ms = timed(lambda: plan.run(x_T, cond, uncond), args.steps, args.warmup)
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    ms_e2e = timed(e2e, args.steps, 1)
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)
    plan.plan_fwd.run_profiled()
    info = plan.plan_fwd.op_info()
    ms_ops = plan.plan_fwd.run_profiled()
    agg = {}
    for (kind, fl, by), t in zip(info, ms_ops):
        a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += fl
        a[3] += by
    if args.dump_ops and rank == 0:
        with open(args.dump_ops, "w") as f:
            f.write("idx,kind,flops,bytes,ms\n")
            for i, ((kind, fl, by), t) in enumerate(zip(info, ms_ops)):
                f.write(f"{i},{kind},{fl},{by},{t}\n")
    fwd_ms = sum(ms_ops)
    pk = peaks()
    conv = agg.get("conv_igemm", [0, 0.0, 0.0, 0.0])
    roof = None
    if conv[1] > 0:
        ach = conv[2] / (conv[1] * 1e-3) / 1e12
        roof = {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM: every conv / Linear of the SD UNet)", "bound": "tensor",
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": None,
                "peak_source": pk["source"] + " (sustained bf16)", "launches": conv[0] * K * args.steps,
                "share_of_step": conv[1] / fwd_ms, "flops_per_launch_avg": conv[2] / conv[0], "ms_per_launch_avg": conv[1] / conv[0]}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not inner:
        rate, dt = sd_cpu_rate(1)
        cpu = {"value": rate, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"one CFG DDIM step (2 UNet forwards, 859.5M params) at batch 1 on the host: {dt:.1f} s, extrapolated to 10 steps"}
    recorded_gflop = sum(fl for _, fl, _ in info) / (2 * B) / 1e9
    if rank == 0 or inner:
        res_line = ({
            "metric": "Stable Diffusion v1 UNet latent images/s (64x64x4 latents = 512x512), 10-step searched "
                      + {"ddim": "DDIM", "plms": "PLMS", "dpm": "DPM-Solver++(2M)"}[args.sd_sampler] + ", CFG 7.5",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "SD-v1 UNet (859.5M params, random-init) on 64x64x4 latents, synthetic 77x768 text context + "
                                   "unconditional context, CFG 7.5 (batched [uncond | cond] forward), 10 searched DDIM steps; one "
                                   "step = full sampling of one batch as one CUDA graph (context K/V projected once)",
                       "batch_per_gpu": B, "ddim_steps": K, "timesteps": SD_CAND,
                       "l2": "activations per launch exceed the 126 MB L2; no explicit flush"},
            "ms_per_unet_fwd": ms / args.steps / (fwd_per_image // 2), "gflop_per_image": fwd_per_image * SD_GFLOP_PER_FWD,
            "tflops_effective": value / world * fwd_per_image * SD_GFLOP_PER_FWD / 1e3, "sampler": args.sd_sampler,
            "flop_accounting": {"survey_gflop_per_image_per_forward": SD_GFLOP_PER_FWD,
                                "recorded_plan_gflop_per_image_per_forward": recorded_gflop},
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": (hx.numel() + hc.numel() + hu.numel()) * 4, "d2h_bytes_per_step": hout.numel() * 4},
            "gpu_launches": plan.launches * args.steps, "plan_build_s": build_s, "next_candidate_plan_build_s": build_next_s, "clocks": clocks, "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_breakdown_one_forward": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / fwd_ms, 4),
                                                 "tflops": (v[2] / (v[1] * 1e-3) / 1e12) if v[2] and v[1] else None,
                                                 "gbs": (v[3] / (v[1] * 1e-3) / 1e9) if v[3] and v[1] else None}
                                             for k, v in agg.items()},
        })
        if inner:
            del plan
            return res_line
        print(json.dumps(res_line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "sdv1":
        run_sdv1(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "lsun256":
        run_lsun(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
