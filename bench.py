#!/usr/bin/env python
"""bench.py — ADM-G 64x64 searched-DDIM candidate sampling throughput (images/s) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): ADM-G ImageNet-64 class-conditional UNet (295.9 M params,
random-init by oracle-independent recipe below), the published 10-step searched candidate
(timesteps + block-skip mask of GD/sample_imagenet64_classifier_guidance_dynamic_subnet.sh:13-14),
batch 256 per GPU, UNet-only (no classifier cond_fn: the classifier is row N1 of SURVEY §8f).
One bench "step" = one full K'=10-step sampling pass over one batch: 10 UNet forwards (9 full +
1 with 9 blocks skipped) + 10 fused DDIM updates + uint8 pack.

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
METRIC = "ADM-G 64x64 images/s, 10-step searched DDIM + block-skip mask"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--dump-ops", default=None, help="write every recorded op's kind/flops/bytes/ms of one step to this CSV")
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
def cpu_reference_rate(batch, n_ddim_steps, warm=True):
    """images/s of the CPU oracle on cand10, measured on `n_ddim_steps` consecutive schedule
    positions (starting at the first sampled step) at batch `batch`, scaled to the 10-step schedule."""
    import torch
    from oracle import diffusion_ref, unet_ref, weights

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = unet_ref.adm_g64_config()
    sd = weights.make_state_dict(unet_ref.param_shapes(cfg), seed=0)
    base = diffusion_ref.base_tables("cosine", 1000)
    tmap, nb = diffusion_ref.respace(base["alphas_cumprod"], CAND10["timesteps"])
    tb = diffusion_ref.diffusion_tables(nb)
    noise = torch.randn(batch, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    y = torch.randint(0, 1000, (batch,), generator=torch.Generator().manual_seed(3))
    unet = lambda x, t, yy, skip: unet_ref.unet_forward(sd, cfg, x, t, yy, skip)
    model_fn = diffusion_ref.make_model_fn(unet, tmap)
    K = len(tmap)

    def run_steps(positions):
        x = noise
        t0 = time.perf_counter()
        for i in positions:
            one = {k: v[i:i + 1] for k, v in tb.items()}
            # a 1-step schedule = step i of the 10-step one in isolation (same ops, same tables)
            x = diffusion_ref.ddim_sample_loop(
                lambda xx, ts, **kw: model_fn(xx, ts, **{**kw}), x.shape, one, [tmap[i]], x, True,
                model_kwargs={"y": y, "skip_layers": CAND10["skip_layers"]})
        return time.perf_counter() - t0

    order = list(range(K))[::-1]
    if warm:
        run_steps(order[:1])
    return run_steps, order, K


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    run_steps, order, K = cpu_reference_rate(batch, 1)
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
        "config": {"workload": "ADM-G 64x64 cand10 (10 searched steps + skip mask), UNet-only, CPU oracle port of the reference",
                   "batch": batch},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"one DDIM step (UNet fwd + update) per bench step at batch {batch}, rotating through the "
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

    from autodiffusion_b200 import create_model_and_diffusion, model_and_diffusion_defaults, ops
    from autodiffusion_b200.sampler import SchedulePlan, resolve_candidate

    d = model_and_diffusion_defaults()
    d.update(ADM_FLAGS)
    model, diffusion = create_model_and_diffusion(**d)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(bench_weights(shapes))
    model.to(dev).eval()
    model.convert_to_fp16()

    B = args.batch
    active, per_step = resolve_candidate(CAND10, diffusion)
    t_build = time.time()
    plan = SchedulePlan(model, active, per_step, B, clip_denoised=True, cond_fn=None, pack_uint8=True)
    torch.cuda.synchronize()
    t_build = time.time() - t_build
    t_rebuild = time.time()
    plan = SchedulePlan(model, active, per_step, B, clip_denoised=True, cond_fn=None, pack_uint8=True)  # cached masks
    torch.cuda.synchronize()
    t_rebuild = time.time() - t_rebuild
    launches_per_step = plan.launches

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise_dev = torch.randn(plan.shape, generator=g, device=dev)
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

    def step_resident():
        plan.run(noise_dev, y_dev)

    def step_e2e():
        plan.run(noise_host, y_host)  # H2D of x_T and y inside
        u8_host.copy_(plan.u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs the images (…progressive.py:427)

    ms, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)
    e2e_value = imgs / (ms_e2e * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "ADM-G 64x64 (295.9M params, random-init), published 10-step searched candidate "
                               "(timesteps + block-skip mask), UNet-only (no classifier cond_fn), one step = full 10-step "
                               "sampling of one batch", "batch_per_gpu": B, "ddim_steps": active.num_timesteps,
                   "l2": "activations per launch (>=400 MB at batch 256) exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world} (independent batches per rank, no data-path collective)"},
        "ms_per_unet_fwd": ms / args.steps / active.num_timesteps,
        "tflops_effective": value / world * GFLOP_PER_IMAGE / 1e3,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": noise_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": u8_host.numel()},
        "gpu_launches": launches_per_step * args.steps,
        "plan_build_s": {"first_candidate": round(t_build, 3), "next_candidate_same_masks": round(t_rebuild, 4)},
        "clocks": clocks,
    }

    if rank == 0 and not args.no_roofline:
        # per-kernel times: the same recorded ops, eager on the current stream with an event pair around each
        pk = peaks()
        # one sampled schedule = its K' cached UNet plans (+ the DDIM updates, timed as one event pair each)
        info, ms_ops = [], []
        for up in plan.steps:
            up.plan.run_profiled()  # warm
        for n, up in enumerate(plan.steps):
            plan.t_in.fill_(plan.t_values[n])
            info += up.plan.op_info()
            ms_ops += up.plan.run_profiled()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan._update(n)
            e1.record()
            torch.cuda.synchronize()
            info.append(("ddim_step", 0.0, 4.0 * 3 * plan.x.numel()))
            ms_ops.append(e0.elapsed_time(e1))
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
        line["roofline"] = {"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM: 3x3/1x1 conv, qkv/proj)", "bound": "tensor",
                            "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                            "traffic": None, "peak_source": pk["source"] + " (sustained bf16: kernel timed inside a long step)",
                            "launches": conv[0], "share_of_step": conv[1] / total_ms,
                            "flops_per_launch_avg": conv[2] / conv[0], "ms_per_launch_avg": conv[1] / conv[0]}
        line["kernel_breakdown"] = {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / total_ms, 4),
                                        "tflops": (v[2] / (v[1] * 1e-3) / 1e12) if v[2] and v[1] else None,
                                        "gbs": (v[3] / (v[1] * 1e-3) / 1e9) if v[3] and v[1] else None} for k, v in agg.items()}
    if rank == 0 and not args.no_cpu_baseline:
        import torch as _t

        run_steps, order, K = cpu_reference_rate(args.ref_batch, 1)
        tt = [run_steps([order[i]]) for i in (0, 3, 6)]  # includes position 6 (t=676), the masked step
        mean_pos = sum(tt) / len(tt)
        line["cpu_baseline"] = {"value": args.ref_batch / (mean_pos * K), "unit": "images/s", "cores": _t.get_num_threads(),
                                "kind": "port",
                                "sample": f"3 of the 10 DDIM steps of the same candidate (incl. the masked one) at batch "
                                          f"{args.ref_batch}, fp32 torch CPU oracle; images/s = batch / (10 x mean step time)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
