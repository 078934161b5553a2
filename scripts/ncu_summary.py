#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of columns the roofline discussion uses.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_<what>_summary.csv [--traffic-json profiles/conv_igemm_traffic.json]

Runs `ncu -i <rep> --page raw --csv` (works on the CPU box) and keeps, per profiled launch: kernel, duration, grid,
registers, tensor-pipe / XU (MUFU) / issue utilisation, DRAM bytes read + written, DRAM and L2 throughput, L2 hit rate,
resident warps. With --traffic-json it also (re)writes the per-launch DRAM traffic of the conv_igemm launches in the
report - the `roofline.traffic` figure bench.py prints - so that number always comes from a committed capture of the
current build."""
import csv
import io
import json
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    conv = []
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{c} [{units[idx[c]]}]" for c in COLS if c in idx])
        for r in data:
            name = r[idx["Kernel Name"]]
            for junk in ("adb::", "<unnamed>::", "(anonymous namespace)::", "unnamed>::", "void "):
                name = name.replace(junk, "")
            w.writerow([name[:60]] + [r[idx[c]] for c in COLS if c in idx])
            if "conv_igemm" in name:
                b = 0.0
                for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    b += float(r[idx[c]]) * UNIT.get(units[idx[c]], 1.0)
                tu = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
                conv.append((name[:40], b, float(r[idx["gpu__time_duration.sum"]]) * tu.get(units[idx["gpu__time_duration.sum"]], 1.0)))
    print(f"{len(data)} launches -> {out}")
    if traffic and conv:
        json.dump({"dram_bytes_per_launch_avg": sum(b for _, b, _ in conv) / len(conv), "launches_sampled": len(conv),
                   "per_launch": [{"kernel": k, "dram_bytes": b, "us": t} for k, b, t in conv],
                   "what": "dram__bytes_read.sum + dram__bytes_write.sum per conv_igemm launch of one `ncu --set full` capture",
                   "source": out}, open(traffic, "w"), indent=1)
        print(f"conv_igemm traffic of {len(conv)} launches -> {traffic}")


if __name__ == "__main__":
    main()
