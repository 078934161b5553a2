#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/prof_inception_once.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inception_only_r2.csv python scripts/prof_inception_once.py > gpurun_out/ncu_inception2.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/launches_inception_only_r2.csv") if l.startswith('"')))
h = rows[0]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mu = h.index("Metric Unit")
c = collections.Counter(); t = collections.Counter()
for r in rows[1:]:
    if len(r) <= mv: continue
    k = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1.0)
    c[k] += 1; t[k] += v
bad = [k for k in c if re.search(r"cudnn|cutlass|cublas|gemm|sm90|sm80|ampere|implicit_convolve|conv2d|xmma", k, re.I) and "conv_igemm" not in k]
print(len(rows) - 1, "launches,", len(c), "distinct kernels; library convolution / GEMM kernels:", bad or "none")
for k, v in t.most_common(25): print(f"{c[k]:6d} {v / 1e3:9.3f} ms  {k[:100]}")
PY
