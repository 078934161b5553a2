#!/bin/bash
# Population evaluation with the native Inception-V3 pool_3 extractor: throughput, the extractor alone, and the
# ncu launch list of a short run (no cuDNN / cuBLAS / ATen convolution kernel may appear in it).
mkdir -p gpurun_out
timeout 600 python scripts/population_eval.py --candidates 3 --num_samples 1000 --batch_size 250 --guided --features inception > gpurun_out/pop_n1_inception_r2.json 2> gpurun_out/pop_n1_inception_r2.err
echo "population with inception features rc=$?"; cut -c1-400 gpurun_out/pop_n1_inception_r2.json; tail -2 gpurun_out/pop_n1_inception_r2.err
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/inception_rate_r2.txt
import torch, time
from autodiffusion_b200.inception import InceptionPool3
m = InceptionPool3().cuda()
u8 = torch.randint(0, 256, (250, 64, 64, 3), dtype=torch.uint8, device="cuda")
for _ in range(2): f = m(u8)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(4): f = m(u8)
torch.cuda.synchronize(); dt = (time.time() - t0) / 4
print(f"InceptionPool3 (native kernels), 250 uint8 64x64 images -> pool_3 [250, 2048]: {dt * 1e3:.1f} ms = {250 / dt:.0f} images/s; features finite: {bool(torch.isfinite(f).all())}")
PY
timeout 300 python scripts/population_eval.py --candidates 1 --num_samples 32 --batch_size 32 --small --guided --features inception > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_inception_r2.csv python scripts/population_eval.py --candidates 1 --num_samples 32 --batch_size 32 --small --guided --features inception > gpurun_out/ncu_inception.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/launches_inception_r2.csv") if l.startswith('"')))
kn = rows[0].index("Kernel Name")
c = collections.Counter(re.sub(r"\(.*", "", r[kn]).replace("void ", "") for r in rows[1:] if len(r) > kn)
bad = [k for k in c if re.search(r"cudnn|cutlass|cublas|gemm|sm90|sm80|ampere|implicit_convolve|conv2d", k, re.I) and "conv_igemm" not in k]
print(len(rows) - 1, "launches,", len(c), "distinct kernels; library convolution / GEMM kernels:", bad or "none")
for k, v in c.most_common(40): print(f"{v:6d}  {k[:110]}")
PY
