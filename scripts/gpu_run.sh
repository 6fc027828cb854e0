#!/usr/bin/env bash
# round-2 call d: full GPU suite on the new tree, kernel sweep with the host-run-ahead timing, the new bench line, ncu of cdf_diff
set -u
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -15 $out/${tag}_pytest_gpu.log
timeout -k 10 400 python scripts/kernel_bench.py --quick --json $out/${tag}_kernel_bench.json > $out/${tag}_kernel_bench.log 2>&1
grep -E "\(16, 320, 128, 128\)|\(16, 128, 256, 256\)|\(16, 128, 32, 32\)|\(8, 192|\(16, 128, 128, 128\)" $out/${tag}_kernel_bench.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_1gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"], d["cpu_baseline"], d["gpu_eager_baseline"])
for k,v in d["kernels"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("shape","bytes")})
PY
timeout 300 python scripts/conv_probe.py --json $out/${tag}_conv_probe.json > $out/${tag}_conv_probe.log 2>&1; cat $out/${tag}_conv_probe.log
timeout 300 python scripts/ncu_target.py cdf,gdn,dense_bwd 1 > $out/${tag}_plain_targets.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bottleneck_fwd_kernel|gdn_bwd_nhwc|gdn_bwd_finalize|dgamma_kernel|gdn_dense_ws_kernel" -c 14 -f -o $out/${tag}_kernels python scripts/ncu_target.py cdf,gdn,dense_bwd 1 > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
