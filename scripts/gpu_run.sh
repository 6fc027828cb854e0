#!/usr/bin/env bash
# round-2 call g: fused first layer after the accumulator fix, full suite, graph-timed kernel sweep, bench with the fused layer on / off
set -u
tag=${1:-r02g}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -8 $out/${tag}_pytest_gpu.log
timeout -k 10 300 python scripts/kernel_bench.py --only conv0 --json $out/${tag}_kernel_bench_conv0.json 2>&1 | tee $out/${tag}_kernel_bench_conv0.log | tail -12
timeout -k 10 400 python scripts/kernel_bench.py --quick --only gdn --json $out/${tag}_kernel_bench_gdn.json > $out/${tag}_kernel_bench_gdn.log 2>&1
grep -E "\(16, 128, 256, 256\)|\(16, 128, 32, 32\)|\(16, 128, 128, 128\)|\(16, 128, 64, 64\)" $out/${tag}_kernel_bench_gdn.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
timeout 400 python bench.py --no-fuse-first-layer --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu_unfused.json 2> $out/${tag}_bench_1gpu_unfused.err; echo "bench unfused rc=$?"
python - <<PY
import json
for f in ("$out/${tag}_bench_1gpu.json", "$out/${tag}_bench_1gpu_unfused.json"):
    d=json.load(open(f))
    print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("fused_first_layer"))
PY
