#!/usr/bin/env bash
# round-2 call j: last layer as 1x1 conv + position-major gathers; full suite; ncu --set full of the fused first-layer kernels
set -u
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -12 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout -k 10 300 python scripts/kernel_bench.py --only deconv --json $out/${tag}_kernel_bench_deconv.json 2>&1 | tee $out/${tag}_kernel_bench_deconv.log | tail -12
timeout 400 python bench.py --no-cpu-baseline > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
python - <<PY
import json
for f in ("$out/${tag}_bench_1gpu.json",):
    try:
        d=json.load(open(f))
        print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["config"].get("fused_first_layer"), d["config"].get("gemm_last_layer"))
    except Exception as e: print(f, "unreadable", e)
PY
timeout 300 python scripts/ncu_target.py conv0,deconv 1 > $out/${tag}_plain_targets.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv0_gdn_kernel|deconv_rgb" -c 8 -f -o $out/${tag}_conv0 python scripts/ncu_target.py conv0,deconv 2 > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
