#!/usr/bin/env bash
# round-2 call u: K1 backward fast path, full suite, smoke, kernel sweep (k1, conv0), bench (+ exhaustive cuDNN autotune A/B), cfg4
set -u
tag=${1:-r02u}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 10 400 python scripts/kernel_bench.py --quick --only k1 --json $out/${tag}_kernel_bench_k1.json 2>&1 | grep -E "\(16, 320, 128, 128\)|\(16, 192, 16, 16\)"
timeout -k 10 300 python scripts/kernel_bench.py --only conv0 --json $out/${tag}_kernel_bench_conv0.json 2>&1 | grep -E "fused"
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
timeout 600 python bench.py --cudnn-benchmark-limit 0 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu_cudnn_exhaustive.json 2> $out/${tag}_bench_1gpu_cudnn_exhaustive.err; echo "bench exhaustive rc=$?"; tail -2 $out/${tag}_bench_1gpu_cudnn_exhaustive.err
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg4_1gpu.json 2> $out/${tag}_bench_cfg4_1gpu.err; echo "cfg4 rc=$?"
python - <<PY
import json
for f in ("bench_1gpu","bench_1gpu_cudnn_exhaustive","bench_cfg4_1gpu"):
    try:
        d=json.load(open("$out/${tag}_%s.json" % f)); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"],1), d["roofline"]["frac"], {k:round(v["frac_of_hbm_peak"],3) for k,v in d["kernels"].items() if k.startswith(("k1_","conv0")) and "frac_of_hbm_peak" in v})
    except Exception as e: print(f, "unreadable", e)
PY
