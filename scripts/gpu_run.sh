#!/usr/bin/env bash
# round-2 call p: hyper branch on a side stream (A/B), wave-fill grid of the NHWC GDN backward, full suite
set -u
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -8 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout -k 10 400 python scripts/kernel_bench.py --quick --only gdn --json $out/${tag}_kernel_bench_gdn.json > $out/${tag}_kernel_bench_gdn.log 2>&1
grep -E "bwd nhwc" $out/${tag}_kernel_bench_gdn.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
timeout 400 python bench.py --no-overlap-hyper --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu_no_overlap.json 2> $out/${tag}_bench_1gpu_no_overlap.err; echo "bench no-overlap rc=$?"; tail -3 $out/${tag}_bench_1gpu_no_overlap.err
python - <<PY
import json
for f in ("bench_1gpu","bench_1gpu_no_overlap"):
    try:
        d=json.load(open("$out/${tag}_%s.json" % f)); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"],1), d["roofline"]["frac"], d["roofline"].get("site_backward_ms_incl_fold_launch"), d["roofline"].get("traffic"), d["config"].get("hyper_branch_on_side_stream"))
    except Exception as e: print(f, "unreadable", e)
PY
