#!/usr/bin/env bash
set -u
tag=${1:-r02w}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
timeout 400 python bench.py --no-overlap-hyper --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu_no_overlap.json 2> $out/${tag}_bench_1gpu_no_overlap.err; echo "bench no-overlap rc=$?"
python - <<PY
import json
for f in ("bench_1gpu","bench_1gpu_no_overlap"):
    try:
        d=json.load(open("$out/${tag}_%s.json" % f)); print(f, {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4))
    except Exception as e: print(f, "unreadable", e)
PY
