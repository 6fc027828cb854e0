#!/usr/bin/env bash
# quick iteration: selected GPU tests + smoke + one bench line without the CPU / eager comparison arms
set -u
tag=${1:-r02aq}
sel=${2:-tests}
out=gpurun_out
mkdir -p $out
timeout -k 10 600 python -m pytest $sel -m gpu -q -x > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -25 $out/${tag}_pytest_gpu.log | cut -c1-220
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_1gpu.json")); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4))
except Exception as e: print("unreadable", e)
PY
