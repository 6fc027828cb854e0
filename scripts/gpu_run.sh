#!/usr/bin/env bash
# round-2 call l: full suite + bench on the tree with the faster hyper tail (in-kernel GDN fold reverted)
set -u
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -5 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_1gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["gpu_eager_baseline"], d["cpu_baseline"]["value"])
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
