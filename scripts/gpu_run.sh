#!/usr/bin/env bash
# One gpurun call that re-measures the tree: full -m gpu suite, smoke(), the kernel sweeps, the bench line (and its A/B switches).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_run.sh r03a'        outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/
set -u
tag=${1:-run}
out=gpurun_out
mkdir -p $out
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 10 600 python scripts/kernel_bench.py --quick --json $out/${tag}_kernel_bench.json > $out/${tag}_kernel_bench.log 2>&1; grep -E "\(16, 320, 128, 128\)|\(16, 128, 256, 256\)|\(16, 128, 128, 128\)|\(8, 192" $out/${tag}_kernel_bench.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_1gpu.err
for sw in --no-overlap-hyper --no-fast-last-layer --no-fuse-first-layer; do
  timeout 400 python bench.py $sw --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_1gpu${sw}.json 2> $out/${tag}_bench_1gpu${sw}.err; echo "bench $sw rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_bench_1gpu*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, round(d["e2e"]["value"],1), round(d["roofline"]["frac"],4))
    except Exception as e: print(f, "unreadable", e)
PY
