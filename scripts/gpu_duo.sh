#!/usr/bin/env bash
# Diagnosis of the 1 -> N step-time difference: the SAME single-GPU bench (no torch.distributed, no NCCL) alone on GPU 0, then as two
# independent processes on GPUs 0 and 1 whose timed regions start at the same wall-clock time.  If the pair is slower than the solo run,
# the difference is the box (power / host), not the data-parallel code.
set -u
tag=${1:-r02be}
out=gpurun_out
mkdir -p $out
B="bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline"
CUDA_VISIBLE_DEVICES=0 python $B > $out/${tag}_solo_gpu0.json 2> $out/${tag}_solo_gpu0.err; echo "solo rc=$?"
start=$(python -c "import time; print(time.time() + 45)")
CUDA_VISIBLE_DEVICES=0 SIC_BENCH_START_AT=$start python $B > $out/${tag}_duo_gpu0.json 2> $out/${tag}_duo_gpu0.err &
p0=$!
CUDA_VISIBLE_DEVICES=1 SIC_BENCH_START_AT=$start python $B > $out/${tag}_duo_gpu1.json 2> $out/${tag}_duo_gpu1.err &
p1=$!
wait $p0; echo "duo0 rc=$?"; wait $p1; echo "duo1 rc=$?"
python - <<PY
import json
for f in ("solo_gpu0","duo_gpu0","duo_gpu1"):
    try:
        d=json.load(open("$out/${tag}_%s.json" % f)); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"])
    except Exception as e: print(f, "unreadable", e)
PY
