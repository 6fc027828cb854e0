#!/usr/bin/env bash
# where does the multi-GPU step overhead come from?  cfg2 at N GPUs: default / no collective at all / one bucket / no side stream / fewer NCCL CTAs
set -u
tag=$1; N=$2
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
B="bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline"
run $B > $out/${tag}_${N}gpu_default.json 2> $out/${tag}_${N}gpu_default.err; echo "default rc=$?"
SIC_DIAG_NO_ALLREDUCE=1 run $B > $out/${tag}_${N}gpu_noallreduce.json 2> $out/${tag}_${N}gpu_noallreduce.err; echo "noallreduce rc=$?"
run $B --bucket-mb 64 > $out/${tag}_${N}gpu_bucket64.json 2> $out/${tag}_${N}gpu_bucket64.err; echo "bucket64 rc=$?"
run $B --no-overlap-hyper > $out/${tag}_${N}gpu_nooverlap.json 2> $out/${tag}_${N}gpu_nooverlap.err; echo "nooverlap rc=$?"
NCCL_MAX_CTAS=8 run $B > $out/${tag}_${N}gpu_maxctas8.json 2> $out/${tag}_${N}gpu_maxctas8.err; echo "maxctas8 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_${N}gpu_*.json")):
    try:
        d=json.load(open(f)); print(f.split("gpu_")[-1], {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), d["clocks"])
    except Exception as e: print(f, "unreadable", e)
PY
