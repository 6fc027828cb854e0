#!/usr/bin/env bash
# usage: gpu_multi.sh <tag> <N>: cfg2 and cfg4 training step + cfg3 codec at N GPUs (one process per GPU, NCCL), same launch line as the driver's
set -u
tag=$1; N=$2
out=gpurun_out
mkdir -p $out
run() { if [ "$N" -gt 1 ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; else python "$@"; fi; }
run bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg2_${N}gpu.json 2> $out/${tag}_bench_cfg2_${N}gpu.err; echo "cfg2 rc=$?"
run bench.py --gpus $N --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg4_${N}gpu.json 2> $out/${tag}_bench_cfg4_${N}gpu.err; echo "cfg4 rc=$?"; tail -2 $out/${tag}_bench_cfg4_${N}gpu.err
run scripts/codec_bench.py 16 > $out/${tag}_codec_cfg3_${N}gpu.json 2> $out/${tag}_codec_cfg3_${N}gpu.err; echo "codec rc=$?"; tail -2 $out/${tag}_codec_cfg3_${N}gpu.err
python - <<PY
import json
for c in ("cfg2","cfg4"):
    try:
        d=json.load(open("$out/${tag}_bench_%s_${N}gpu.json" % c))
        print(c, "N=$N", {k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"].get("gradient_buckets"))
    except Exception as e: print(c, "unreadable", e)
try:
    d=json.load(open("$out/${tag}_codec_cfg3_${N}gpu.json")); print("codec", {k:v for k,v in d.items() if not isinstance(v,(list,dict))})
except Exception as e: print("codec unreadable", e)
PY
