#!/usr/bin/env bash
# usage: gpu_buckets.sh <tag> <N>: cfg2 step at N GPUs for several gradient-bucket sizes / NCCL CTA limits (how much of the all-reduce is hidden?)
set -u
tag=$1; N=$2
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for mb in 64 13 3; do
  run bench.py --gpus $N --steps 20 --warmup 5 --bucket-mb $mb --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg2_${N}gpu_bucket${mb}.json 2> $out/${tag}_bench_cfg2_${N}gpu_bucket${mb}.err; echo "bucket $mb rc=$?"
done
NCCL_MAX_CTAS=4 run bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg2_${N}gpu_bucket6_maxctas4.json 2> $out/${tag}_bench_cfg2_${N}gpu_bucket6_maxctas4.err; echo "maxctas4 rc=$?"
NCCL_MAX_CTAS=4 run bench.py --gpus $N --steps 20 --warmup 5 --bucket-mb 64 --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_bench_cfg2_${N}gpu_bucket64_maxctas4.json 2> $out/${tag}_bench_cfg2_${N}gpu_bucket64_maxctas4.err; echo "maxctas4 b64 rc=$?"
run scripts/codec_bench.py 16 > $out/${tag}_codec_cfg3_${N}gpu.json 2> $out/${tag}_codec_cfg3_${N}gpu.err; echo "codec rc=$?"; tail -2 $out/${tag}_codec_cfg3_${N}gpu.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_bench_cfg2_${N}gpu_*.json")):
    try:
        d=json.load(open(f)); print(f.split("gpu_")[-1], {k:d[k] for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), d["config"].get("gradient_buckets"))
    except Exception as e: print(f, "unreadable", e)
d=json.load(open("$out/${tag}_codec_cfg3_${N}gpu.json")); print("codec sharded api", d.get("sharded_api_with_host_gather"), d.get("gpu"))
PY
