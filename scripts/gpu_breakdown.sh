#!/usr/bin/env bash
# step breakdown (torch profiler) + ncu launch list of the bench command on the current tree
set -u
tag=${1:-r02ap}
out=gpurun_out
mkdir -p $out
CHANNELS_LAST=1 CUDNN_BENCHMARK=1 ROWS=140 timeout 300 python scripts/profile_step.py ours > $out/${tag}_step_breakdown_torchprofiler.txt 2>&1; echo "profile rc=$?"; head -3 $out/${tag}_step_breakdown_torchprofiler.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_shares.py $out/${tag}_launches.csv 2 > $out/${tag}_launches.txt 2>&1; head -40 $out/${tag}_launches.txt; tail -1 $out/${tag}_launches.txt
