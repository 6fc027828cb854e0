#!/usr/bin/env bash
set -u
tag=${1:-r02v}
out=gpurun_out
mkdir -p $out
timeout 300 python scripts/ncu_target.py conv0,k1 1 > $out/${tag}_plain_targets.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv0_gdn_kernel|bottleneck_bwd_kernel" -c 6 -f -o $out/${tag}_conv0_k1bwd python scripts/ncu_target.py conv0,k1 1 > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $out/${tag}_ncu.log
CHANNELS_LAST=1 CUDNN_BENCHMARK=1 ROWS=80 timeout 300 python scripts/profile_step.py ours > $out/${tag}_step_breakdown_torchprofiler.txt 2>&1; echo "profile rc=$?"; head -3 $out/${tag}_step_breakdown_torchprofiler.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-eager-baseline > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_shares.py $out/${tag}_launches.csv 2 > $out/${tag}_launches.txt 2>&1; head -30 $out/${tag}_launches.txt; tail -1 $out/${tag}_launches.txt
