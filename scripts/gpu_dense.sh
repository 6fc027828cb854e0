#!/usr/bin/env bash
set -u
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
timeout 300 python scripts/dgamma_bias_probe.py 2>&1 | tail -5
timeout -k 10 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log | cut -c1-200
timeout -k 10 300 python scripts/kernel_bench.py --quick --only dense --json $out/${tag}_kernel_bench_dense.json 2>&1 | grep -E "bwd|pipelined"
