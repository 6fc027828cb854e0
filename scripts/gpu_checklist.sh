#!/usr/bin/env bash
# One gpurun call that measures everything changed since the last GPU run of round 1 (profiles/README.md, "Changes made after
# the last GPU run").  Usage (from the repo root, ~3 GPU-minutes):
#   gpurun --timeout 600 -- 'bash scripts/gpu_checklist.sh r02'
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
# 1. parity first: the default suite, then the opt-in dense backward (SIC_EXPERIMENTAL=1)
timeout -k 10 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
SIC_EXPERIMENTAL=1 timeout -k 10 200 python -m pytest tests/test_gpu_gdn.py -m gpu -q -k fused_backward > $out/${tag}_pytest_dense_bwd.log 2>&1
echo "dense bwd rc=$?" | tee -a $out/${tag}_pytest_dense_bwd.log; tail -3 $out/${tag}_pytest_dense_bwd.log
# 2. kernels against the roofline (K1 sweep top, GDN NCHW/NHWC fwd+bwd, dense fwd both variants + C=192)
timeout -k 10 300 python scripts/kernel_bench.py --quick --json $out/${tag}_kernel_bench.json > $out/${tag}_kernel_bench.log 2>&1
grep -E "k1_bwd|gdn_.*nhwc|dense" $out/${tag}_kernel_bench.log | grep -E "\(16, 320, 128, 128\)|\(16, 128, 256, 256\)|\(16, 128, 128, 128\)|\(8, 192" 
SIC_DENSE_BWD=1 timeout -k 10 200 python scripts/kernel_bench.py --quick --only dense --json $out/${tag}_kernel_bench_dense_bwd_fused.json 2>&1 | grep -E "dense_bwd|FAILED"
# 3. the headline line
timeout 300 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench_1gpu.json
for pad in 4 8; do timeout 300 python bench.py --pad-rgb $pad --no-cpu-baseline > $out/${tag}_bench_1gpu_padrgb$pad.json 2> $out/${tag}_bench_1gpu_padrgb$pad.err; python -c "import json,sys; d=json.load(open('$out/${tag}_bench_1gpu_padrgb$pad.json')); print('pad-rgb $pad:', d['ms_per_step'], 'ms/step', d['value'], 'patches/s')"; done
# 4. one ncu --set full capture of the changed kernels (never a bench number)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gdn_(fwd|bwd)_nhwc|bottleneck_bwd" -c 4 -f \
    -o $out/${tag}_kernels python scripts/ncu_target.py all 1 > $out/${tag}_ncu_kernels.log 2>&1; echo "ncu rc=$?"
