#!/bin/bash
# ncu --set full captures of the training-path kernels:  gpurun -- 'bash tools/gpu_prof2.sh tag'
set -u
TAG=${1:-r01j}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/prof_one.py bbwd --iters 2 > $OUT/${TAG}_bbwd_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:build_bwd_kernel" -s 5 -c 2 -f \
    -o $OUT/${TAG}_bbwd python tools/prof_one.py bbwd --iters 2 > $OUT/${TAG}_bbwd_ncu.log 2>&1
echo "bbwd -> $?"
timeout 300 python tools/prof_one.py fusedbwd_acc --iters 3 > $OUT/${TAG}_fusedbwdacc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lookup_fused_bwd" -s 2 -c 1 -f \
    -o $OUT/${TAG}_fusedbwdacc python tools/prof_one.py fusedbwd_acc --iters 3 > $OUT/${TAG}_fusedbwdacc_ncu.log 2>&1
echo "fusedbwd_acc -> $?"
