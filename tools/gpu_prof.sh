#!/bin/bash
# ncu --set full captures of single ops:  gpurun -- 'bash tools/gpu_prof.sh tag "fwd:0 bwd:0 bwd:2 build:0"'
set -u
TAG=${1:-r01}; shift
OUT=gpurun_out; mkdir -p $OUT
for spec in ${1:-fwd:0}; do
  op=${spec%%:*}; lvl=${spec##*:}
  timeout 300 python tools/prof_one.py $op --lvl $lvl --iters 3 > $OUT/${TAG}_${op}${lvl}_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lookup|build_pyramid|gaussian|lowmem|pack_fmaps" -s 2 -c 1 -f \
      -o $OUT/${TAG}_${op}${lvl} python tools/prof_one.py $op --lvl $lvl --iters 3 > $OUT/${TAG}_${op}${lvl}_ncu.log 2>&1
  echo "$spec -> $?"
done
