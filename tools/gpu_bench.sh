#!/bin/bash
# bench.py on the GPU box (+ optional reference arm). Usage: gpurun [--gpus N] --timeout 1200 -- 'bash tools/gpu_bench.sh <tag> <N> [extra bench args]'
set -u
TAG=$1; N=${2:-1}; shift; shift || true
OUT=gpurun_out; mkdir -p $OUT
if [ "$N" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
else
  RUN="python bench.py --gpus 1"
fi
timeout 900 $RUN "$@" > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench exit $?"; tail -c 6000 $OUT/${TAG}_bench_n$N.json; tail -15 $OUT/${TAG}_bench_n$N.err
