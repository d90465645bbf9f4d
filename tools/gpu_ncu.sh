#!/bin/bash
# ncu --set full capture of one kernel of a diag script.  Usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <script.py> [args]
set -u
TAG=$1; K=$2; SKIP=$3; shift; shift; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python "$@" > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c 1 -f -o $OUT/$TAG python "$@" > $OUT/${TAG}_ncu.log 2>&1
echo "$TAG -> $?"; tail -3 $OUT/${TAG}_plain.log
