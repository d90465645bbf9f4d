#!/bin/bash
# A/B timing of library variants: bash tools/gpu_exp.sh <script.py> <lib1.so> [lib2.so ...]
set -u
S=$1; shift
for L in "$@"; do LGU_CORR_LIB=$PWD/$L timeout 300 python $S 2>&1 | tail -4; done
