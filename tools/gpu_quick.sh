#!/bin/bash
# Quick GPU-box visit: the GPU parity suite + smoke.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh [tag] [pytest args]'
set -u
TAG=${1:-r02}; shift || true
OUT=gpurun_out; mkdir -p $OUT
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(tests -x)
timeout 800 python -m pytest -m gpu -q "${ARGS[@]}" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -40 $OUT/${TAG}_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -5 $OUT/${TAG}_smoke.log
