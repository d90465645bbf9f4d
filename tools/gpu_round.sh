#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + CPU arm), ncu launch list of the bench command.
# Usage (from the dev container):  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
BRC=$?
echo "bench exit $BRC"; tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "ref exit $?"; tail -c 1500 $OUT/${TAG}_bench_ref.json
if [ $BRC -eq 0 ]; then
  echo "== ncu launch list"
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > $OUT/${TAG}_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > $OUT/${TAG}_ncu.log 2>&1
  echo "ncu exit $?"
fi
