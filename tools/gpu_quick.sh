#!/bin/bash
# Quick GPU check: parity tests + one bench line (no profiler).  usage: bash tools/gpu_quick.sh <tag> [pytest -k expr] [extra bench args]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then
  python -m pytest tests -m gpu -x -q -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
else
  python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
fi
tail -15 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 3 $3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python tools/bench_summary.py $OUT/bench.json || tail -30 $OUT/bench.err
