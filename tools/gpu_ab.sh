#!/bin/bash
# A/B of an environment switch on the headline bench: bash tools/gpu_ab.sh <tag> <ENVVAR> [pytest -k expr]
TAG=${1:-ab}
VAR=${2:-GAP_B200_PDL}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q ${3:+-k "$3"} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for rep in 1 2; do
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --named-configs none > $OUT/bench_on_$rep.json 2> $OUT/bench_on_$rep.err; echo "on rc=$?"
  env $VAR=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --named-configs none > $OUT/bench_off_$rep.json 2> $OUT/bench_off_$rep.err; echo "off rc=$?"
done
python tools/bench_summary.py $OUT/bench_on_1.json $OUT/bench_off_1.json $OUT/bench_on_2.json $OUT/bench_off_2.json | grep -v "parity\|clocks"
