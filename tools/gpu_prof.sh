#!/bin/bash
# ncu captures of one bench step: bash tools/gpu_prof.sh <tag> [kernel regex] [count]
TAG=${1:-p}
RX=${2:-"k_soap|k_dgemm|k_neigh"}
CNT=${3:-6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --named-configs none"
# launch list (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py $BARGS > $OUT/ncu_launches.log 2>&1
# full capture of the path's kernels in one timed step (after bootstrap + 3 warm-up steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 21 -c $CNT -f -o $OUT/prof \
    python bench.py $BARGS > $OUT/ncu_full.log 2>&1
ls -la $OUT
