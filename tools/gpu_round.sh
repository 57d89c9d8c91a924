#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one full capture of the path's kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt
nproc >> $OUT/gpu.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
cat $OUT/bench_reference.json
# launch list (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
# full capture of the path's kernels in one timed step (after bootstrap + 3 warm-up steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_soap|k_dgemm|k_neigh" -s 21 -c 6 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
