#!/bin/bash
# One gpurun call for a round's 1-GPU record: the whole GPU parity suite, the default bench line (config A + named configs + parity block + CPU
# leg), the reference arm, the ncu launch list and one full capture of the path's kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt
nproc >> $OUT/gpu.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python tools/bench_summary.py $OUT/bench.json || tail -30 $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "reference rc=$?"
cat $OUT/bench_reference.json
bash tools/gpu_prof.sh $TAG
