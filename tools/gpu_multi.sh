#!/bin/bash
# Multi-GPU check on one box: bash tools/gpu_multi.sh <tag> <n_gpus> [extra bench args]
# bench at N ranks with the peer-memory reduction and (unless the 4th argument is nonccl) with NCCL only, plus the sharded-MD replica check.
TAG=${1:-m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
run bench.py --gpus $N --steps 20 --warmup 3 $3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench p2p rc=$?"
if [ "$4" != "nonccl" ]; then GAP_B200_P2P=0 run bench.py --gpus $N --steps 20 --warmup 3 --named-configs none > $OUT/bench_n${N}_nccl.json 2> $OUT/bench_n${N}_nccl.err; echo "bench nccl rc=$?"; fi
run tools/md_sharded_check.py 6 10 > $OUT/md_n$N.json 2> $OUT/md_n$N.err; echo "md rc=$?"
python tools/bench_summary.py $OUT/bench_n$N.json || tail -30 $OUT/bench_n$N.err
[ -f $OUT/bench_n${N}_nccl.json ] && python tools/bench_summary.py $OUT/bench_n${N}_nccl.json
tail -2 $OUT/md_n$N.json; tail -5 $OUT/md_n$N.err
