#!/bin/bash
# bench at N ranks with the peer-reduction variants side by side; usage: bash tools/gpu_multi2.sh <tag> <N> [named-configs for the first run]
TAG=${1:-m2}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
GAP_B200_P2P_LL_MIN_RANKS=2 run bench.py --gpus $N --steps 20 --warmup 3 --named-configs ${3:-none} > $OUT/bench_n${N}_ll.json 2> $OUT/bench_n${N}_ll.err; echo "ll rc=$?"

GAP_B200_P2P_LL_MAX_DOUBLES=0 run bench.py --gpus $N --steps 20 --warmup 3 --named-configs none > $OUT/bench_n${N}_one.json 2> $OUT/bench_n${N}_one.err; echo "one-shot rc=$?"
GAP_B200_P2P_LL_MIN_RANKS=2 run tools/md_sharded_check.py 6 10 > $OUT/md_n$N.json 2> $OUT/md_n$N.err; echo "md rc=$?"
for v in ll one; do python tools/bench_summary.py $OUT/bench_n${N}_$v.json | head -3 | cut -c1-330 || tail -30 $OUT/bench_n${N}_$v.err; done
tail -1 $OUT/md_n$N.json | cut -c1-400; tail -3 $OUT/md_n$N.err
