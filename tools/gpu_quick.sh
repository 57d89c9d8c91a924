#!/bin/bash
# Quick GPU check: parity tests + one bench line (no profiler).  usage: bash tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then
  python -m pytest tests -m gpu -x -q -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
else
  python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
fi
tail -15 $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j = json.load(open("$OUT/bench.json"))
    print("value %.0f atoms/s  ms/step %.4f  e2e %.0f" % (j["value"], j["ms_per_step"], j["e2e"]["value"]))
    print(j["roofline"]["stage_ms"], "frac", j["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench.err").read()[-2000:])
PY
