#!/bin/bash
TAG=${1:-r1l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_$TAG.log
timeout 500 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "rec", d["recorded"])
PY
timeout 600 python tools/c3_chromosomes.py --streams 8 --out gpurun_out/c3_$TAG.json 2>&1 | tail -1 | cut -c1-700
