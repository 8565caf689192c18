#!/bin/bash
# usage: run_test_bench.sh TAG  -> pytest gpu + bench, summary printed
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print({k: round(v*1000,1) for k,v in d["stage_ms"].items()})
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
tail -3 gpurun_out/bench_$TAG.err
