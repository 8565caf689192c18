#!/bin/bash
TAG=${1:-r1q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_$TAG.log
timeout 500 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "rec", d["recorded"]["value"], "launches", d["gpu_launches"])
print({k: round(x*1000,1) for k,x in d["stage_ms"].items()})
PY
timeout 300 python tools/scan_latency.py --only "C1,C2,C4/8" --out gpurun_out/scan_$TAG.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], round(d['sweep_ms_device']*1000,1), 'us', round(d['sweeps_per_s']))
"
