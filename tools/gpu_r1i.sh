#!/bin/bash
TAG=${1:-r1i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_$TAG.log
for v in 0 1; do
HML_EMIT_SHARE=$v timeout 500 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_share${v}_$TAG.json 2> gpurun_out/bench_share${v}_$TAG.err; echo "bench share=$v exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_share${v}_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print({k: round(x*1000,1) for k,x in d["stage_ms"].items()})
PY
done
timeout 900 python tools/scan_latency.py --only "C5,C1" --out gpurun_out/scan_latency_$TAG.json > gpurun_out/scan_latency_$TAG.log 2>&1; echo "scan exit $?"; tail -2 gpurun_out/scan_latency_$TAG.log | cut -c1-1100
