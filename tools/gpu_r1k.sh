#!/bin/bash
TAG=${1:-r1k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_$TAG.log
for v in 0 1; do
HML_PDL=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$v bench.py --gpus 2 --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench2_pdl${v}_$TAG.json 2> gpurun_out/bench2_pdl${v}_$TAG.err; echo "bench2 pdl=$v exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench2_pdl${v}_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["carry_exchange"])
PY
done
timeout 300 python tools/scan_latency.py --only "C5" --out gpurun_out/scan_$TAG.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], round(d['sweep_ms_device']*1000,1), 'us', round(d['sweeps_per_s'],1), {k: round(v) for k,v in d['stage_us'].items()})
"
