#!/bin/bash
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_multivariate.py -m gpu -x -q > gpurun_out/pytest_mgpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_mgpu_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench2_$TAG.json 2> gpurun_out/bench2_$TAG.err; echo "bench2 exit $?"; tail -2 gpurun_out/bench2_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench2_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["carry_exchange"], d["scaling"])
print({k: round(v*1000,1) for k,v in d["stage_ms"].items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --mode independent --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench2i_$TAG.json 2> gpurun_out/bench2i_$TAG.err; echo "bench2 independent exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench2i_$TAG.json"))
print("independent: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["scaling"])
PY
