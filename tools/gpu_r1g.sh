#!/bin/bash
# 2 GPUs: segment-split parity (peer + nccl transports) and both multi-GPU bench modes, short
TAG=${1:-r1g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_mgpu_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench2_$TAG.json 2> gpurun_out/bench2_$TAG.err; echo "bench2 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench2_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["carry_exchange"], d["scaling"])
print({k: round(v*1000,1) for k,v in d["stage_ms"].items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/c3_chromosomes.py --streams 6 --out gpurun_out/c3_2gpu_$TAG.json > gpurun_out/c3_2gpu_$TAG.log 2>&1; echo "c3 exit $?"; tail -1 gpurun_out/c3_2gpu_$TAG.log | cut -c1-600
