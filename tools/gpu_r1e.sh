#!/bin/bash
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_$TAG.log
timeout 900 python tools/scan_latency.py --only "C5,C2" --out gpurun_out/scan_latency_$TAG.json > gpurun_out/scan_latency_$TAG.log 2>&1; echo "scan exit $?"; tail -3 gpurun_out/scan_latency_$TAG.log | cut -c1-1200
