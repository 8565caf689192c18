#!/bin/bash
TAG=${1:-r1d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_$TAG.log
timeout 900 python tools/scan_latency.py --out gpurun_out/scan_latency_$TAG.json > gpurun_out/scan_latency_$TAG.log 2>&1; echo "scan exit $?"; tail -3 gpurun_out/scan_latency_$TAG.log | cut -c1-400
timeout 600 bash tools/c2_cli.sh gpurun_out/c2_cli_$TAG.json > gpurun_out/c2_cli_$TAG.log 2>&1; echo "c2 exit $?"; tail -4 gpurun_out/c2_cli_$TAG.log | cut -c1-900
