#!/bin/bash
TAG=${1:-r1f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_$TAG.log
timeout 900 python tools/c3_chromosomes.py --streams 4 --out gpurun_out/c3_$TAG.json > gpurun_out/c3_$TAG.log 2>&1; echo "c3 exit $?"; tail -3 gpurun_out/c3_$TAG.log | cut -c1-1500
timeout 600 python tools/c3_chromosomes.py --streams 8 --out gpurun_out/c3s8_$TAG.json > gpurun_out/c3s8_$TAG.log 2>&1; echo "c3 exit $?"; tail -1 gpurun_out/c3s8_$TAG.log | cut -c1-700
