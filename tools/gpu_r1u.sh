#!/bin/bash
TAG=${1:-r1u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_$TAG.log
timeout 200 python tools/time_marginals.py chain 2>&1 | tail -5
