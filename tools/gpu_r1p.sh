#!/bin/bash
TAG=${1:-r1p}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_$TAG.log
timeout 300 python tools/time_marginals.py chain 2>&1 | tail -6
