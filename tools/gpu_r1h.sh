#!/bin/bash
TAG=${1:-r1h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_$TAG.log
timeout 500 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "recorded", d["recorded"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_chunks_wide|k_block_emit|k_bwd_maps|k_fwd_replay" -s 8 -c 4 -f -o gpurun_out/${TAG}_c5_full \
  python tools/scan_latency.py --only C5 --out gpurun_out/scan_c5_ncu_$TAG.json > gpurun_out/${TAG}_c5_full.out 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/${TAG}_c5_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_c5_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_c5_full.ncu-rep --page details --csv > gpurun_out/${TAG}_c5_full_details.csv 2>/dev/null
ls -la gpurun_out | grep $TAG
