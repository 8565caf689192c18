#!/bin/bash
# final capture of the round: both bench arms, ncu launch list, ncu --set full of one whole sweep
TAG=${1:-r1z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.out 2>&1; echo "ncu list exit $?"
timeout 700 ncu --set full --clock-control none --import-source on \
  -k regex:"k_cand_|k_block_emit|k_fwd_|k_bwd_|k_reduce_" -s 56 -c 13 -f -o gpurun_out/${TAG}_full \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.out 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "rec", d["recorded"]["value"])
print({k: round(v*1000,1) for k,v in d["stage_ms"].items()})
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print(d.get("cpu_baseline"))
PY
head -c 400 gpurun_out/bench_ref_$TAG.json
