#!/bin/bash
# BASELINE.json configs[1] (array-CGH-shaped 1e7 probes, K = 5, auto priors) end to end through the command line:
# the reference binary (oracle/_ref/hammlet, CPU, 1 thread) and bin/hammlet (one B200) on the same text file, same flags.
# Wall-clock seconds of the whole process (parse + load + 1000 sweeps + marginals output) -> gpurun_out/c2_cli.json
set -u
OUT=${1:-gpurun_out/c2_cli.json}
W=$(mktemp -d)
python - "$W" <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from hammlet_b200.synth import piecewise_gaussian
x = piecewise_gaussian(10_000_000, 5, 5000, seed=2)
np.savetxt(sys.argv[1] + "/c2.txt", x, fmt="%.5f")
PY
ARGS="-f $W/c2.txt -a -R 2 -s 5 -i F 1000 10 -O M -w"
t0=$(date +%s.%N); hammlet_b200/bin/hammlet $ARGS -o $W/our- .csv > $W/our.log 2>&1; rc_o=$?; t1=$(date +%s.%N)
taskset -c 0 oracle/_ref/hammlet $ARGS -o $W/ref- .csv > $W/ref.log 2>&1; rc_r=$?; t2=$(date +%s.%N)
# load-only runs (1 sweep) so that the per-sweep rate can be separated from parsing
ARGS1="-f $W/c2.txt -a -R 2 -s 5 -i F 1 0 -O M -w"
t3=$(date +%s.%N); hammlet_b200/bin/hammlet $ARGS1 -o $W/our1- .csv > /dev/null 2>&1; t4=$(date +%s.%N)
taskset -c 0 oracle/_ref/hammlet $ARGS1 -o $W/ref1- .csv > /dev/null 2>&1; t5=$(date +%s.%N)
python - "$W" "$OUT" $rc_o $rc_r $t0 $t1 $t2 $t3 $t4 $t5 <<'PY'
import json, sys
W, out, rc_o, rc_r = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
t = [float(v) for v in sys.argv[5:]]
def lines(p):
    try: return sum(1 for _ in open(p))
    except OSError: return -1
d = {"config": "C2: 1e7 probes, K=5, auto priors, -i F 1000 10, marginals output; text input 1e7 lines",
     "ours_wall_s": t[1] - t[0], "reference_wall_s": t[2] - t[1], "ours_rc": rc_o, "reference_rc": rc_r,
     "ours_load_plus_1_sweep_s": t[4] - t[3], "reference_load_plus_1_sweep_s": t[5] - t[4],
     "ours_sweeps_per_s": 999 / max((t[1] - t[0]) - (t[4] - t[3]), 1e-9),
     "reference_sweeps_per_s": 999 / max((t[2] - t[1]) - (t[5] - t[4]), 1e-9),
     "ours_marginal_lines": lines(W + "/our-marginals.csv"), "reference_marginal_lines": lines(W + "/ref-marginals.csv"),
     "reference": "oracle/_ref/hammlet (g++ -O3, 1 thread pinned with taskset -c 0)"}
json.dump(d, open(out, "w"))
print(json.dumps(d))
PY
tail -2 $W/our.log
rm -rf $W
