#!/bin/bash
# BASELINE.json configs[1] (array-CGH-shaped 1e7 probes, K = 5, auto priors) end to end through the command line:
# the reference binary (oracle/_ref/hammlet, CPU, 1 thread) and bin/hammlet (one B200) on the same input, same flags.
# Wall-clock seconds of the whole process (parse + load + 1000 sweeps + marginals output) for text, gzip'd text and raw
# float32 input, and the sweep rate of the 1000 sweeps alone: ours from the binary's own -timing line (steady clock
# around sampleHMM), the reference's by differencing two runs that differ only in the number of sweeps, each preceded by
# an untimed run so that file cache and CUDA context creation are warm.  -> gpurun_out/c2_cli.json
set -u
OUT=${1:-gpurun_out/c2_cli.json}
W=$(mktemp -d)
python - "$W" <<'PY'
import gzip, sys, numpy as np
sys.path.insert(0, ".")
from hammlet_b200.synth import piecewise_gaussian
x = piecewise_gaussian(10_000_000, 5, 5000, seed=2)
np.savetxt(sys.argv[1] + "/c2.txt", x, fmt="%.5f")
open(sys.argv[1] + "/c2.txt.gz", "wb").write(gzip.compress(open(sys.argv[1] + "/c2.txt", "rb").read(), compresslevel=1))
np.loadtxt(sys.argv[1] + "/c2.txt", dtype=np.float32).tofile(sys.argv[1] + "/c2.f32")
PY
COMMON="-a -R 2 -s 5 -O M -w"
now() { date +%s.%N; }
hammlet_b200/bin/hammlet -f $W/c2.txt $COMMON -i F 1 0 -o $W/warm- .csv > /dev/null 2>&1     # warm-up: CUDA context, file cache
declare -A WALL
for fmt in txt gz f32; do
  case $fmt in txt) SRC="-f $W/c2.txt";; gz) SRC="-f $W/c2.txt.gz";; f32) SRC="-f $W/c2.f32 -F f32";; esac
  t0=$(now); hammlet_b200/bin/hammlet $SRC $COMMON -i F 1000 10 -timing -o $W/our-$fmt- .csv > $W/our-$fmt.log 2> $W/our-$fmt.err; rc=$?; t1=$(now)
  WALL[$fmt]="$t0 $t1 $rc"
done
taskset -c 0 oracle/_ref/hammlet -f $W/c2.txt $COMMON -i F 1 0 -o $W/warm- .csv > /dev/null 2>&1
t2=$(now); taskset -c 0 oracle/_ref/hammlet -f $W/c2.txt $COMMON -i F 1000 10 -o $W/ref- .csv > $W/ref.log 2>&1; rc_r=$?; t3=$(now)
t4=$(now); taskset -c 0 oracle/_ref/hammlet -f $W/c2.txt $COMMON -i F 100 10 -o $W/ref1- .csv > /dev/null 2>&1; t5=$(now)
python - "$W" "$OUT" $rc_r $t2 $t3 $t4 $t5 ${WALL[txt]} ${WALL[gz]} ${WALL[f32]} <<'PY'
import json, re, sys
W, out, rc_r = sys.argv[1], sys.argv[2], int(sys.argv[3])
t = [float(v) for v in sys.argv[4:8]]
w = sys.argv[8:]
def lines(p):
    try: return sum(1 for _ in open(p))
    except OSError: return -1
def timing(fmt):
    m = re.search(r"\[timing\] F 1000 sweeps in ([0-9.eE+-]+) s", open(f"{W}/our-{fmt}.err").read())
    return float(m.group(1)) if m else None
ours = {}
for i, fmt in enumerate(("txt", "gz", "f32")):
    t0, t1, rc = float(w[3 * i]), float(w[3 * i + 1]), int(w[3 * i + 2])
    s = timing(fmt)
    ours[fmt] = {"wall_s": t1 - t0, "rc": rc, "sampling_s": s, "sweeps_per_s": 1000 / s if s else None,
                 "load_and_output_s": (t1 - t0 - s) if s else None}
ref_1000, ref_100 = t[1] - t[0], t[3] - t[2]
d = {"config": "C2: 1e7 probes, K=5, auto priors, -i F 1000 10, marginals output; input as text (1e7 lines), gzip'd text, raw float32",
     "ours": ours, "reference_wall_s": ref_1000, "reference_rc": rc_r,
     "reference_sweeps_per_s": 900 / max(ref_1000 - ref_100, 1e-9),
     "reference_sweep_rate_how": "(wall of -i F 1000 10) - (wall of -i F 100 10) over 900 sweeps, both after a warm-up run",
     "ours_sweep_rate_how": "the binary's -timing line: steady clock around the 1000 sweeps",
     "ours_marginal_lines": lines(W + "/our-txt-marginals.csv"), "reference_marginal_lines": lines(W + "/ref-marginals.csv"),
     "reference": "oracle/_ref/hammlet (g++ -O3, 1 thread pinned with taskset -c 0)"}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
PY
rm -rf $W
