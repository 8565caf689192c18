"""Pins the extrapolation bench.py's reference arm makes: the reference's own sampleHMM (oracle/_ref/ref_probe, one
thread) on the first Ts observations of the bench workload for Ts = 3e7, 1e8, 3e8 and the full 1e9, measured sweep
rate against the rate extrapolated from the smaller samples (sweep cost linear in the number of blocks).  CPU only:
needs ~25 GB of RAM and ~15 minutes for the 1e9 run.  Writes one JSON document to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import refprobe  # noqa: E402
import torch  # noqa: E402

T, K, L = 10**9, 5, 5000
sizes = [int(float(a)) for a in sys.argv[1:]] or [30_000_000, 100_000_000, 300_000_000, 1_000_000_000]
dev = "cuda" if torch.cuda.is_available() else "cpu"
x = bench.generate(torch, T, K, L, seed=4, device=dev, limit=max(sizes)).cpu().numpy()
rows = []
for Ts in sizes:
    t0 = time.time()
    r = refprobe.run("bench", x[:Ts], K=K, seed=1, burn=100, timed=20, reps=2, method="F", raw32=True)
    secs = float(np.min(r["bench_secs"]))
    rows.append({"sample_T": Ts, "blocks": int(r["bench_blocks"][0]), "sweeps_per_s": 20 / secs,
                 "us_per_block": secs / 20 / int(r["bench_blocks"][0]) * 1e6,
                 "extrapolated_to_T": 20 / secs * Ts / T, "wall_s": time.time() - t0})
    print(rows[-1], file=sys.stderr, flush=True)
full = next((r for r in rows if r["sample_T"] == T), None)
doc = {"workload": f"first Ts of the bench.py sequence (T={T}, K={K}, mean segment {L}), reference sampleHMM, 1 thread, "
                   f"100 burn-in + best of 2 x 20 timed sweeps", "host": bench.cpu_model(), "generator_device": dev,
       "rows": rows}
if full:
    doc["measured_full_config_sweeps_per_s"] = full["sweeps_per_s"]
    doc["extrapolated_over_measured"] = {str(r["sample_T"]): r["extrapolated_to_T"] / full["sweeps_per_s"] for r in rows}
print(json.dumps(doc, indent=1))
