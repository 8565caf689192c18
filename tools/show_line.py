"""Prints the parts of a bench.py JSON line one reads first."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 4), "| e2e", round(d["e2e"]["value"], 1),
      "| launches", d.get("gpu_launches"), "| recorded", (d.get("recorded") or {}).get("value"))
if d.get("stage_ms"):
    print({k: round(v * 1000, 1) for k, v in d["stage_ms"].items()})
r = d.get("roofline") or {}
print("roofline", r.get("kernel"), "frac", r.get("frac"), "traffic", r.get("traffic"), "| clocks", d["clocks"]["sm_mhz"],
      d["clocks"]["reasons"], "| invariants", d.get("invariants_ok"))
print("cpu_baseline", d.get("cpu_baseline"))
if d.get("forward_filter"):
    print("forward filter", d["forward_filter"]["after_timed_regions"], "->", d["forward_filter"]["after_stage_pass"])
