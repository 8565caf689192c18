"""profiles/ncu_traffic.json from an `ncu --set full` raw CSV of one bench sweep (tools/gpu.sh full): DRAM bytes read +
written per launch of every per-sweep kernel, stamped with the git head and a hash of the kernel sources the capture
was taken from, so bench.py can tell a stale capture from a current one.
usage: python tools/ncu_traffic.py RAW.csv OUT.json [T] [K]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

raw, out = sys.argv[1], sys.argv[2]
T = int(float(sys.argv[3])) if len(sys.argv) > 3 else 10**9
K = int(sys.argv[4]) if len(sys.argv) > 4 else 5
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


doc = {}
for r in rows[2:]:
    m = re.search(r"(k_[a-z_0-9]+)", r[col["Kernel Name"]])
    if not m:
        continue
    name = m.group(1)
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    du = units[col["gpu__time_duration.sum"]]
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
    rec = doc.setdefault(name, {"T": T, "K": K, "launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ncu_us": 0.0})
    rec["launches"] += 1
    rec["dram_read"] += rd
    rec["dram_write"] += wr
    rec["ncu_us"] += dur_us
for rec in doc.values():
    n = rec.pop("launches")
    rec["dram_read_bytes_per_launch"] = rec.pop("dram_read") / n
    rec["dram_write_bytes_per_launch"] = rec.pop("dram_write") / n
    rec["dram_bytes_per_launch"] = rec["dram_read_bytes_per_launch"] + rec["dram_write_bytes_per_launch"]
    rec["ncu_us_per_launch"] = rec.pop("ncu_us") / n
    rec["launches_in_capture"] = n
# bench.py names the roofline kernel after its stage ("k_" + stage name)
for alias, kernel in (("k_fwd_chunks", "k_fwd_chunks_prefix"), ("k_fwd_replay", "k_fwd_replay_prefix"), ("k_fwd_spec", "k_fwd_replay_prefix"),
                      ("k_detect_cand", "k_cand_count"), ("k_detect_scatter", "k_cand_scatter")):
    if kernel in doc and alias not in doc:
        doc[alias] = dict(doc[kernel], alias_of=kernel)
try:
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
except OSError:
    head = ""
doc["_git_head"] = head or os.environ.get("HML_GIT_HEAD", "unknown (captured from a snapshot without .git)")
doc["_kernel_sources_sha256"] = bench.kernel_sources_hash()
doc["_source"] = os.path.basename(raw)
with open(out, "w") as f:
    json.dump(doc, f, indent=1, sort_keys=True)
print(json.dumps({k: v.get("dram_bytes_per_launch") for k, v in doc.items() if isinstance(v, dict)}, indent=1))
