"""bench.py's output contract: the keys of the JSON line (checked on the committed line of the final tree, which a GPU
box produced) and the reference arm, which needs no GPU and is run here on a small sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_committed_bench_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r3s_bench_line.json")   # bench.py with default flags on the final tree of round 2
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(d)
    assert {"gpu_launches", "clocks", "roofline", "cpu_baseline", "invariants_ok", "forward_filter", "recorded"} <= set(d)
    assert d["invariants_ok"] is True and d["invariants"]["sum_trans"] == d["config"]["observations"]
    assert d["forward_filter"]["after_timed_regions"]["repeated"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["unit"] == "sweeps/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] >= d["steps"] * 8          # eight kernels per sweep
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert abs(d["value"] - d["n_gpus"] * 1000.0 / d["ms_per_step"]) / d["value"] < 1e-6


def test_reference_arm_runs_without_a_gpu():
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_probe")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_probe not built")
    def arm(T, sample):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--T", T, "--ref-sample",
                            sample, "--steps", "5", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert p.returncode == 0, p.stderr[-2000:]
        lines = [ln for ln in p.stdout.strip().split("\n") if ln.startswith("{")]
        assert len(lines) == 1
        return json.loads(lines[0])

    d = arm("1e6", "1e6")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0
    # the whole workload was swept: a measurement, and the line says so at the top level
    assert d["extrapolated"] is False and d["same_config"] is True and d["sample_T"] == 1_000_000
    assert abs(d["value"] - d["measured_on_sample_sweeps_per_s"]) < 1e-9
    # a sample of the workload: the scaled number is flagged as extrapolated where nobody can miss it
    e = arm("2e6", "1e6")
    assert e["extrapolated"] is True and e["same_config"] is False and e["sample_T"] == 1_000_000
    assert abs(e["value"] - e["measured_on_sample_sweeps_per_s"] * 0.5) < 1e-9


def test_traffic_capture_is_bound_to_the_kernel_sources():
    """profiles/ncu_traffic.json carries the hash of the kernel sources it was captured from; bench.py reports the
    DRAM traffic only while that hash matches the tree (a stale capture gives traffic = null and says why)."""
    sys.path.insert(0, ROOT)
    import bench
    doc = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert len(bench.kernel_sources_hash()) == 64
    assert "_kernel_sources_sha256" in doc and "_git_head" in doc
