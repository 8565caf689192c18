"""Sweep rate of a Gibbs chain on one GPU for the small and medium configurations, host-driven against
device-resident: `Chain.run` (C++ host chain: 8 kernels + one host round trip per sweep, parameters drawn on the host
with libstdc++ <random>) and `hml_chain_run` (parameters on the device; the whole sweep in one persistent kernel where
the block structure has at most 64 tiles, n sweeps per launch).  One JSON line per configuration.
usage: python tools/chain_rate.py [c1 c2 c3chr c4over8]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from hammlet_b200 import capi  # noqa: E402
from hammlet_b200.synth import piecewise_gaussian  # noqa: E402

CONFIGS = {
    "c1": dict(T=1_000_000, K=3, L=5000, seed=1, note="BASELINE configs[0]: 1e6 points, K=3"),
    "c2": dict(T=10_000_000, K=5, L=5000, seed=2, note="BASELINE configs[1]: 1e7 probes, K=5"),
    "c3chr": dict(T=50_000_000, K=5, L=50_000, seed=121, note="BASELINE configs[2]: one chromosome-sized sequence (chr21-like)"),
    "c4over8": dict(T=125_000_000, K=5, L=5000, seed=4, note="BASELINE configs[3]: the share of one of 8 GPUs, as a sequence of its own"),
}


def main():
    names = sys.argv[1:] or ["c1", "c2", "c3chr"]
    for name in names:
        cfg = CONFIGS[name]
        T, K = cfg["T"], cfg["K"]
        x = piecewise_gaussian(T, K, cfg["L"], seed=cfg["seed"])
        h = capi.Handle(0)
        h.load(x)
        tau = capi.Chain.auto_prior(h, 0.2, 0.9)
        levels = ((np.arange(K) - (K - 1) / 2.0)).astype(np.float64)
        A0 = np.full((K, K), 0.0002 / (K - 1)) + np.eye(K) * (0.9998 - 0.0002 / (K - 1))
        # ---- host-driven chain (HAMMLET_HOST_PARAMS keeps sampleHMM off the device-resident chain)
        os.environ["HAMMLET_HOST_PARAMS"] = "1"
        chain = capi.Chain(h, K, tau, seed=100)
        chain.set(levels.astype(np.float32), np.full(K, 0.09, np.float32), A0.astype(np.float32), np.full(K, 1.0 / K, np.float32))
        chain.run(200)
        h.sync()
        n = 2000 if T <= 10_000_000 else 500
        t0 = time.perf_counter()
        nb_host = chain.run(n)
        h.sync()
        host_rate = n / (time.perf_counter() - t0)
        chain.close()
        del os.environ["HAMMLET_HOST_PARAMS"]
        # ---- device-resident chain
        h.chain_init(K, tau, seed=100)
        h.chain_set(mean=levels, var=np.full(K, 0.09), A=A0, pi=np.full(K, 1.0 / K))
        h.chain_run(200)
        t0 = time.perf_counter()
        out = h.chain_run(n)
        dev_rate = n / (time.perf_counter() - t0)
        # ... one sweep per call: what a caller that records every sweep gets
        t0 = time.perf_counter()
        for _ in range(200):
            out1 = h.chain_run(1)
        single_rate = 200 / (time.perf_counter() - t0)
        h.set_timing(True)
        h.chain_run(50)
        st = h.chain_phase_ns().astype(np.int64)
        h.set_timing(False)
        names = ["model+count", "barrier1", "scatter", "barrier2", "stats+emit+chunk_ops", "barrier3", "rows+maps+chunk_maps",
                 "barrier4", "states+reduce", "barrier5", "final+params", "barrier6"]
        phases = {names[i]: float(st[i + 1] - st[i]) / 1000.0 for i in range(12)} if st[12] > st[0] > 0 else None
        if phases and st[13] > 0:   # inside "rows+maps+chunk_maps" (CTA 0's first quarter)
            phases["  of which rows"] = float(st[13] - st[6]) / 1000.0
            phases["  of which maps"] = float(st[14] - st[13]) / 1000.0
            phases["  of which chunk maps"] = float(st[15] - st[14]) / 1000.0
            phases["  of which tile barrier + tile map"] = float(st[7] - st[15]) / 1000.0
        ok = bool(out["trans"].sum() == T and out["counts"].sum() == T and out["stat_n"].sum() == T)
        print(json.dumps({"config": name, "note": cfg["note"], "T": T, "K": K, "blocks_host_chain": int(nb_host),
                          "blocks_device_chain": int(out["nblocks"]), "sweeps_timed": n,
                          "host_chain_sweeps_per_s": host_rate, "host_chain_us_per_sweep": 1e6 / host_rate,
                          "device_chain_sweeps_per_s": dev_rate, "device_chain_us_per_sweep": 1e6 / dev_rate,
                          "device_chain_fused_sweeps": int(out["fused"]), "device_chain_one_sweep_per_call_per_s": single_rate,
                          "speedup": dev_rate / host_rate, "fused_phase_us": phases, "invariants_ok": ok and int(out1["fused"]) in (0, 1)}), flush=True)
        h.close()


if __name__ == "__main__":
    main()
