#!/usr/bin/env python
"""Per-stage kernel times of one Gibbs sweep at the block counts SURVEY.md §8d asks for (B ~ 1.3 k, 15 k, 190 k,
1.5 M, 1e7+), i.e. the shapes of BASELINE.json configs C1, C2, C4/8, C4 and C5, each as a dynamic FBG chain on one
B200.  CUDA events on the library's stream around every stage (hml_set_timing); one JSON line per configuration.

  python tools/scan_latency.py [--out gpurun_out/scan_latency.json] [--skip-c4]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the generator of the bench workload)

CONFIGS = [  # name, T, K, mean segment length, level spacing, sweeps timed
    ("C1  1e6 K=3", 1_000_000, 3, 5000, 1.0, 200),
    ("C2  1e7 K=5", 10_000_000, 5, 5000, 1.0, 200),
    ("C4/8  1.25e8 K=5", 125_000_000, 5, 5000, 1.0, 200),
    ("C4  1e9 K=5", 1_000_000_000, 5, 5000, 1.0, 100),
    ("C5  1e8 K=20 low compression", 100_000_000, 20, 50, 0.3, 30),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "scan_latency.json"))
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated substrings of the configuration names to run")
    args = ap.parse_args()
    import torch
    from hammlet_b200 import capi
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lines = []
    for name, T, K, L, spacing, steps in CONFIGS:
        if args.skip_c4 and T >= 1_000_000_000:
            continue
        if args.only and not any(tok in name for tok in args.only.split(",")):
            continue
        bench.SPACING = spacing
        t0 = time.time()
        x = bench.generate(torch, T, K, L, seed=4, device=device)
        torch.cuda.synchronize()
        h = capi.Handle(0)
        h.load_device(x.data_ptr(), T)
        del x
        torch.cuda.empty_cache()
        tau = capi.Chain.auto_prior(h, 0.2, 0.9)
        chain = capi.Chain(h, K, tau, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=100)
        chain.set(((np.arange(K) - (K - 1) / 2.0) * spacing).astype(np.float32), np.full(K, 0.09, np.float32),
                  (np.full((K, K), 0.0002 / (K - 1)) + np.eye(K) * (0.9998 - 0.0002 / (K - 1))).astype(np.float32),
                  np.full(K, 1.0 / K, np.float32))
        chain.run(10)
        stream = torch.cuda.ExternalStream(h.stream(), device=device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(stream)
        chain.run(steps)
        ev1.record(stream)
        torch.cuda.synchronize()
        dev_ms = ev0.elapsed_time(ev1) / steps
        w0 = time.perf_counter()
        chain.run(steps)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3 / steps
        h.set_timing(True)
        stage, nb = {}, []
        for _ in range(min(steps, 50)):
            nb.append(chain.run(1))
            for nm, ms in h.timing():
                stage.setdefault(nm, []).append(ms)
        h.set_timing(False)
        st = {k: float(np.mean(v)) * 1e3 for k, v in stage.items()}
        fwd = sum(v for k, v in st.items() if k.startswith("fwd_"))
        bwd = sum(v for k, v in st.items() if k.startswith("bwd_"))
        line = {"config": name, "T": T, "K": K, "blocks_per_sweep": float(np.mean(nb)), "compression": T / float(np.mean(nb)),
                "sweep_ms_device": dev_ms, "sweep_ms_wall": wall_ms, "sweeps_per_s": 1e3 / dev_ms,
                "forward_scan_us": fwd, "backward_scan_us": bwd, "stage_us": st,
                "setup_seconds": time.time() - t0}
        print(json.dumps(line), flush=True)
        lines.append(line)
        chain.close()
        h.close()
        torch.cuda.empty_cache()
    # ---- multivariate data (`-s C P D`): handle-level sweeps with a fixed model (the C chain of
    # include/hammlet_host.h is univariate); same stage table
    for name, T, P, D in (("MD 1e8 x 2 dims, P=2 (K=4)", 100_000_000, 2, 2), ("MD 5e7 x 3 dims, P=3 (K=27)", 50_000_000, 3, 3)):
        if args.only and not any(tok in name for tok in args.only.split(",")):
            continue
        K = P ** D
        gen = torch.Generator(device=device)
        gen.manual_seed(11)
        L = 5000
        change = torch.rand(T, generator=gen, device=device) < (1.0 / L)
        seg = torch.cumsum(change.to(torch.int32), 0, dtype=torch.int32).long()
        levels = torch.randint(0, P, (int(seg[-1].item()) + 1, D), generator=gen, device=device).to(torch.float32)
        x = (levels[seg] - (P - 1) / 2.0) + 0.3 * torch.randn(T, D, generator=gen, device=device)
        del change, seg, levels
        x = x.contiguous()
        torch.cuda.synchronize()
        h = capi.Handle(0)
        h.load_device_md(x.data_ptr(), T, D)
        del x
        torch.cuda.empty_cache()
        mapping = capi.combinations_mapping(P, D)
        mu = np.arange(P) - (P - 1) / 2.0
        var = np.full(P, 0.09)
        A = np.full((K, K), 0.0002 / (K - 1)) + np.eye(K) * (0.9998 - 0.0002 / (K - 1))
        pi = np.full(K, 1.0 / K)
        thr = float(np.sqrt(np.float32(2) * np.log(np.float32(T)) * np.float32(0.09)))
        for i in range(5):
            out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=1, sweep=i, mapping=mapping)
        steps = 50
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for i in range(steps):
            out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=1, sweep=10 + i, mapping=mapping)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3 / steps
        h.set_timing(True)
        stage = {}
        for i in range(20):
            h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=1, sweep=100 + i, mapping=mapping)
            for nm, ms in h.timing():
                stage.setdefault(nm, []).append(ms)
        h.set_timing(False)
        st = {k: float(np.mean(v)) * 1e3 for k, v in stage.items()}
        line = {"config": name, "T": T, "D": D, "P": P, "K": K, "blocks_per_sweep": int(out["nblocks"]),
                "sweep_ms_wall": wall_ms, "sweeps_per_s": 1e3 / wall_ms, "stage_us": st,
                "how": "hml_fb_sweep from Python (ctypes) with a fixed model, wall clock incl. the call overhead"}
        print(json.dumps(line), flush=True)
        lines.append(line)
        h.close()
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
