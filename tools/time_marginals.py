#!/usr/bin/env python
"""Times the pieces of a recorded sweep on one B200: sweep, run formation (hml_get_segments), device-side marginal
merge (hml_marginals_add), at T = 1e9 / K = 5 (or --T)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=float, default=1e9)
    args = ap.parse_args()
    import torch
    from hammlet_b200 import capi
    T, K = int(args.T), 5
    dev = torch.device("cuda", 0)
    x = bench.generate(torch, T, K, 5000, seed=4, device=dev)
    h = capi.Handle(0)
    h.load_device(x.data_ptr(), T)
    del x
    mu = (np.arange(K) - 2.0)
    var = np.full(K, 0.09)
    A = np.full((K, K), 0.0002 / 4) + np.eye(K) * (0.9998 - 0.0002 / 4)
    pi = np.full(K, 0.2)
    thr = float(np.sqrt(np.float32(2) * np.log(np.float32(T)) * np.float32(0.09)))
    h.marginals_reset(K)

    def t(f, n=1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    for i in range(12):
        ts = t(lambda: h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr * (1 + 0.001 * i), seed=i, sweep=i))
        ta = t(h.marginals_add)
        tg = t(h.segments)
        n, _, it = h.marginals()
        print(f"iter {i}: sweep {ts:.3f} ms  marginals_add {ta:.3f} ms  get_segments {tg:.3f} ms  segments {n.size} iterations {it}", flush=True)


def chain_level(T=int(1e9)):
    import torch
    from hammlet_b200 import capi
    K = 5
    dev = torch.device("cuda", 0)
    x = bench.generate(torch, T, K, 5000, seed=4, device=dev)
    h = capi.Handle(0)
    h.load_device(x.data_ptr(), T)
    del x
    tau = capi.Chain.auto_prior(h, 0.2, 0.9)
    chain = capi.Chain(h, K, tau, seed=100)
    chain.set((np.arange(K) - 2.0).astype(np.float32), np.full(K, 0.09, np.float32),
              (np.full((K, K), 0.0002 / 4) + np.eye(K) * (0.9998 - 0.0002 / 4)).astype(np.float32), np.full(K, 0.2, np.float32))
    chain.run(20)
    for n in (20, 20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); chain.run(n); torch.cuda.synchronize()
        print(f"chain.run({n}): {(time.perf_counter() - t0) / n * 1e3:.3f} ms per sweep", flush=True)
    for n in (1, 1, 1, 1, 5, 20, 20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); nb, ns = chain.run_recorded(n, thinning=1); torch.cuda.synchronize()
        print(f"chain.run_recorded({n}): {(time.perf_counter() - t0) / n * 1e3:.3f} ms per sweep, blocks {nb}, marginal segments {ns}", flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); chain.run_recorded(21, thinning=3); torch.cuda.synchronize()
    print(f"chain.run_recorded(21, thinning 3): {(time.perf_counter() - t0) / 21 * 1e3:.3f} ms per sweep", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "chain":
        chain_level()
        sys.exit(0)
    main()
