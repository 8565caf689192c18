"""A reduced selection of the GPU suite for compute-sanitizer (tools/gpu.sh sanitize): one case per kernel family,
sized so that `--tool memcheck` and `--tool racecheck` finish in minutes — load (univariate, multivariate, non-power-of-
two lengths), the three boundary-detection modes and a candidate-list rebuild, block statistics, the multi-kernel sweep
for K = 5 (operator scan and speculative filter incl. its fall-back, replay and Philox, log-likelihood and kept rows; the
kernels whose last CTA finishes the step), K = 20 (wide path) and multivariate data,
the mixture sampler, runs and device-side state marginals of recorded sweeps, the persistent fused sweep and the
device-resident chain (spin barriers across CTAs, shared-memory hand-offs), capacity growth and a handle reload.
Every result is still checked against the oracle where that is cheap, so a sanitizer-clean run is also a correct one."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import oracle  # noqa: E402
from hammlet_b200 import capi  # noqa: E402
from hammlet_b200.synth import model_guess, piecewise_gaussian, piecewise_gaussian_md  # noqa: E402


def main():
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    h = capi.Handle(0)
    # ---- load + detection modes
    for T in (1, 11, 4097, 70_001):
        x = piecewise_gaussian(T, 3, 50, seed=T % 97 + 1)
        h.load(x)
        assert np.array_equal(h.weights().view(np.uint32), O32.weights(x).view(np.uint32))
    T = 150_000
    x = piecewise_gaussian(T, 5, 200, seed=3)
    w = O32.weights(x)
    h.load(x)
    for mode in (capi.DETECT_STREAM, capi.DETECT_PYRAMID, capi.DETECT_CANDIDATES):
        h.set_detect_mode(mode)
        for thr in (0.9, 0.3, 1.4, 1e-30):
            B = h.create_blocks(thr)
            assert np.array_equal(h.blocks(stats=False).astype(np.uint64), O32.boundaries(w, np.float32(thr))), (mode, thr)
            assert B > 0
    h.set_detect_mode(capi.DETECT_CANDIDATES)
    print("load + detection ok", flush=True)
    # ---- multi-kernel sweeps: K = 5 (replay, rows, loglik), Philox dynamic, mixture
    mu, var, A, pi = model_guess(5, seed=5)
    thr = 0.8
    B = h.create_blocks(thr)
    starts = O32.boundaries(w, np.float32(thr))
    n, s, q = O64.block_stats(O64.integral(x), starts, T)
    u = np.random.default_rng(1).random(B)
    ref = O64.fb_sweep(n, s, q, mu, var, A, pi, 1, u)
    out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_LOGLIK | capi.SWEEP_KEEP_ROWS, replay=u)
    assert np.array_equal(h.states(), ref["states"]) and np.array_equal(out["trans"], ref["trans"])
    h.rows(5)
    for i in range(3):
        out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.8 + 0.05 * i, seed=3, sweep=i)
        assert out["trans"].sum() == T
    out = h.mix_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.8, seed=3, sweep=9)
    assert out["counts"].sum() == T
    # ---- the forward filter both ways (operator scan; speculative pieces + repair pass), and a model on which the
    # speculative pass has to give up and the sweep repeats itself through the operator scan
    B = h.create_blocks(thr)
    for mode in (capi.FORWARD_OPERATORS, capi.FORWARD_SPECULATIVE):
        h.set_forward_mode(mode)
        out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_LOGLIK | capi.SWEEP_KEEP_ROWS, replay=u)
        assert np.array_equal(h.states(), ref["states"]) and np.array_equal(out["trans"], ref["trans"]), mode
    flat_var = np.full(5, 1e6, np.float32)
    flat_A = (np.full((5, 5), 1e-7) + np.eye(5) * (1 - 5e-7)).astype(np.float32)
    failed = h.forward_info()[2]
    out = h.fb_sweep(mu, flat_var, flat_A, pi, replay=u)
    assert h.forward_info()[2] == failed + 1 and out["trans"].sum() == T
    h.set_forward_mode(capi.FORWARD_AUTO)
    print("K=5 sweeps ok", flush=True)
    # ---- recorded sweeps: runs + marginals on the device
    h.marginals_reset(5)
    for i in range(3):
        h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.8, seed=4, sweep=i)
        sizes, _ = h.segments()
        assert sizes.sum() == T
        h.marginals_add()
    ms, mc, it = h.marginals()
    assert ms.sum() == T and it == 3 and np.all(mc.sum(1) == 3)
    print("recorded sweeps ok", flush=True)
    # ---- the persistent kernel: single fused sweeps (replay and Philox) and a device-resident chain
    B = h.create_blocks(thr)
    out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_FUSED, replay=u)
    assert np.array_equal(h.states(), ref["states"]) and np.array_equal(out["trans"], ref["trans"])
    a = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.85, seed=5, sweep=1)
    sa = h.states()
    b = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED, threshold=0.85, seed=5, sweep=1)
    assert np.array_equal(sa, h.states()) and np.array_equal(a["trans"], b["trans"])
    tau = capi.Chain.auto_prior(h, 0.2, 0.9)
    h.chain_init(5, tau, seed=7)
    h.chain_set(mean=mu, var=var, A=A, pi=pi)
    out = h.chain_run(6)
    assert out["trans"].sum() == T and out["counts"].sum() == T
    print("fused sweep + chain ok (fused sweeps: %d)" % out["fused"], flush=True)
    # ---- K = 20 (wide path), low compression
    T2 = 60_000
    x2 = piecewise_gaussian(T2, 20, 6, seed=9, spacing=0.3)
    mu2, var2, A2, pi2 = model_guess(20, seed=20, spacing=0.3)
    h.load(x2)
    out = h.fb_sweep(mu2, var2, A2, pi2, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_LOGLIK, threshold=0.3, seed=1, sweep=0)
    assert out["trans"].sum() == T2
    print("K=20 sweep ok", flush=True)
    # ---- capacity growth + reload on a used handle
    x3 = piecewise_gaussian(300_000, 3, 3, seed=5)
    h.load(x3)
    mu3, var3, A3, pi3 = model_guess(3, seed=3)
    out = h.fb_sweep(mu3, var3, A3, pi3, flags=capi.SWEEP_DYNAMIC, threshold=0.1, seed=1, sweep=0)
    assert out["nblocks"] > 65536 and out["trans"].sum() == 300_000
    # every observation a block: more tiles than SMs, i.e. block maps, chunk maps + scan and replay + statistics as
    # separate launches (short lists take the CTA-per-tile map kernel)
    out = h.fb_sweep(mu3, var3, A3, pi3, flags=capi.SWEEP_DYNAMIC, threshold=0.0, seed=1, sweep=1)
    assert out["nblocks"] == 300_000 and out["trans"].sum() == 300_000
    h.load(x[:20000])
    out = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.8, seed=1, sweep=0)
    assert out["trans"].sum() == 20000
    print("capacity growth + reload ok", flush=True)
    # ---- multivariate data
    P, D, T4 = 2, 2, 30_000
    xm = piecewise_gaussian_md(T4, P, D, 250, 7, quantum_bits=10)
    h.load(xm)
    K4 = P ** D
    mapping = np.array([[(s // (P ** d)) % P for d in range(D)] for s in range(K4)], np.int32)
    mu4, var4, _, _ = model_guess(P, seed=2)
    _, _, A4, pi4 = model_guess(K4, seed=2)
    out = h.fb_sweep(mu4, var4, A4, pi4, flags=capi.SWEEP_DYNAMIC, threshold=0.9, seed=2, sweep=0, mapping=mapping)
    assert out["trans"].sum() == T4
    print("multivariate sweep ok", flush=True)
    h.close()
    print("SANITIZE CASES OK", flush=True)


if __name__ == "__main__":
    main()
