"""The whole sweep in one persistent kernel (HML_SWEEP_FUSED, csrc/hml_fused.cuh) and the device-resident Gibbs chain
built on it (hml_chain_*; SURVEY.md §8f.4: conjugate updates and parameter draws on the device).

Bars: a fused sweep is THE sweep — under uniform replay its states are the oracle's (the reference's), its counts are
exact, its sums within 1e-9; with Philox uniforms it gives the multi-kernel sweep's states bit for bit.  The parameter
phase reproduces the reference's Normal-Inverse-Gamma / Dirichlet posterior algebra (checked through the moments of
many draws from known statistics) and chains agree with the reference in distribution (tests/test_host_cli.py)."""
import numpy as np
import pytest

import oracle
from hammlet_b200 import capi
from hammlet_b200.synth import model_guess, piecewise_gaussian

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def rel_err(a, b, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    s = np.where(s == 0, 1.0, s)
    return float(np.max(np.abs(a - b) / s)) if a.size else 0.0


@pytest.fixture(scope="module")
def dev():
    h = capi.Handle(0)
    yield h
    h.close()


@pytest.mark.parametrize("T,K,L,thr,use_self", [
    (3000, 3, 100, 0.9, 1), (50_000, 5, 500, 1.2, 1), (50_000, 5, 500, 0.4, 0), (20_000, 2, 50, 0.8, 1),
    (200_000, 4, 300, 0.6, 1), (30_000, 8, 300, 0.7, 1), (60_000, 6, 100, 0.5, 1), (2_000_000, 5, 500, 0.5, 1),
    (1, 3, 5, 1.0, 1), (2, 3, 5, 1.0, 1), (33, 3, 5, 0.1, 1), (1_000_000, 3, 5000, 1.5, 1)])
def test_fused_sweep_replay_vs_oracle(dev, T, K, L, thr, use_self):
    """Same inputs and bars as test_gpu_parity.py::test_fb_sweep_replay_vs_oracle, through the persistent kernel."""
    x = piecewise_gaussian(T, K, L, seed=T % 89 + K)
    mu, var, A, pi = model_guess(K, seed=K)
    dev.load(x)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    ref_starts = O32.boundaries(O32.weights(x), thr)
    n, s, q = O64.block_stats(O64.integral(x), ref_starts, T)
    B = dev.create_blocks(thr)
    assert B == ref_starts.size
    u = np.random.default_rng(123).random(B)
    ref = O64.fb_sweep(n, s, q, mu, var, A, pi, use_self, u)
    if ref["rc"] != 0:
        pytest.skip("the oracle took the uniform fallback: that sweep belongs to the sequential kernel")
    launches = dev.launch_count()
    out = dev.fb_sweep(mu, var, A, pi, use_self=use_self, flags=capi.SWEEP_FUSED, replay=u)
    assert out["nblocks"] == B
    if B <= 65536:
        assert dev.launch_count() - launches == 1, "one cooperative launch for the whole sweep"
    assert np.array_equal(dev.states(), ref["states"])                # sample-exact under replay
    assert np.array_equal(out["trans"], ref["trans"]) and np.array_equal(out["counts"], ref["counts"])
    assert np.array_equal(out["stat_n"], ref["stat_n"]) and out["trans"].sum() == T
    nz = ref["stat_n"] > 0
    assert rel_err(out["stat_sq"][nz], ref["stat_sq"][nz]) <= RTOL
    assert rel_err(out["stat_sum"][nz], ref["stat_sum"][nz], scale=np.sqrt(ref["stat_n"] * ref["stat_sq"])[nz]) <= RTOL
    starts, bs, bq = dev.blocks()
    assert np.array_equal(starts.astype(np.uint64), ref_starts)
    msq = float(np.mean(x.astype(np.float64) ** 2))
    assert rel_err(bq, q, scale=msq) <= RTOL
    seg_n, seg_s = dev.segments()
    rn, rs = oracle.merge_runs(ref["states"], n)
    assert np.array_equal(seg_n.astype(np.int64), rn) and np.array_equal(seg_s.astype(np.int64), rs)


@pytest.mark.parametrize("T,K,L", [(400_000, 5, 300), (3_000_000, 3, 2000), (10_000_000, 5, 5000)])
def test_fused_dynamic_philox_equals_the_multi_kernel_sweep(dev, T, K, L):
    """Philox uniforms are indexed by the block number, so the persistent kernel and the 13-kernel sweep must sample the
    same states from the same model — on thresholds that move (dynamic blocks), over a few sweeps."""
    x = piecewise_gaussian(T, K, L, seed=K + 11)
    mu, var, A, pi = model_guess(K, seed=K + 1)
    dev.load(x)
    for i, thr in enumerate((1.1, 1.05, 1.2, 0.95)):
        a = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=7, sweep=i)
        sa, ba = dev.states(), dev.blocks(stats=False)
        b = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED, threshold=thr, seed=7, sweep=i)
        sb, bb = dev.states(), dev.blocks(stats=False)
        assert a["nblocks"] == b["nblocks"] and np.array_equal(ba, bb)
        assert np.array_equal(sa, sb), f"sweep {i}: states differ"
        for k in ("trans", "counts", "stat_n"):
            assert np.array_equal(a[k], b[k]), k
        assert rel_err(b["stat_sum"], a["stat_sum"], scale=1.0) <= 1e-12 and rel_err(b["stat_sq"], a["stat_sq"]) <= 1e-12


def test_fused_sweep_stands_down_where_it_does_not_apply(dev):
    """More blocks than 64 tiles, a threshold the candidate list cannot serve: the call still returns the sweep (through
    the multi-kernel path); K > 8 or a log-likelihood request are refused."""
    T = 3_000_000
    x = piecewise_gaussian(T, 3, 6, seed=4)
    mu, var, A, pi = model_guess(3, seed=4)
    dev.load(x)
    a = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.3, seed=1, sweep=0)
    assert a["nblocks"] > 65536
    sa = dev.states()
    b = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED, threshold=0.3, seed=1, sweep=0)
    assert b["nblocks"] == a["nblocks"] and np.array_equal(dev.states(), sa)
    c = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED, threshold=1e-30, seed=1, sweep=0)
    O32 = oracle.Oracle(False)
    assert c["nblocks"] == O32.boundaries(O32.weights(x), np.float32(1e-30)).size > T // 2
    with pytest.raises(capi.HmlError):
        dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED | capi.SWEEP_LOGLIK, threshold=1.0)
    mu9, var9, A9, pi9 = model_guess(9, seed=1)
    with pytest.raises(capi.HmlError):
        dev.fb_sweep(mu9, var9, A9, pi9, flags=capi.SWEEP_DYNAMIC | capi.SWEEP_FUSED, threshold=1.0)


def test_chain_is_reproducible_and_independent_of_batching():
    """hml_chain_run: n sweeps in one launch, or one by one, or in uneven batches — the same chain (Philox streams keyed
    by the sweep number), and the statistics of the last sweep satisfy the counting invariants."""
    T, K = 1_000_000, 3
    x = piecewise_gaussian(T, K, 5000, seed=1)
    finals = []
    for batches in ([40], [1] * 40, [7, 13, 20]):
        h = capi.Handle(0)
        try:
            h.load(x)
            tau = capi.Chain.auto_prior(h, 0.2, 0.9)
            h.chain_init(K, tau, seed=42)
            fused = 0
            for n in batches:
                out = h.chain_run(n)
                fused += out["fused"]
            # 1e6 observations, ~1.3 k blocks: the persistent kernel, except for the odd sweep right after the draw from
            # the priors whose variance gives a threshold no candidate list is worth building for
            assert fused >= 35
            assert out["trans"].sum() == T and out["counts"].sum() == T and out["stat_n"].sum() == T
            g = h.chain_get()
            assert g["sweeps"] == 41          # the leading draw from the priors + 40 sweeps
            finals.append((g["mean"].copy(), g["var"].copy(), g["A"].copy(), g["pi"].copy(), h.states().copy(), out["nblocks"]))
        finally:
            h.close()
    for f in finals[1:]:
        for a, b in zip(finals[0][:5], f[:5]):
            assert np.array_equal(a, b)
        assert f[5] == finals[0][5]


def test_chain_parameter_phase_moments():
    """The parameter phase against the closed forms of the reference's conjugate algebra (Conjugate.hpp:120-205): with the
    state sequence pinned by an extremely peaked model on well separated data, the statistics of every sweep are (almost)
    the same, so over many sweeps the drawn variances, means, pi and rows of A must have the moments of
    InvGamma(alpha', beta'), N(mu0', var / nu'), Dirichlet(alpha_I + occupancy), Dirichlet(prior + transitions)."""
    T, K = 400_000, 3
    rng = np.random.default_rng(5)
    seg = np.repeat(rng.integers(0, K, T // 2000), 2000)
    x = (seg * 4.0 + rng.normal(0, 0.3, T)).astype(np.float32)      # levels 0, 4, 8: no state confusion
    h = capi.Handle(0)
    try:
        h.load(x)
        prior = np.array([2.0, 1.0, 4.0, 0.01], np.float32)
        h.chain_init(K, prior, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=9)
        h.chain_set(mean=[0.0, 4.0, 8.0], var=[0.09] * 3, A=np.full((3, 3), 0.001) + np.eye(3) * 0.997, pi=[1 / 3] * 3)
        draws = []
        out = h.chain_run(20)                                         # burn-in
        for _ in range(600):
            out = h.chain_run(1)
            g = h.chain_get()
            order = np.argsort(g["mean"])
            draws.append((g["mean"][order], g["var"][order], g["pi"][order], g["A"][np.ix_(order, order)], out))
        assert out["fused"] == 1
        means = np.array([d[0] for d in draws]); vars_ = np.array([d[1] for d in draws]); pis = np.array([d[2] for d in draws])
        last = draws[-1][4]
        # expected posterior from the (stable) statistics of a sweep, in the reference's formulas
        order = np.argsort(h.chain_get()["mean"])
        n = last["stat_n"][order].astype(np.float64); s = last["stat_sum"][order]; q = last["stat_sq"][order]
        a0, b0, m0, nu0 = [float(v) for v in prior]
        xbar = s / n
        a1 = a0 + n / 2
        b1 = b0 + ((q + (n * nu0 / (n + nu0)) * (xbar - m0) ** 2) - np.minimum(s * s / n, q)) / 2
        m1 = (nu0 * m0 + s) / (nu0 + n)
        exp_var = b1 / (a1 - 1)                                      # mean of InvGamma(a1, b1)
        assert np.allclose(vars_.mean(0), exp_var, rtol=0.01), (vars_.mean(0), exp_var)
        sd_var = exp_var / np.sqrt(a1 - 2)
        assert np.allclose(vars_.std(0), sd_var, rtol=0.15), (vars_.std(0), sd_var)
        assert np.allclose(means.mean(0), m1, atol=4 * np.sqrt(exp_var / (nu0 + n)) / np.sqrt(len(draws)) + 1e-4)
        assert np.allclose(means.std(0), np.sqrt(exp_var / (nu0 + n)), rtol=0.15)
        occ = last["counts"][order].astype(np.float64) + 0.5
        assert np.allclose(pis.mean(0), occ / occ.sum(), atol=5e-3)
        Arows = np.array([d[3] for d in draws]).mean(0)
        tr = last["trans"][np.ix_(order, order)].astype(np.float64) + 0.5
        assert np.allclose(Arows, tr / tr.sum(1, keepdims=True), atol=2e-3)
    finally:
        h.close()


def test_chain_takes_the_multi_kernel_path_when_it_must():
    """Low compression (more than 64 tiles of blocks): every sweep goes through the 13-kernel path, the parameter phase
    still runs on the device; the chain is the same Philox chain either way."""
    T, K = 3_000_000, 3
    x = piecewise_gaussian(T, K, 6, seed=4)
    h = capi.Handle(0)
    try:
        h.load(x)
        tau = capi.Chain.auto_prior(h, 0.2, 0.9)
        h.chain_init(K, tau, seed=3)
        out = h.chain_run(5)
        assert out["fused"] < 5 and out["trans"].sum() == T and out["counts"].sum() == T
        assert h.chain_get()["sweeps"] == 6
    finally:
        h.close()
