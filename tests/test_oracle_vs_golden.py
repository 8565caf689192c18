"""Pins the oracle (oracle/hammlet_oracle_impl.h) to the reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from the reference's own
classes (oracle/ref_probe.cpp, compiled from /root/reference/src).  Every comparison here is
bit-for-bit, for real_t=float (what `hammlet` ships) and real_t=double (the fp64 oracle)."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle
import refprobe

QB = 10


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(path):
    g = np.load(path, allow_pickle=False)
    x = g["xq"].astype(np.float32) / (1 << QB)
    return g, x


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


WEIGHT_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "weights_T*.npz")))
SWEEP_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fb_*.npz")) +
                     glob.glob(os.path.join(os.path.dirname(__file__), "golden", "mix_*.npz")))


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("path", WEIGHT_CASES, ids=os.path.basename)
def test_weights_bitwise(path, fp64):
    g, x = load(path)
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    c = O.maxlet(x)
    w = O.breakpoint_weights(c)
    if "coeffs" + tag in g:
        assert same(c, g["coeffs" + tag]) and same(w, g["weights" + tag])
    else:
        assert digest(c) == str(g["coeffs_sha" + tag]) and digest(w) == str(g["weights_sha" + tag])
        assert same(np.flatnonzero(np.isinf(w)), g["inf_pos" + tag])
    if x.size >= 2:
        assert O.sigma_hat(c) == g["sigma_hat" + tag][0]


def test_forced_breakpoints_quirk():
    """the `R < T` test also kills complete wavelets ending exactly at T."""
    O = oracle.Oracle(False)
    for T, expect in ((8, [0, 4, 6, 7]), (16, [0, 8, 12, 14, 15]), (11, [0, 8, 10])):
        w = O.weights(np.arange(T, dtype=np.float32))
        assert list(np.flatnonzero(np.isinf(w))) == expect


@pytest.mark.parametrize("fp64", [False, True])
def test_blocks_and_autoprior(golden_dir, fp64):
    g, x = load(os.path.join(golden_dir, "blocks_T140000.npz"))
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    T = x.size
    w, integ = O.weights(x), O.integral(x)
    for i, thr in enumerate(g["thrs"]):
        st = O.boundaries(w, thr)
        assert same(st.astype(np.uint32), g[f"starts{i}_{tag}"])
        # the boundary set is exactly {0} U {t : w[t] >= thr}  
        assert same(st, np.union1d([0], np.flatnonzero(~(w < O.dt(thr)))).astype(np.uint64))
        n, s, q = O.block_stats(integ, st, T)
        assert n.sum() == T
        assert same(s, g[f"sum{i}_{tag}"]) and same(q, g[f"sumsq{i}_{tag}"])
    thr = O.auto_prior_threshold(T, O.sigma_hat(O.maxlet(x)))
    st = O.boundaries(w, thr)
    assert same(st.astype(np.uint32), g["ap_starts" + tag])
    n, s, _ = O.block_stats(integ, st, T)
    assert same(O.auto_prior(n, s), g["autoprior" + tag])


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("path", [p for p in SWEEP_CASES if "dyn" not in p], ids=os.path.basename)
def test_single_sweep_bitwise(path, fp64):
    g, x = load(path)
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    T, K = x.size, int(g["K"])
    w, integ = O.weights(x), O.integral(x)
    st = O.boundaries(w, g["thr"])
    assert same(st.astype(np.int64), g["starts"])
    n, s, q = O.block_stats(integ, st, T)
    assert same(s, g["sum" + tag]) and same(q, g["sumsq" + tag])
    if str(g["method"]) == "F":
        r = O.fb_sweep(n, s, q, g["mu"], g["var"], g["A"], g["pi"], int(g["use_self"]), g["uniforms" + tag])
        assert r["rc"] == 0
        assert same(r["rows"], g["rows" + tag])
        assert same(r["states"], g["states" + tag])
    else:
        r = O.mix_sweep(n, s, q, g["mu"], g["var"], g["uniforms" + tag])
    assert r["trans"].sum() == T and r["counts"].sum() == T  # incl. the phantom 0 -> q0 transition
    pt, pa, pp = O.posterior(r, g["tau_theta"], g["tau_A"], float(g["tau_pi"][0]))
    assert same(pt.ravel(), g["post_theta" + tag])
    assert same(pa.ravel(), g["post_A" + tag])
    assert same(pp, g["post_pi" + tag])
    seg_n, seg_s = oracle.merge_runs(r["states"], n)
    M = oracle.Marginals(T)
    M.add(seg_n, seg_s)
    assert M.text() == str(g["file_marginals" + tag])
    assert oracle.sequence_line(seg_n, seg_s) == str(g["file_sequences" + tag])
    assert "\t".join(str(int(v)) for v in n) + "\n" == str(g["file_blocks" + tag])


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("path", [p for p in SWEEP_CASES if "dyn" in p], ids=os.path.basename)
def test_multi_sweep_dynamic(path, fp64):
    """Five sweeps with dynamic blocks: theta/A/pi of sweep i+1 are the reference's own draws (taken
    from the fixture; parameter draws are libstdc++ <random>, host-side in the product), the
    threshold is re-derived from them every sweep (HMM.hpp:100-102), marginals accumulate as the
    common refinement of all recorded segmentations (StateMarginals.hpp:51-137)."""
    g, x = load(path)
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    T, K, nsw = x.size, int(g["K"]), int(g["nsweeps"])
    w, integ = O.weights(x), O.integral(x)
    drawn = g["drawn" + tag].reshape(nsw, -1)
    mu, var, A, pi = g["mu"], g["var"], g["A"], g["pi"]
    M = oracle.Marginals(T)
    seq_lines, uo, so = [], 0, 0
    for it in range(nsw):
        thr = O.threshold(T, var)
        st = O.boundaries(w, thr)
        n, s, q = O.block_stats(integ, st, T)
        B = st.size
        u = g["all_uniforms" + tag][uo:uo + B]
        uo += B
        if str(g["method"]) == "F":
            r = O.fb_sweep(n, s, q, mu, var, A, pi, 1, u, want_rows=False)
            assert same(r["states"], g["all_states" + tag][so:so + B])
            so += B
        else:
            r = O.mix_sweep(n, s, q, mu, var, u)
        seg_n, seg_s = oracle.merge_runs(r["states"], n)
        M.add(seg_n, seg_s)
        seq_lines.append(oracle.sequence_line(seg_n, seg_s))
        d = drawn[it]
        mu, var = d[0:2 * K:2].astype(O.dt), d[1:2 * K:2].astype(O.dt)
        pi, A = d[2 * K:3 * K].astype(O.dt), d[3 * K:].reshape(K, K).astype(O.dt)
    assert uo == g["all_uniforms" + tag].size
    assert "".join(seq_lines) == str(g["file_sequences" + tag])
    assert M.text() == str(g["file_marginals" + tag])


def test_nig_update_edge_cases():
    O = oracle.Oracle(False)
    rc, hp = O.nig_update([2, 1, 0, 1], 0.0, 0.0, 0)
    assert rc == 1 and list(hp) == [2, 1, 0, 1]        # warning path: unchanged (Conjugate.hpp:129-136)
    assert O.nig_update([2, 1, 0, 1], 1.0, 1.0, 0)[0] == -1   # values without a count throw
    assert O.nig_update([2, 1, 0, 1], 1.0, -1.0, 3)[0] == -1  # negative sum of squares throws
    rc, hp = O.nig_update([2, 1, 0, 1], 3.0, 2.9, 3)          # (sum^2)/N > sumSq is clamped (:152-156)
    assert rc == 0 and hp[1] >= 1.0


@pytest.mark.skipif(not (refprobe.available(True, True) and refprobe.available(True, False)),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_probe_flavours_agree():
    """The instrumented Trellis stand-in of ref_probe*_t (oracle/ref_probe.cpp) must not change what the reference
    samples: same states, same posteriors, same Records files as the probe built on the reference's own Trellis."""
    from hammlet_b200.synth import model_guess, piecewise_gaussian
    x = piecewise_gaussian(6000, 3, 80, seed=11)
    mu, var, A, pi = model_guess(3, seed=3)
    kw = dict(K=3, seed=5, theta=np.stack([mu, var], 1).ravel(), A=A.ravel(), pi=pi, thr=0.7, self=1, method="F",
              tau_theta=[2.0, 1.0, 0.0, 1.0], tau_A=[0.5, 0.5], tau_pi=0.5, nsweeps=3, dynamic=1)
    r = refprobe.run("sweep", x, fp64=True, trellis=True, **kw)
    r0 = refprobe.run("sweep", x, fp64=True, trellis=False, **kw)
    for k in ("all_states", "post_theta", "post_A", "post_pi", "drawn", "all_uniforms"):
        assert np.array_equal(r[k], r0[k]), k
    assert r["files"] == r0["files"]


@pytest.mark.skipif(not refprobe.available(False, False), reason="oracle/_ref not built (needs /root/reference)")
def test_probe_float32_file_loader_equals_text_loader():
    """bench.py's reference arm hands ref_probe a raw float32 file (rendered as text piecewise, so a 1e9-observation
    run needs no T-sized text buffer); the reference must see the same numbers as through the fp64 route."""
    from hammlet_b200.synth import piecewise_gaussian
    x = piecewise_gaussian(150_000, 5, 300, seed=3)
    a = refprobe.run("bench", x, K=5, seed=1, burn=5, timed=3, reps=1, method="F")
    b = refprobe.run("bench", x, K=5, seed=1, burn=5, timed=3, reps=1, method="F", raw32=True)
    assert a["bench_blocks"][0] == b["bench_blocks"][0] and a["sigma_hat"][0] == b["sigma_hat"][0]
