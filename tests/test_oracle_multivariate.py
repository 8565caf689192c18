"""Pins the oracle's multivariate path (`-s C P D`: D data dimensions, P shared emission parameters, K = P**D states;
Mapping.hpp:89-117, wavelet.hpp:150-163, IntegralArray.hpp:136-212, EFD.hpp:83-93, ForwardBackward.hpp:189-191) to
fixtures made from the reference's own classes (oracle/make_golden.py md).  Bit-for-bit, both real_t flavours."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle

QB = 10
MD_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "md_*.npz")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(path):
    g = np.load(path, allow_pickle=False)
    return g, g["xq"].astype(np.float32) / (1 << QB)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_mapping_is_reversed_p_ary_digits():
    m = oracle.Oracle.mapping(3, 2)
    assert m.tolist() == [[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1], [0, 2], [1, 2], [2, 2]]


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("path", [p for p in MD_CASES if "dyn" not in p], ids=os.path.basename)
def test_md_single_sweep_bitwise(path, fp64):
    g, x = load(path)
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    T, D = x.shape
    P, K = int(g["P"]), int(g["K"])
    assert D == int(g["D"]) and K == P ** D
    c = O.maxlet(x)
    w = O.breakpoint_weights(c)
    assert digest(c) == str(g["coeffs_sha" + tag]) and digest(w) == str(g["weights_sha" + tag])
    assert O.sigma_hat(c) == g["sigma_hat" + tag][0]
    # a dimension alone never exceeds the maximum over dimensions
    for d in range(D):
        cd = O.maxlet(np.ascontiguousarray(x[:, d]))
        assert np.all(cd <= c)
    st = O.boundaries(w, g["thr"])
    assert same(st.astype(np.int64), g["starts"])
    integ = O.integral_md(x)
    n, s, q = O.block_stats_md(integ, st, T)
    assert same(s, g["sum" + tag]) and same(q, g["sumsq" + tag])
    mapping = O.mapping(P, D)
    if str(g["method"]) == "F":
        r = O.fb_sweep_md(n, s, q, mapping, g["mu"], g["var"], g["A"], g["pi"], int(g["use_self"]), g["uniforms" + tag])
        assert r["rc"] == 0
        assert same(r["rows"], g["rows" + tag])
        assert same(r["states"], g["states" + tag])
    else:
        r = O.mix_sweep_md(n, s, q, mapping, g["mu"], g["var"], g["uniforms" + tag])
    assert r["trans"].sum() == T and r["counts"].sum() == T
    assert r["stat_n"].sum() == T * D       # every dimension of every observation lands in exactly one parameter
    pt, pa, pp = O.posterior(r, g["tau_theta"], g["tau_A"], float(g["tau_pi"][0]))
    assert same(pt.ravel(), g["post_theta" + tag])
    assert same(pa.ravel(), g["post_A" + tag])
    assert same(pp, g["post_pi" + tag])
    seg_n, seg_s = oracle.merge_runs(r["states"], n)
    M = oracle.Marginals(T)
    M.add(seg_n, seg_s)
    assert M.text() == str(g["file_marginals" + tag])
    assert oracle.sequence_line(seg_n, seg_s) == str(g["file_sequences" + tag])
    # auto priors pool the block means of all dimensions (AutoPriors.hpp:99-104)
    thr = O.auto_prior_threshold(T, O.sigma_hat(c))
    st = O.boundaries(w, thr)
    assert same(st.astype(np.uint32), g["ap_starts" + tag])
    n, s, _ = O.block_stats_md(integ, st, T)
    assert same(O.auto_prior_md(n, s), g["autoprior" + tag])


@pytest.mark.parametrize("fp64", [False, True])
@pytest.mark.parametrize("path", [p for p in MD_CASES if "dyn" in p], ids=os.path.basename)
def test_md_multi_sweep_dynamic(path, fp64):
    g, x = load(path)
    O = oracle.Oracle(fp64)
    tag = "64" if fp64 else "32"
    T, D = x.shape
    P, K, nsw = int(g["P"]), int(g["K"]), int(g["nsweeps"])
    w, integ, mapping = O.weights(x), O.integral_md(x), O.mapping(P, D)
    drawn = g["drawn" + tag].reshape(nsw, -1)
    mu, var, A, pi = g["mu"], g["var"], g["A"], g["pi"]
    M = oracle.Marginals(T)
    seq_lines, uo, so = [], 0, 0
    for it in range(nsw):
        thr = O.threshold(T, var)           # min over the P parameters (Theta.hpp:226-234)
        st = O.boundaries(w, thr)
        n, s, q = O.block_stats_md(integ, st, T)
        B = st.size
        u = g["all_uniforms" + tag][uo:uo + B]
        uo += B
        r = O.fb_sweep_md(n, s, q, mapping, mu, var, A, pi, 1, u, want_rows=False)
        assert same(r["states"], g["all_states" + tag][so:so + B])
        so += B
        seg_n, seg_s = oracle.merge_runs(r["states"], n)
        M.add(seg_n, seg_s)
        seq_lines.append(oracle.sequence_line(seg_n, seg_s))
        d = drawn[it]
        mu, var = d[0:2 * P:2].astype(O.dt), d[1:2 * P:2].astype(O.dt)
        pi, A = d[2 * P:2 * P + K].astype(O.dt), d[2 * P + K:].reshape(K, K).astype(O.dt)
    assert uo == g["all_uniforms" + tag].size
    assert "".join(seq_lines) == str(g["file_sequences" + tag])
    assert M.text() == str(g["file_marginals" + tag])
