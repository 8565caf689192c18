"""GPU parity for multivariate data (`-s C P D`): the CUDA path through the C ABI against the oracle
and the reference-made fixtures tests/golden/md_*.npz.  Same bars as the univariate path: maxlet weights and block
boundaries bit-exact, integer counts exact, sampled states identical under uniform replay, fp64 block sums / forward
rows / log-likelihood / per-parameter sums within RTOL = 1e-9 of the real_t=double reference."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle
from hammlet_b200 import capi
from hammlet_b200.synth import model_guess_md, piecewise_gaussian_md

pytestmark = pytest.mark.gpu
RTOL = 1e-9
QB = 10
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MD_CASES = sorted(glob.glob(os.path.join(GOLD, "md_*.npz")))


@pytest.fixture(scope="module")
def dev():
    h = capi.Handle(0)
    yield h
    h.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_err(a, b, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    s = np.where(s == 0, 1.0, s)
    return float(np.max(np.abs(a - b) / s)) if a.size else 0.0


@pytest.mark.parametrize("T,D", [(1, 2), (2, 3), (9, 2), (4097, 2), (65537, 3), (300_001, 5), (4096 * 4096 + 3, 2)])
def test_md_weights_bit_exact(dev, T, D):
    """deinterleave + k_maxlet_level per dimension + max over dimensions vs wavelet.hpp:97-188 (fp32, bitwise)."""
    x = piecewise_gaussian_md(T, 2, D, 80, seed=T % 1000 + D)
    O = oracle.Oracle(False)
    c_ref = O.maxlet(x)
    w_ref = O.breakpoint_weights(c_ref, 1.0)
    dev.load(x)
    assert dev.nr_dims() == D
    assert np.array_equal(bits(dev.coeffs()), bits(c_ref))
    assert np.array_equal(bits(dev.weights()), bits(w_ref))
    if T >= 2:
        assert abs(dev.sigma_hat() - O.sigma_hat(c_ref)) <= 1e-12 * abs(O.sigma_hat(c_ref))
    # a univariate load afterwards must not see leftovers of the multivariate one
    dev.load(np.ascontiguousarray(x[:, 0]))
    assert dev.nr_dims() == 1
    assert np.array_equal(bits(dev.weights()), bits(O.weights(np.ascontiguousarray(x[:, 0]))))


@pytest.mark.parametrize("T,D,L", [(5, 2, 2), (100_000, 2, 150), (1_000_003, 3, 400)])
def test_md_boundaries_and_block_sums(dev, T, D, L):
    x = piecewise_gaussian_md(T, 3, D, L, seed=T % 89 + 1)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    w = O32.weights(x)
    integ = O64.integral_md(x)
    dev.load(x)
    for thr in (0.3, 1.0, 2.5, np.inf):
        B = dev.create_blocks(thr)
        ref = O32.boundaries(w, thr)
        starts = dev.blocks(stats=False)
        assert B == ref.size and np.array_equal(starts.astype(np.uint64), ref)
        if ref.size > 300_000:
            continue
        n, rs, rq = O64.block_stats_md(integ, ref, T)
        for d in range(D):
            msq = float(np.mean(x[:, d].astype(np.float64) ** 2))
            s, q = dev.block_sums(d)
            assert rel_err(q, rq[:, d], scale=msq) <= RTOL
            assert rel_err(s, rs[:, d], scale=np.maximum(np.sqrt(n * rq[:, d]), np.sqrt(msq))) <= RTOL
    with pytest.raises(capi.HmlError):
        dev.block_sums(D)


@pytest.mark.parametrize("path", [p for p in MD_CASES if "dyn" not in p], ids=os.path.basename)
def test_md_sweep_vs_reference_fixture(dev, path):
    g = np.load(path, allow_pickle=False)
    x = g["xq"].astype(np.float32) / (1 << QB)
    T, D = x.shape
    P, K = int(g["P"]), int(g["K"])
    mapping = capi.combinations_mapping(P, D)
    assert np.array_equal(mapping, oracle.Oracle.mapping(P, D))
    dev.load(x)
    w = dev.weights()
    assert hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest() == str(g["weights_sha32"])
    B = dev.create_blocks(float(g["thr"]))
    starts = dev.blocks(stats=False)
    assert np.array_equal(starts.astype(np.int64), g["starts"])
    n = np.diff(np.append(g["starts"], T))
    for d in range(D):
        s, q = dev.block_sums(d)
        msq = float(np.mean(x[:, d].astype(np.float64) ** 2))
        assert rel_err(q, g["sumsq64"][:, d], scale=msq) <= RTOL
        assert rel_err(s, g["sum64"][:, d], scale=np.maximum(np.sqrt(n * g["sumsq64"][:, d]), np.sqrt(msq))) <= RTOL
    O64 = oracle.Oracle(True)
    u = g["uniforms64"]
    mixture = str(g["method"]) == "M"
    if mixture:
        ref = O64.mix_sweep_md(n, g["sum64"], g["sumsq64"], mapping, g["mu"], g["var"], u)
        out = dev.mix_sweep(g["mu"], g["var"], g["A"], g["pi"], replay=u, mapping=mapping)
    else:
        ref = O64.fb_sweep_md(n, g["sum64"], g["sumsq64"], mapping, g["mu"], g["var"], g["A"], g["pi"],
                              int(g["use_self"]), u)
        out = dev.fb_sweep(g["mu"], g["var"], g["A"], g["pi"], use_self=bool(g["use_self"]),
                           flags=capi.SWEEP_LOGLIK | capi.SWEEP_KEEP_ROWS, replay=u, mapping=mapping)
        assert np.array_equal(ref["states"], g["states64"])        # the oracle is pinned to the fixture
        ref_rows = g["rows64"]                                      # forward rows incl. the rescale quirk
        scale = np.maximum(ref_rows.max(axis=1, keepdims=True), 1e-300) * 1e-3
        assert rel_err(dev.rows(K), ref_rows, scale=scale) <= RTOL
        assert abs(out["loglik"] - ref["loglik"]) <= RTOL * abs(ref["loglik"])
    assert out["nblocks"] == B
    assert np.array_equal(dev.states(), ref["states"])
    assert np.array_equal(out["trans"], ref["trans"]) and np.array_equal(out["counts"], ref["counts"])
    assert np.array_equal(out["stat_n"], ref["stat_n"]) and out["stat_n"].sum() == T * D
    assert rel_err(out["stat_sum"], ref["stat_sum"], scale=1.0) <= RTOL
    assert rel_err(out["stat_sq"], ref["stat_sq"], scale=1.0) <= RTOL


def test_md_dynamic_chain_vs_reference_fixture(dev):
    """Four dynamic sweeps with the reference's own parameter draws and uniforms: identical states every sweep."""
    g = np.load(os.path.join(GOLD, "md_fb_T10000_P2_D2_dyn4.npz"), allow_pickle=False)
    x = g["xq"].astype(np.float32) / (1 << QB)
    T, D = x.shape
    P, K, nsw = int(g["P"]), int(g["K"]), int(g["nsweeps"])
    mapping = capi.combinations_mapping(P, D)
    O = oracle.Oracle(True)
    dev.load(x)
    drawn = g["drawn64"].reshape(nsw, -1)
    mu, var, A, pi = g["mu"], g["var"], g["A"], g["pi"]
    uo = so = 0
    for it in range(nsw):
        thr = oracle.Oracle(True).threshold(T, var)
        B = dev.create_blocks(np.float32(thr))
        u = g["all_uniforms64"][uo:uo + B]
        uo += B
        dev.fb_sweep(mu, var, A, pi, replay=u, mapping=mapping)
        assert np.array_equal(dev.states(), g["all_states64"][so:so + B])
        so += B
        d = drawn[it]
        mu, var = d[0:2 * P:2], d[1:2 * P:2]
        pi, A = d[2 * P:2 * P + K], d[2 * P + K:].reshape(K, K)
    assert uo == g["all_uniforms64"].size


@pytest.mark.parametrize("P,D", [(2, 4), (2, 5), (5, 2), (3, 3)])
def test_md_many_states_vs_oracle(dev, P, D):
    """K = P**D up to 32 states (the padded-state kernels 16, 20, 32), Philox and replay modes."""
    T, K = 150_000, P ** D
    x = piecewise_gaussian_md(T, P, D, 300, seed=P * 10 + D, quantum_bits=QB)
    mu, var, A, pi = model_guess_md(P, D, seed=P + D)
    mapping = capi.combinations_mapping(P, D)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    dev.load(x)
    thr = 1.0
    B = dev.create_blocks(thr)
    st = O32.boundaries(O32.weights(x), thr)
    assert B == st.size
    n, s, q = O64.block_stats_md(O64.integral_md(x), st, T)
    u = np.random.default_rng(5).random(B)
    ref = O64.fb_sweep_md(n, s, q, mapping, mu, var, A, pi, 1, u)
    out = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_LOGLIK, replay=u, mapping=mapping)
    assert np.array_equal(dev.states(), ref["states"])
    assert np.array_equal(out["trans"], ref["trans"]) and np.array_equal(out["counts"], ref["counts"])
    assert np.array_equal(out["stat_n"], ref["stat_n"])
    assert abs(out["loglik"] - ref["loglik"]) <= RTOL * abs(ref["loglik"])
    assert rel_err(out["stat_sum"], ref["stat_sum"], scale=1.0) <= RTOL
    # Philox mode, dynamic blocks: size-independent properties
    out = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=3, sweep=1, mapping=mapping)
    assert out["nblocks"] == B and out["trans"].sum() == T and out["counts"].sum() == T
    assert out["stat_n"].sum() == T * D
    tot = x.astype(np.float64).sum()
    assert abs(out["stat_sum"].sum() - tot) <= 1e-9 * max(1.0, abs(tot))


def test_md_model_errors(dev):
    x = piecewise_gaussian_md(5000, 2, 2, 100, seed=1)
    mu, var, A, pi = model_guess_md(2, 2, seed=1)
    dev.load(x)
    dev.create_blocks(1.0)
    with pytest.raises(capi.HmlError):     # multivariate data needs a mapping
        dev.fb_sweep(np.zeros(4), np.ones(4), A, pi)
    with pytest.raises(capi.HmlError):     # mapping of the wrong dimensionality
        dev.fb_sweep(mu, var, np.full((2, 2), 0.5), np.full(2, 0.5), mapping=capi.combinations_mapping(2, 1))
    bad = capi.combinations_mapping(2, 2).copy()
    bad[3, 1] = 2
    with pytest.raises(capi.HmlError):     # parameter index out of range
        dev.fb_sweep(mu, var, A, pi, mapping=bad)
    with pytest.raises(capi.HmlError):     # more dimensions than the library supports
        dev.load(np.zeros((10, capi.MAX_DIMS + 1), np.float32))
