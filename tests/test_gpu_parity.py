"""GPU parity tests: the CUDA path, called through the C ABI (hammlet_b200/capi.py -> libhammlet_b200.so),
against the oracle (oracle/, pinned to the reference by tests/test_oracle_vs_golden.py) and against the
committed golden fixtures, which hold outputs of the reference's own classes.

Bars (BASELINE.json north_star): breakpoint weights, block boundaries and every integer count are
bit-exact; fp64 block statistics, forward rows and the forward log-likelihood agree with the
real_t=double reference within RTOL = 1e-9; with identical uniforms (replay mode) the sampled state
sequence is identical.
"""
import glob
import os

import numpy as np
import pytest

import oracle
from hammlet_b200 import capi
from hammlet_b200.synth import model_guess, piecewise_gaussian

pytestmark = pytest.mark.gpu
RTOL = 1e-9
QB = 10
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    h = capi.Handle(0)
    yield h
    h.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def load_gold(name):
    g = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return g, g["xq"].astype(np.float32) / (1 << QB)


def rel_err(a, b, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    s = np.where(s == 0, 1.0, s)
    return float(np.max(np.abs(a - b) / s)) if a.size else 0.0


# ------------------------------------------------------------------------------------------ load kernels

@pytest.mark.parametrize("T", [1, 2, 3, 8, 11, 16, 1000, 4095, 4096, 4097, 8192, 65534, 65535, 65536, 65537, 200001,
                               5_000_003, 4096 * 4096 + 5])
def test_weights_bit_exact(dev, T):
    """k_maxlet_level + k_bp_weights vs wavelet.hpp:97-188 / :68-93 (fp32, bitwise, incl. forced infinities)."""
    x = piecewise_gaussian(T, 3, 50, seed=T % 1000 + 1)
    O = oracle.Oracle(False)
    c_ref = O.maxlet(x)
    w_ref = O.breakpoint_weights(c_ref, 1.0)
    dev.load(x)
    assert np.array_equal(bits(dev.coeffs()), bits(c_ref))
    assert np.array_equal(bits(dev.weights()), bits(w_ref))
    if T >= 2:
        assert abs(dev.sigma_hat() - O.sigma_hat(c_ref)) <= 1e-12 * abs(O.sigma_hat(c_ref))


def test_weight_multiplier(dev):
    x = piecewise_gaussian(30000, 3, 100, seed=3)
    O = oracle.Oracle(False)
    dev.load(x, 0.75)
    assert np.array_equal(bits(dev.weights()), bits(O.weights(x, 0.75)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "weights_T*.npz"))), ids=os.path.basename)
def test_weights_vs_reference_fixture(dev, path):
    g = np.load(path, allow_pickle=False)
    x = g["xq"].astype(np.float32) / (1 << QB)
    dev.load(x)
    w = dev.weights()
    if "weights32" in g:
        assert np.array_equal(bits(w), bits(g["weights32"]))
        assert np.array_equal(bits(dev.coeffs()), bits(g["coeffs32"]))
    else:
        import hashlib
        assert hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest() == str(g["weights_sha32"])
        assert np.array_equal(np.flatnonzero(np.isinf(w)), g["inf_pos32"])


# ------------------------------------------------------------------------------------------ boundaries + statistics

@pytest.mark.parametrize("mode", [capi.DETECT_CANDIDATES, capi.DETECT_PYRAMID, capi.DETECT_STREAM],
                         ids=["candidates", "pyramid", "stream"])
@pytest.mark.parametrize("T,L", [(1, 5), (7, 3), (4096, 40), (100_000, 200), (3_000_017, 500)])
def test_boundaries_exact_and_stats(dev, T, L, mode):
    dev.set_detect_mode(mode)
    x = piecewise_gaussian(T, 5, L, seed=T % 97 + 2)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    w = O32.weights(x)
    integ = O64.integral(x)
    # Sums come from differences of running sums (as in the reference, Statistics/IntegralArray.hpp:104-124),
    # so their absolute error is set by the running-sum magnitude; the tolerance is RTOL relative to
    # max(|value|, one average observation's contribution).
    msq = float(np.mean(x.astype(np.float64) ** 2))
    dev.load(x)
    for thr in (0.05, 0.4, 1.0, 1.7, 4.0, np.inf, np.nan, -1.0):
        B = dev.create_blocks(thr)
        ref = O32.boundaries(w, thr)
        starts, s, q = dev.blocks()
        assert B == ref.size
        assert np.array_equal(starts.astype(np.uint64), ref)          # bit-exact, ordered
        if ref.size <= 400_000:
            n, rs, rq = O64.block_stats(integ, ref, T)
            assert rel_err(q, rq, scale=msq) <= RTOL
            assert rel_err(s, rs, scale=np.maximum(np.sqrt(n * rq), np.sqrt(msq))) <= RTOL
    dev.set_detect_mode(capi.DETECT_CANDIDATES)


def test_pyramid_and_stream_detection_agree_on_hostile_weights(dev):
    """NaN / inf / huge observations give NaN and inf weights; a NaN weight is a boundary for every threshold
    (BreakpointArray.hpp:224-231 tests !(w < thr)).  Both detection modes must return the same list."""
    T = 777_777
    x = piecewise_gaussian(T, 4, 3000, seed=8)
    rng = np.random.default_rng(1)
    x[rng.integers(0, T, 40)] = np.nan
    x[rng.integers(0, T, 40)] = np.inf
    x[rng.integers(0, T, 40)] = -np.inf
    x[rng.integers(0, T, 40)] = 3e38
    dev.load(x)
    w = dev.weights()
    for thr in (0.3, 1.2, 50.0, 1e30, np.inf, np.nan):
        lists = []
        for mode in (capi.DETECT_STREAM, capi.DETECT_PYRAMID, capi.DETECT_CANDIDATES):
            dev.set_detect_mode(mode)
            B = dev.create_blocks(thr)
            lists.append(dev.blocks(stats=False))
            assert B == lists[-1].size
        expect = np.flatnonzero(~(w < np.float32(thr)))
        expect = expect if expect.size and expect[0] == 0 else np.concatenate([[0], expect])
        assert np.array_equal(lists[0], lists[1]) and np.array_equal(lists[0], lists[2])
        assert np.array_equal(lists[0].astype(np.int64), expect)
    dev.set_detect_mode(capi.DETECT_PYRAMID)
    B = dev.create_blocks(1.2)
    mode, hot = dev.detect_info()
    # the pyramid holds sub-block maxima rounded UP to bf16: every sub-block with a boundary is hot, and the
    # rounding flags only a few more
    true_hot = np.unique(dev.blocks(stats=False) // 32).size
    assert mode == capi.DETECT_PYRAMID and true_hot <= hot <= true_hot * 1.1 + 8
    dev.set_detect_mode(capi.DETECT_CANDIDATES)


def test_candidate_list_follows_the_threshold(dev):
    """Candidate mode: the list built at 0.75 x threshold serves every later threshold >= its floor, is rebuilt when
    the threshold drops below the floor or the list has become much longer than the block list, and grows the block
    arrays when it does not fit.  The boundary list must equal the oracle's for every threshold of the walk."""
    T = 6_000_011
    x = piecewise_gaussian(T, 5, 40, seed=17)
    O32 = oracle.Oracle(False)
    w = O32.weights(x)
    dev.load(x)
    assert dev.detect_info()[0] == capi.DETECT_CANDIDATES
    sizes = []
    #       build   reuse  reuse  below the floor  far above: kept once, then found too long, reuse  tiny: no list worth keeping
    for thr in (1.0, 0.9999, 1.3, 0.7, 6.0, 6.5, 7.0, 0.02, 0.021, 1.0):
        B = dev.create_blocks(thr)
        ref = O32.boundaries(w, np.float32(thr))
        assert B == ref.size
        assert np.array_equal(dev.blocks(stats=False).astype(np.uint64), ref)
        sizes.append((thr, B, dev.detect_info()[1]))
    cand = [c for _, _, c in sizes]
    assert cand[0] == cand[1] == cand[2]                      # one list served three thresholds
    assert cand[3] > cand[0]                                  # 0.7: rebuilt at a lower floor
    assert cand[4] == cand[3]                                 # 6.0: the block count that says "too long" is the previous sweep's
    assert cand[5] < cand[3] / 4 and cand[6] == cand[5]       # 6.5: rebuilt, 7.0: reused
    assert cand[7] == 0 and cand[8] == 0                      # 0.02, 0.021: the list would hold most of the sequence (> T / 2):
    #                                                           the candidate path stands down, the pyramid pass runs
    assert cand[9] == cand[0]                                 # 1.0 again: rebuilt at the floor of the first list
    assert all(c >= B for _, B, c in sizes if c)
    # a dynamic sweep uses the same path
    mu, var, A, pi = model_guess(5, seed=5)
    out = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.95, seed=1, sweep=0)
    assert out["nblocks"] == O32.boundaries(w, np.float32(0.95)).size and out["trans"].sum() == T


def test_block_stats_vs_exact_sums(dev):
    """Against exactly rounded sums (math.fsum), blocks crossing many 4096-cells included."""
    import math
    T = 300_000
    x = piecewise_gaussian(T, 3, 20000, seed=11, quantum_bits=12) + np.float32(100.0)   # large offset: hard for prefix sums
    dev.load(x)
    dev.create_blocks(2.5)
    starts, s, q = dev.blocks()
    ends = np.append(starts[1:], T)
    xd = x.astype(np.float64)
    for a, b, sv, qv in list(zip(starts, ends, s, q))[:2000]:
        es, eq = math.fsum(xd[a:b]), math.fsum(xd[a:b] ** 2)
        assert abs(sv - es) <= RTOL * abs(es) and abs(qv - eq) <= RTOL * abs(eq)


def test_blocks_vs_reference_fixture(dev):
    g, x = load_gold("blocks_T140000.npz")
    dev.load(x)
    for i, thr in enumerate(g["thrs"]):
        dev.create_blocks(float(thr))
        starts, s, q = dev.blocks()
        assert np.array_equal(starts, g[f"starts{i}_32"])             # the float reference's own block list
        assert np.array_equal(starts, g[f"starts{i}_64"])
        n = np.diff(np.append(starts, x.size))
        msq = float(np.mean(x.astype(np.float64) ** 2))
        assert rel_err(q, g[f"sumsq{i}_64"], scale=msq) <= RTOL
        assert rel_err(s, g[f"sum{i}_64"], scale=np.maximum(np.sqrt(n * g[f"sumsq{i}_64"]), np.sqrt(msq))) <= RTOL


def test_capacity_growth(dev):
    """More blocks than the initial per-block capacity (65536): buffers grow and the pass is re-run."""
    T = 1_000_000
    x = piecewise_gaussian(T, 3, 5, seed=5)
    O32 = oracle.Oracle(False)
    dev.load(x)
    B = dev.create_blocks(0.2)
    ref = O32.boundaries(O32.weights(x), 0.2)
    assert B == ref.size and B > 65536
    assert np.array_equal(dev.blocks(stats=False).astype(np.uint64), ref)


def test_reload_longer_sequence_same_k():
    """A used handle that loads a longer sequence: the K-sized per-sweep buffers must follow the new block capacity
    (they were left at the old one while the kernels indexed up to the new one).  T = 1000 gives the smallest
    capacity (65536 blocks); the second sequence has ~4x that many blocks at the same K."""
    h = capi.Handle(0)
    try:
        mu, var, A, pi = model_guess(5, seed=5)
        x0 = piecewise_gaussian(1000, 5, 50, seed=1)
        h.load(x0)
        _check_fb(h, x0, mu, var, A, pi, 0.8, 1)
        x1 = piecewise_gaussian(2_000_000, 5, 6, seed=2)
        h.load(x1)
        out, ref = _check_fb(h, x1, mu, var, A, pi, 0.3, 1)
        assert out["nblocks"] > 200_000
        # and back to a short one
        h.load(x0)
        _check_fb(h, x0, mu, var, A, pi, 0.8, 1)
    finally:
        h.close()


def test_tiny_threshold_takes_the_pyramid_pass():
    """A threshold whose candidate list would be about as long as the sequence (every weight reaches 0.75 thr): the
    candidate path stands down instead of growing the block arrays to 1.25 T, and the boundaries are still exact."""
    h = capi.Handle(0)
    try:
        T = 300_000
        x = piecewise_gaussian(T, 3, 200, seed=9)
        O32 = oracle.Oracle(False)
        w = O32.weights(x)
        h.load(x)
        for thr in (1e-30, 1e-42, float(np.nextafter(np.float32(0), np.float32(1))), 1e-3):
            B = h.create_blocks(thr)
            ref = O32.boundaries(w, thr)
            assert B == ref.size and np.array_equal(h.blocks(stats=False).astype(np.uint64), ref)
            mode, looked = h.detect_info()
            assert mode == capi.DETECT_CANDIDATES
        # a usable threshold afterwards goes back to the candidate list
        B = h.create_blocks(0.9)
        assert B == O32.boundaries(w, 0.9).size and h.detect_info()[1] < T // 8
    finally:
        h.close()


# ------------------------------------------------------------------------------------------ sweeps

def _check_fb(dev, x, mu, var, A, pi, thr, use_self, seed=123, uniforms=None):
    T, K = x.size, len(mu)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    ref_starts = O32.boundaries(O32.weights(x), thr)
    n, s, q = O64.block_stats(O64.integral(x), ref_starts, T)
    B = dev.create_blocks(thr)
    assert B == ref_starts.size
    if uniforms is None:
        uniforms = np.random.default_rng(seed).random(B)
    ref = O64.fb_sweep(n, s, q, mu, var, A, pi, use_self, uniforms)
    out = dev.fb_sweep(mu, var, A, pi, use_self=use_self, flags=capi.SWEEP_LOGLIK | capi.SWEEP_KEEP_ROWS,
                       replay=uniforms)
    assert out["nblocks"] == B and out["fallbacks"] == ref["rc"]
    rows = dev.rows(K)
    scale = np.maximum(ref["rows"].max(axis=1, keepdims=True), 1e-300) * 1e-3
    assert rel_err(rows, ref["rows"], scale=scale) <= RTOL          # forward rows incl. the rescale quirk
    assert np.array_equal(dev.states(), ref["states"])                # sample-exact under replay
    assert np.array_equal(out["trans"], ref["trans"]) and np.array_equal(out["counts"], ref["counts"])
    assert np.array_equal(out["stat_n"], ref["stat_n"])
    assert out["trans"].sum() == T
    nz = ref["stat_n"] > 0
    assert rel_err(out["stat_sq"][nz], ref["stat_sq"][nz]) <= RTOL
    assert rel_err(out["stat_sum"][nz], ref["stat_sum"][nz], scale=np.sqrt(ref["stat_n"] * ref["stat_sq"])[nz]) <= RTOL
    if ref["rc"] == 0:
        assert abs(out["loglik"] - ref["loglik"]) <= RTOL * abs(ref["loglik"])
    seg_n, seg_s = dev.segments()
    rn, rs = oracle.merge_runs(ref["states"], n)
    assert np.array_equal(seg_n.astype(np.int64), rn) and np.array_equal(seg_s.astype(np.int64), rs)
    return out, ref


@pytest.mark.parametrize("T,K,L,thr,use_self", [
    (3000, 3, 100, 0.9, 1), (50_000, 5, 500, 1.2, 1), (50_000, 5, 500, 0.4, 0), (20_000, 2, 50, 0.8, 1),
    (200_000, 4, 300, 0.6, 1), (30_000, 8, 300, 0.7, 1), (60_000, 6, 100, 0.5, 1), (40_000, 12, 100, 0.8, 1),
    (40_000, 16, 100, 0.8, 1), (70_000, 20, 100, 0.6, 1), (30_000, 32, 100, 0.8, 1), (2_000_000, 5, 500, 0.5, 1),
    (1, 3, 5, 1.0, 1), (2, 3, 5, 1.0, 1), (33, 3, 5, 0.1, 1)])
def test_fb_sweep_replay_vs_oracle(dev, T, K, L, thr, use_self):
    x = piecewise_gaussian(T, K, L, seed=T % 89 + K)
    mu, var, A, pi = model_guess(K, seed=K)
    dev.load(x)
    _check_fb(dev, x, mu, var, A, pi, thr, use_self)


@pytest.mark.parametrize("T,K,use_self", [(1_200_000, 20, 1), (600_000, 12, 1), (600_000, 16, 0), (500_000, 32, 1)])
def test_fb_sweep_many_tiles_large_k(dev, T, K, use_self):
    """K > 8 with hundreds of tiles (BASELINE configs[4] shape: level spacing 0.3 = one sigma, low compression):
    the operator tree of k_fwd_chunks_wide and the multi-CTA tile scan (k_tilescan_groups / _top / _apply).
    The emission exponents E_s grow with the spread of the levels (|E| ~ N mu^2 / sigma^2), and one ulp of E is a
    relative error |E| * 2^-53 of exp(E - max E); with levels one sigma apart the 1e-9 bar holds for 32 states."""
    x = piecewise_gaussian(T, K, 6, seed=T % 89 + K, spacing=0.3)
    mu, var, A, pi = model_guess(K, seed=K, spacing=0.3)
    dev.load(x)
    out, _ = _check_fb(dev, x, mu, var, A, pi, 0.3, use_self)
    assert out["nblocks"] > 148 * 1024      # more tiles than tile-scan groups


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "fb_T*.npz"))), ids=os.path.basename)
def test_fb_sweep_vs_reference_fixture(dev, path):
    """Against dumps of the real_t=double reference itself (rows, states, posteriors' inputs)."""
    g = np.load(path, allow_pickle=False)
    if int(g["dynamic"]):
        pytest.skip("multi-sweep fixtures are covered by test_dynamic_multi_sweep_vs_reference_fixture")
    x = g["xq"].astype(np.float32) / (1 << QB)
    K = int(g["K"])
    dev.load(x)
    B = dev.create_blocks(float(g["thr"]))
    assert np.array_equal(dev.blocks(stats=False).astype(np.int64), g["starts"])
    out = dev.fb_sweep(g["mu"], g["var"], g["A"], g["pi"], use_self=int(g["use_self"]),
                       flags=capi.SWEEP_KEEP_ROWS, replay=g["uniforms64"])
    ref_rows = g["rows64"]
    scale = np.maximum(ref_rows.max(axis=1, keepdims=True), 1e-300) * 1e-3
    assert rel_err(dev.rows(K), ref_rows, scale=scale) <= RTOL
    assert np.array_equal(dev.states(), g["states64"])
    # posterior hyper-parameters computed by the oracle's conjugate algebra from OUR statistics must
    # reproduce the reference's posteriors (Conjugate.hpp:120-205)
    O64 = oracle.Oracle(True)
    pt, pa, pp = O64.posterior(out, g["tau_theta"], g["tau_A"], float(g["tau_pi"][0]))
    assert rel_err(pt.ravel(), g["post_theta64"]) <= 1e-8
    assert np.array_equal(pa.ravel(), g["post_A64"]) and np.array_equal(pp, g["post_pi64"])


@pytest.mark.parametrize("name", ["fb_T20000_K3_dyn5.npz", "mix_T20000_K3_dyn5.npz"])
def test_dynamic_multi_sweep_vs_reference_fixture(dev, name):
    """Five dynamic sweeps driven by the reference's own parameter draws: thresholds re-derived per sweep
    in fp32 (BreakpointArray.hpp:195-199), uniforms replayed, record files reproduced byte for byte."""
    g, x = load_gold(name)
    K, nsw, T = int(g["K"]), int(g["nsweeps"]), x.size
    O32 = oracle.Oracle(False)
    dev.load(x)
    drawn = g["drawn32"].reshape(nsw, -1)
    mu, var, A, pi = g["mu"], g["var"], g["A"], g["pi"]
    M = oracle.Marginals(T)
    seq, blocks_txt, uo = [], [], 0
    for it in range(nsw):
        thr = O32.threshold(T, var.astype(np.float32))
        B = dev.create_blocks(thr)
        u = g["all_uniforms32"][uo:uo + B]
        uo += B
        fn = dev.fb_sweep if str(g["method"]) == "F" else dev.mix_sweep
        fn(mu, var, A, pi, use_self=1, replay=u)
        seg_n, seg_s = dev.segments()
        M.add(seg_n.astype(np.int64), seg_s)
        seq.append(oracle.sequence_line(seg_n, seg_s))
        starts = dev.blocks(stats=False)
        blocks_txt.append("\t".join(str(int(v)) for v in np.diff(np.append(starts, T))) + "\n")
        d = drawn[it]
        mu, var = d[0:2 * K:2].astype(np.float32), d[1:2 * K:2].astype(np.float32)
        pi, A = d[2 * K:3 * K].astype(np.float32), d[3 * K:].reshape(K, K).astype(np.float32)
    assert uo == g["all_uniforms32"].size                    # same block counts as the float reference, every sweep
    assert "".join(blocks_txt) == str(g["file_blocks32"])
    # The float reference samples from fp32 statistics that are only ~1e-3 accurate, so
    # at near-ties its states may differ from the fp64 path: require position-wise agreement >= 99.9 %.
    def expand(text):
        out = []
        for line in text.strip().split("\n"):
            toks = [t.split(":") for t in line.split("\t")]
            out.append(np.repeat([int(b) for _, b in toks], [int(a) for a, _ in toks]))
        return np.concatenate(out)
    ours, theirs = expand("".join(seq)), expand(str(g["file_sequences32"]))
    assert ours.size == theirs.size == nsw * T
    assert np.mean(ours == theirs) >= 0.999
    if np.array_equal(ours, theirs):
        assert M.text() == str(g["file_marginals32"])


def test_fb_philox_reproducible(dev):
    """Counter-based uniforms: u_b = Philox4x32-10(seed, sweep, block); feeding the same uniforms to the
    oracle (in its consumption order, last block first) reproduces the device's states."""
    T, K = 80_000, 5
    x = piecewise_gaussian(T, K, 400, seed=21)
    mu, var, A, pi = model_guess(K, seed=4)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    dev.load(x)
    thr = 0.8
    out = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=77, sweep=5)
    B = out["nblocks"]
    st = dev.states()
    u = np.array([capi.philox_uniform(77, 5, 0, b) for b in range(B)])
    starts = O32.boundaries(O32.weights(x), thr)
    assert B == starts.size
    n, s, q = O64.block_stats(O64.integral(x), starts, T)
    ref = O64.fb_sweep(n, s, q, mu, var, A, pi, 1, u[::-1].copy())
    assert np.array_equal(st, ref["states"])
    out2 = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=77, sweep=5)
    assert np.array_equal(dev.states(), st) and np.array_equal(out2["stat_sum"], out["stat_sum"])   # deterministic
    out3 = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=77, sweep=6)
    assert not np.array_equal(dev.states(), st)


@pytest.mark.parametrize("T,K,thr", [(50_000, 5, 1.2), (30_000, 3, 0.3), (40_000, 20, 0.7)])
def test_mix_sweep_vs_oracle(dev, T, K, thr):
    x = piecewise_gaussian(T, K, 300, seed=K + 40)
    mu, var, A, pi = model_guess(K, seed=K + 1)
    O32, O64 = oracle.Oracle(False), oracle.Oracle(True)
    dev.load(x)
    B = dev.create_blocks(thr)
    starts = O32.boundaries(O32.weights(x), thr)
    n, s, q = O64.block_stats(O64.integral(x), starts, T)
    u = np.random.default_rng(9).random(B)
    ref = O64.mix_sweep(n, s, q, mu, var, u)
    out = dev.mix_sweep(mu, var, A, pi, replay=u)
    assert np.array_equal(dev.states(), ref["states"])
    assert np.array_equal(out["trans"], ref["trans"]) and np.array_equal(out["counts"], ref["counts"])
    nz = ref["stat_n"] > 0
    assert rel_err(out["stat_sq"][nz], ref["stat_sq"][nz]) <= RTOL
    # Philox stream 1, block order
    out = dev.mix_sweep(mu, var, A, pi, seed=3, sweep=2)
    u = np.array([capi.philox_uniform(3, 2, 1, b) for b in range(B)])
    assert np.array_equal(dev.states(), O64.mix_sweep(n, s, q, mu, var, u)["states"])


def test_uniform_fallback_path(dev):
    """A zero forward sum (ForwardBackward.hpp:106-111) resets the filter to uniform; the device then
    re-runs the exact sequential recursion.  A = identity with well separated levels forces it."""
    K = 3
    rng = np.random.default_rng(2)
    lev = np.repeat(np.array([0, 1, 2, 0, 2, 1, 0], dtype=np.int64), 3000)
    x = (lev * 5.0 + 0.05 * rng.standard_normal(lev.size)).astype(np.float32)
    mu, var = np.array([0.0, 5.0, 10.0]), np.array([0.0025, 0.0025, 0.0025])
    A, pi = np.eye(K), np.array([0.2, 0.3, 0.5])
    dev.load(x)
    out, ref = _check_fb(dev, x, mu, var, A, pi, 2.0, 1)
    assert ref["rc"] > 0 and out["fallbacks"] == ref["rc"]


# ------------------------------------------------------------------------------------------ size-independent properties

def test_large_sequence_properties(dev):
    """BASELINE-scale shapes are checked through invariants: ordered boundaries, sizes summing to T,
    transition counts summing to T (incl. the phantom first transition), occupancy == sizes by state."""
    T, K = 40_000_000, 5
    x = piecewise_gaussian(T, K, 5000, seed=2)
    mu, var, A, pi = model_guess(K, seed=2)
    dev.load(x)
    w = dev.weights()
    thr = float(np.sqrt(2 * np.log(np.float32(T)) * var.min()))
    out = dev.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=1, sweep=0)
    starts = dev.blocks(stats=False)
    expect = np.flatnonzero(~(w < np.float32(thr)))
    assert expect[0] == 0 and np.array_equal(starts, expect.astype(np.uint32))
    sizes = np.diff(np.append(starts.astype(np.int64), T))
    st = dev.states()
    assert out["trans"].sum() == T and out["counts"].sum() == T
    assert np.array_equal(np.bincount(st, weights=sizes, minlength=K).astype(np.uint64), out["counts"])
    assert np.trace(out["trans"]) + np.count_nonzero(np.diff(st)) + (1 if st[0] != 0 else 0) == T
    seg_n, seg_s = dev.segments()
    assert seg_n.sum() == T and np.all(np.diff(seg_s) != 0)
    # idempotence: static re-run on the cached structure with the same counters gives the same result
    out2 = dev.fb_sweep(mu, var, A, pi, seed=1, sweep=0)
    assert np.array_equal(dev.states(), st) and np.array_equal(out2["trans"], out["trans"])


# ------------------------------------------------------------------------------------------------
# the C++ host side as C entry points (include/hammlet_host.h)

@pytest.mark.gpu
def test_cpp_chain_runs_sample_hmm(dev):
    from hammlet_b200 import gibbs
    from hammlet_b200.synth import piecewise_gaussian
    T, K = 400_000, 4
    x = piecewise_gaussian(T, K, 800, seed=21)
    h = capi.Handle(0)
    h.load(x)
    tau_cpp = capi.Chain.auto_prior(h, 0.2, 0.9)
    tau_py = gibbs.auto_prior(h, 0.2, 0.9)
    assert np.allclose(tau_cpp, tau_py, rtol=1e-6), (tau_cpp, tau_py)   # AutoPriors.hpp:18-110, two host mirrors
    chain = capi.Chain(h, K, tau_cpp, seed=3)
    mean0, var0, A0, pi0 = chain.get()
    assert np.all(var0 > 0) and np.allclose(A0.sum(1), 1, atol=1e-5) and abs(pi0.sum() - 1) < 1e-5
    nb = chain.run(30, method="M")        # the default scheme starts with mixture sweeps (main.cpp:57)
    nb = chain.run(60, method="F")
    mean, var, A, pi = chain.get()
    assert 0 < nb < T and np.all(np.isfinite(mean)) and np.all(var > 0)
    # the chain must have found the generating levels (spacing 1, sigma 0.3): every true level has a state near it
    levels = np.arange(K) - (K - 1) / 2.0
    assert all(np.min(np.abs(mean - lv)) < 0.15 for lv in levels), mean
    # static structure ("S" token): the block count stays what the frozen threshold gave
    nb_s = chain.run(1, method="F", dynamic=False)
    assert chain.run(5, method="F", dynamic=False) == nb_s
    # same seed, same chain
    chain2 = capi.Chain(h, K, tau_cpp, seed=3)
    chain2.run(30, method="M")
    chain2.run(60, method="F")
    assert all(np.array_equal(a, b) for a, b in zip(chain2.get(), (mean, var, A, pi)))
    chain.close()
    chain2.close()
    h.close()


def test_cpp_chain_records_marginals_from_device_runs(dev, tmp_path):
    """hammlet_chain_run_recorded: recorded sweeps hand one (size, state) entry per equal-state run to
    Records / StateMarginals (hml_get_segments forms the runs on the device).  The saved marginals must be a valid
    file of the reference's format: sizes sum to T, every line counts every recorded iteration exactly once; and
    they must match a chain that records block by block (the Python route: states + block sizes -> oracle.Marginals)."""
    from hammlet_b200.synth import piecewise_gaussian
    T, K = 300_000, 3
    x = piecewise_gaussian(T, K, 1500, seed=33)
    h = capi.Handle(0)
    h.load(x)
    tau = capi.Chain.auto_prior(h, 0.2, 0.9)
    chain = capi.Chain(h, K, tau, seed=9)
    chain.run(40, method="M")
    chain.run(40, method="F")
    nb, nseg = chain.run_recorded(30, thinning=3, method="F")
    assert 0 < nb < T and nseg >= 1
    # the last recorded sweep is the last sweep: its runs are still on the device
    sizes, states = h.segments()
    assert sizes.sum() == T and np.all(states[1:] != states[:-1])
    rs, rst = oracle.merge_runs(h.states(), np.diff(np.append(h.blocks(stats=False).astype(np.int64), T)))
    assert np.array_equal(sizes.astype(np.int64), rs) and np.array_equal(states.astype(np.int64), rst)
    path = tmp_path / "marginals.csv"
    chain.save_marginals(str(path))
    rows = [list(map(int, line.split("\t"))) for line in path.read_text().strip().split("\n")]
    assert len(rows) == nseg
    assert sum(r[0] for r in rows) == T and all(sum(r[1:]) == 10 for r in rows)
    # every boundary of the last recorded segmentation is a boundary of the common refinement
    assert set(np.cumsum(sizes)[:-1].tolist()) <= set(np.cumsum([r[0] for r in rows])[:-1].tolist())
    chain.close()
    h.close()


def test_concurrent_chains_equal_sequential_chains():
    """hammlet_chains_run: independent sequences swept several at a time (own handle, stream, RNG each) must end in
    exactly the state that the same chains reach one after the other."""
    from hammlet_b200.synth import piecewise_gaussian
    K = 3
    seqs = [piecewise_gaussian(T, K, 400, seed=40 + i) for i, T in enumerate((250_000, 90_000, 400_000, 4096, 130_001))]

    def build():
        hs, cs = [], []
        for i, x in enumerate(seqs):
            h = capi.Handle(0)
            h.load(x)
            tau = capi.Chain.auto_prior(h, 0.2, 0.9)
            hs.append(h)
            cs.append(capi.Chain(h, K, tau, seed=70 + i))
        return hs, cs

    h1, c1 = build()
    h2, c2 = build()
    for c in c1:
        c.run(15, method="M")
        c.run(25, method="F")
    capi.run_chains(c2, 15, threads=3, method="M")
    capi.run_chains(c2, 25, threads=5, method="F")
    for a, b in zip(c1, c2):
        assert all(np.array_equal(u, v) for u, v in zip(a.get(), b.get()))
    for a, b in zip(h1, h2):
        assert np.array_equal(a.states(), b.states())
    for c in c1 + c2:
        c.close()
    for h in h1 + h2:
        h.close()


def test_device_marginals_match_the_common_refinement(dev):
    """hml_marginals_add over sweeps with changing block structures and models against oracle.Marginals (the
    reference's observable semantics, pinned to its marginals files by tests/test_oracle_vs_golden.py)."""
    T, K = 500_000, 4
    x = piecewise_gaussian(T, K, 700, seed=55)
    mu, var, A, pi = model_guess(K, seed=2)
    dev.load(x)
    dev.marginals_reset(K)
    sizes, counts, it = dev.marginals()
    assert it == 0 and sizes.tolist() == [T] and counts.tolist() == [[0] * K]
    M = oracle.Marginals(T)
    rng = np.random.default_rng(1)
    for sweep, thr in enumerate((1.2, 0.5, 2.0, 0.9, 0.9, 3.5, 0.3)):
        out = dev.fb_sweep(mu * (1 + 0.02 * sweep), var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr, seed=int(rng.integers(1 << 30)),
                           sweep=sweep)
        dev.marginals_add()
        n = np.diff(np.append(dev.blocks(stats=False).astype(np.int64), T))
        M.add(*oracle.merge_runs(dev.states(), n))
        sizes, counts, it = dev.marginals()
        rs, rc = M.lines()
        assert it == sweep + 1 and sizes.sum() == T
        assert np.array_equal(sizes.astype(np.int64), rs)
        assert np.array_equal(counts[:, :rc.shape[1]], rc) and not counts[:, rc.shape[1]:].any()
    # a mixture sweep joins the same structure
    dev.mix_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=1.0, seed=5, sweep=0)
    dev.marginals_add()
    n = np.diff(np.append(dev.blocks(stats=False).astype(np.int64), T))
    M.add(*oracle.merge_runs(dev.states(), n))
    sizes, counts, it = dev.marginals()
    rs, rc = M.lines()
    assert np.array_equal(sizes.astype(np.int64), rs) and np.array_equal(counts[:, :rc.shape[1]], rc)
    assert np.all(counts.sum(1) == it)
    # loading new data starts over
    dev.load(x[:5000])
    with pytest.raises(capi.HmlError):
        dev.marginals_add()
