"""CPU model of the device-side marginal merge (csrc/hml_sweep.cu: k_mg_rank -> k_seg_scan -> k_mg_write), step for
step in numpy, against oracle.Marginals (the reference's observable semantics).  It pins the ALGORITHM — index
formulas, parent segments, what counts as a new boundary — independently of the GPU tests, so that a rewrite of the
kernels can be checked on paper first."""
import numpy as np
import pytest

import oracle


def merge(P, cnt, R, rstate):
    """(P[n] sorted segment starts, cnt[n, K]) + one iteration (R[m] sorted run starts, rstate[m]) -> (P2, cnt2)."""
    n, m, K = P.size, R.size, cnt.shape[1]
    run_of_old = np.searchsorted(R, P, side="right") - 1            # k_mg_rank: upper bound - 1
    olds_below = np.searchsorted(P, R, side="left")                 # k_mg_rank: lower bound
    is_new = np.ones(m, dtype=np.int64)
    hit = olds_below < n
    is_new[hit] = (P[olds_below[hit]] != R[hit]).astype(np.int64)
    new_before = np.concatenate([[0], np.cumsum(is_new)])           # k_seg_scan: exclusive, [m] = total
    n2 = n + int(new_before[m])
    P2 = np.full(n2, -1, dtype=np.int64)
    cnt2 = np.zeros((n2, K), dtype=np.int64)
    out_old = np.arange(n) + new_before[run_of_old + 1]             # k_mg_write, old starts
    P2[out_old] = P
    cnt2[out_old] = cnt
    cnt2[out_old, rstate[run_of_old]] += 1
    j = np.flatnonzero(is_new)                                      # k_mg_write, new run starts
    out_new = olds_below[j] + new_before[j]
    P2[out_new] = R[j]
    cnt2[out_new] = cnt[olds_below[j] - 1]
    cnt2[out_new, rstate[j]] += 1
    assert np.all(P2 >= 0) and np.all(np.diff(P2) > 0)              # every slot written once, in order
    return P2, cnt2


@pytest.mark.parametrize("seed", range(6))
def test_merge_model_equals_common_refinement(seed):
    rng = np.random.default_rng(seed)
    T, K = int(rng.integers(50, 5000)), int(rng.integers(2, 7))
    P, cnt = np.zeros(1, dtype=np.int64), np.zeros((1, K), dtype=np.int64)
    M = oracle.Marginals(T)
    for it in range(12):
        nruns = int(rng.integers(1, min(T, 60) + 1))
        starts = np.sort(np.concatenate([[0], rng.choice(np.arange(1, T), size=nruns - 1, replace=False)])) if nruns > 1 else np.zeros(1, np.int64)
        states = rng.integers(0, K, size=starts.size)
        keep = np.ones(starts.size, dtype=bool)                    # maximal runs: neighbours differ
        keep[1:] = states[1:] != states[:-1]
        starts, states = starts[keep].astype(np.int64), states[keep]
        sizes = np.diff(np.append(starts, T))
        M.add(sizes, states)
        P, cnt = merge(P, cnt, starts, states)
        rs, rc = M.lines()
        assert np.array_equal(np.diff(np.append(P, T)), rs)
        assert np.array_equal(cnt[:, :rc.shape[1]], rc) and not cnt[:, rc.shape[1]:].any()
        assert np.all(cnt.sum(1) == it + 1)
