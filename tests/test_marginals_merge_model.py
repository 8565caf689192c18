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


# ---- a sequence split over ranks (hml_api.cu: ensure_runs with RunCtx, mg_merge_global; hml_sweep.cu: k_seg_count /
# k_seg_write in segment mode).  Every rank accumulates the marginals of its own positions; a rank > 0 always has a run
# start at its local position 0 — its first block if that block begins there, otherwise a virtual run in the previous
# rank's last state — and remembers whether that position ever started a run of the whole sequence.

def split_iteration(block_starts, block_states, borders):
    """One iteration given as blocks (global starts, states) -> per rank (local run starts, run states, border_real)."""
    T_end = borders[-1]
    out = []
    for r in range(len(borders) - 1):
        lo, hi = borders[r], borders[r + 1]
        own = np.flatnonzero((block_starts >= lo) & (block_starts < hi))      # a block belongs to the rank where it starts
        st, ss = block_starts[own] - lo, block_states[own]
        real = False
        if r == 0:
            heads = np.ones(own.size, dtype=bool)
            heads[1:] = ss[1:] != ss[:-1]
            R, S = st[heads], ss[heads]
        else:
            prev = block_states[np.searchsorted(block_starts, lo, side="left") - 1] if own.size == 0 or st[0] != 0 else \
                block_states[own[0] - 1]
            virt = own.size == 0 or st[0] != 0
            heads = np.ones(own.size, dtype=bool)
            heads[1:] = ss[1:] != ss[:-1]
            if own.size:
                heads[0] = (ss[0] != prev) if virt else True
                real = (not virt) and ss[0] != prev
            R, S = st[heads], ss[heads]
            if virt:
                R, S = np.concatenate([[0], R]), np.concatenate([[prev], S])
        out.append((R.astype(np.int64), S.astype(np.int64), real))
    assert T_end > 0
    return out


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_split_sequence_marginals_merge_to_the_whole_sequence(seed, world):
    rng = np.random.default_rng(100 * world + seed)
    T, K = int(rng.integers(40 * world, 3000)), int(rng.integers(2, 5))
    cuts = np.sort(rng.choice(np.arange(1, T), size=world - 1, replace=False))
    borders = np.concatenate([[0], cuts, [T]]).astype(np.int64)
    M = oracle.Marginals(T)
    local = [(np.zeros(1, np.int64), np.zeros((1, K), np.int64)) for _ in range(world)]
    border_real = np.zeros(world, dtype=bool)
    for it in range(10):
        nb = int(rng.integers(1, 40))
        bs = np.sort(np.concatenate([[0], rng.choice(np.arange(1, T), size=nb - 1, replace=False)])).astype(np.int64) if nb > 1 \
            else np.zeros(1, np.int64)
        if it % 3 == 0 and world > 1:                                  # a block that begins exactly at a rank border
            bs = np.unique(np.concatenate([bs, [borders[1 + it % (world - 1)]]]))
        bq = rng.integers(0, K if it % 4 else 2, size=bs.size)         # few states: many runs continue across borders
        sizes = np.diff(np.append(bs, T))
        rn, rs = oracle.merge_runs(bq, sizes)
        M.add(rn, rs)
        parts = split_iteration(bs, bq, borders)
        # the runs of the whole sequence from the ranks' lists (hml_get_segments in segment mode)
        gs, gq = [], []
        for r, (R, S, real) in enumerate(parts):
            keep = np.ones(R.size, dtype=bool)
            if r > 0 and not real:
                keep[0] = False
            gs.append(R[keep] + borders[r])
            gq.append(S[keep])
            border_real[r] |= real
            local[r] = merge(local[r][0], local[r][1], R, S)
        gs, gq = np.concatenate(gs), np.concatenate(gq)
        assert np.array_equal(np.diff(np.append(gs, T)), rn) and np.array_equal(gq, rs)
    # mg_merge_global: concatenate, joining the segment at a rank's first position to its left neighbour unless real
    starts, counts = [], []
    for r in range(world):
        P, cnt = local[r]
        for i in range(P.size):
            if i == 0 and r > 0 and not border_real[r]:
                assert np.array_equal(counts[-1], cnt[0]), "a run crossing a rank border has equal counts on both sides"
                continue
            starts.append(P[i] + borders[r])
            counts.append(cnt[i])
    sizes = np.diff(np.append(np.array(starts), T))
    rs_, rc_ = M.lines()
    counts = np.array(counts)
    assert np.array_equal(sizes, rs_)
    assert np.array_equal(counts[:, :rc_.shape[1]], rc_) and not counts[:, rc_.shape[1]:].any()
