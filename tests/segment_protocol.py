"""TEST INFRASTRUCTURE — the segment-split protocol of SURVEY.md §8e.2 restated in numpy over torch.distributed.

Each rank owns a contiguous, 4096-aligned segment of one sequence.  What crosses rank boundaries per sweep is
exactly what the CUDA library all-gathers (hammlet_b200/csrc/hml_api.cu: exchange_cb / fetch_result):
  heads   {blocks of the rank, head length, head sum x, head sum x^2}: the observations in front of a rank's
          first boundary belong to the last block of the previous owner
  ops     one K x K operator per rank (product of M_t = A diag(e_t) over its blocks): the forward vector
          entering rank r is pi * Op_0 * ... * Op_{r-1}, normalised
  maps    one K -> K map per rank (composition of the per-block backward maps): the state following rank r
          is resolved through the maps of the later ranks
  stats   per-rank statistics, summed in rank order on every rank
  rows    (speculative forward filter) instead of `ops`: the last forward row of every rank + its block count
The per-block arithmetic follows the reference (ForwardBackward.hpp:64-212) in fp64; this file is the executable
statement of the protocol the kernels implement, run under gloo on CPU (tests/test_segments_gloo.py) and compared
with the unsplit oracle.
"""
import numpy as np


def discrete_draw(w, u):
    """std::discrete_distribution rule (libstdc++): normalise, partial sums, last := 1, first cp[k] >= u."""
    s = w.sum()
    if not s > 0:
        return 0
    cp = np.cumsum(w / s)
    cp[-1] = 1.0
    return int(np.searchsorted(cp, u, side="left"))


def emissions(n, sx, sq, mean, var, A, use_self):
    """e_t(s) = exp(E_s - max E) (EFD.hpp:23-38, FB.hpp:74-84) and the rescale A_ss^(N-1) (FB.hpp:115-119)."""
    N = n.astype(np.float64)[:, None]
    lognorm = np.log(np.sqrt(var)) + mean * mean / (2 * var)
    loga = np.log(np.diag(A)) if use_self else np.zeros(len(mean))
    E = (2.0 * mean * sx[:, None] - sq[:, None]) / (2.0 * var) - N * lognorm + (N - 1.0) * loga
    mx = E.max(axis=1, keepdims=True)
    return np.exp(E - mx), np.exp((N - 1.0) * loga), mx[:, 0]


def _parallel(x, y, tol=1e-13):
    l, r = x * y.sum(), y * x.sum()
    return bool(np.all(np.abs(l - r) <= tol * l + 1e-300))


def speculative_rows(gather, rank, world, heads, e, A, pi, sub, warm):
    """The speculative forward filter on a split sequence (csrc/hml_sweep_impl.cuh: spec_entry, k_fwd_fixup,
    fwd_fixup_head_cta): pieces of `sub` blocks from guessed starts, repair pass, then ONE all-gather of (last row,
    block count) and the repair of the rank's first piece from the last row of the nearest earlier rank with blocks.
    Returns (rows, failures summed over the ranks); rows are only meaningful when failures == 0."""
    B, K = e.shape
    rows = np.empty((B, K))
    fails = 0
    for first in range(0, B, sub):                               # pass 1
        if first <= warm:
            a, b0 = (np.asarray(pi, np.float64).copy() if rank == 0 else np.full(K, 1.0 / K)), 0
        else:
            a, b0 = np.full(K, 1.0 / K), first - warm
        for b in range(b0, first):
            f = (a @ A) * e[b]
            a = f / f.max() if f.max() > 0 else np.full(K, 1.0 / K)
        for b in range(first, min(B, first + sub)):
            f = (a @ A) * e[b]
            if not f.sum() > 0:
                fails += 1
                f = np.full(K, 1.0 / K)
            a = f / f.sum()
            rows[b] = a
    stored = rows.copy()

    def repair(first, entry, may_run_out):
        """-> 1 if the piece starting at `first` did not meet its guess before its last block"""
        a = entry
        steps = min(B, first + sub) - first
        for t in range(steps):
            f = (a @ A) * e[first + t]
            if not f.sum() > 0:
                return 1
            a = f / f.sum()
            met = _parallel(a, stored[first + t])
            if not met and t + 1 == steps and not may_run_out:
                return 1
            if met:
                return 0
            rows[first + t] = a
        return 0

    for first in range(sub, B, sub):                             # pass 2 (every piece but the rank's first)
        fails += repair(first, stored[first - 1], first + sub >= B)
    pub = gather((rows[B - 1].copy() if B else np.zeros(K), B))  # the rank's last row as it stands after pass 2
    if rank > 0 and B > 0:
        if B <= sub:
            fails += 1                                           # a single piece: the published row came from the guess
        else:
            src = rank - 1
            while src > 0 and not pub[src][1] > 0:
                src -= 1
            fails += repair(0, pub[src][0], False)
    return rows, sum(gather(fails))


def run_rank(dist, rank, world, seg_start, seg_len, T, starts_local, x_local, mean, var, A, pi, use_self, u_global,
             forward="operators", sub=8, warm=4):
    """One FBG sweep of this rank's segment; returns the local states and the rank-order-summed statistics.
    forward="speculative": the forward rows come from speculative_rows; if any rank reports a failure the function
    returns (None, None) on every rank — the host then repeats the sweep with forward="operators"."""
    K = len(mean)
    x64 = x_local.astype(np.float64)
    csum, csq = np.concatenate([[0.0], np.cumsum(x64)]), np.concatenate([[0.0], np.cumsum(x64 * x64)])
    B = len(starts_local)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ---- (i) heads
    e0 = int(starts_local[0]) if B else seg_len
    heads = gather((B, e0, csum[e0], csq[e0]))
    first_block = sum(h[0] for h in heads[:rank])
    global_blocks = sum(h[0] for h in heads)
    later_blocks = any(h[0] > 0 for h in heads[rank + 1:])
    ends = np.append(starts_local[1:], seg_len).astype(np.int64) if B else np.empty(0, np.int64)
    st = np.asarray(starts_local, dtype=np.int64)
    n = (ends - st).astype(np.int64)
    sx, sq = csum[ends] - csum[st], csq[ends] - csq[st]
    if B:
        for r in range(rank + 1, world):           # the last block continues up to the next boundary
            n[-1] += heads[r][1]
            sx[-1] += heads[r][2]
            sq[-1] += heads[r][3]
            if heads[r][0] > 0:
                break
    e, sp, mx = emissions(n, sx, sq, mean, var, A, use_self)

    # ---- (ii) segment operator (normalised product; scaling does not change normalised forward vectors)
    op = np.eye(K)
    for t in range(B):
        op = (op @ A) * e[t]
        op /= op.max()
    ops = gather(op)
    a = np.asarray(pi, dtype=np.float64).copy()
    for r in range(rank):
        if heads[r][0] > 0:
            a = a @ ops[r]
            a /= a.sum()
    alpha = np.empty((B, K))
    loglik = 0.0
    for t in range(B):
        f = (a @ A) * e[t]
        fs = f.sum()
        a = f / fs
        loglik += mx[t] + np.log(fs)
        alpha[t] = a
    if forward == "speculative":
        # (the operator exchange above still ran: this model keeps both so that the test can compare the rows)
        spec, failures = speculative_rows(gather, rank, world, heads, e, A, pi, sub, warm)
        if failures > 0:
            return None, None
        assert B == 0 or np.max(np.abs(spec - alpha) / np.maximum(alpha, 1e-300)) <= 1e-10, "speculative rows differ"
        alpha = spec

    # ---- (iii) backward maps, segment map, state following the segment
    maps = np.empty((B, K), dtype=np.int64)
    for t in range(B):
        gb = first_block + t
        u = u_global[global_blocks - 1 - gb]       # FB.hpp:140-162 consumes uniforms from the last block backwards
        last = (t == B - 1) and not later_blocks
        ap = alpha[t] if (last or not use_self) else alpha[t] * sp[t]
        if last:
            maps[t, :] = discrete_draw(ap, u)
        else:
            for j in range(K):
                maps[t, j] = discrete_draw(ap * A[:, j], u)
    seg_map = np.arange(K)
    for t in range(B - 1, -1, -1):                 # f_first o ... o f_last
        seg_map = maps[t][seg_map]
    all_maps = gather(seg_map)
    q_end = 0
    for r in range(world - 1, rank, -1):
        q_end = int(all_maps[r][q_end])
    states = np.empty(B, dtype=np.int16)
    q = q_end
    for t in range(B - 1, -1, -1):
        q = int(maps[t][q])
        states[t] = q

    # ---- (iv) statistics: transitions counted at their source block (FB.hpp:177-200)
    trans = np.zeros((K, K), dtype=np.uint64)
    counts = np.zeros(K, dtype=np.uint64)
    ssum, ssq = np.zeros(K), np.zeros(K)
    for t in range(B):
        s = int(states[t])
        trans[s, s] += np.uint64(n[t] - 1)
        counts[s] += np.uint64(n[t])
        ssum[s] += sx[t]
        ssq[s] += sq[t]
        if t + 1 < B:
            trans[s, int(states[t + 1])] += np.uint64(1)
        elif later_blocks:
            trans[s, q_end] += np.uint64(1)
        if t == 0 and rank == 0:
            trans[0, s] += np.uint64(1)            # phantom 0 -> q_0
    parts = gather((trans, counts, ssum, ssq, loglik))
    tot = dict(trans=sum(p[0] for p in parts), counts=sum(p[1] for p in parts),
               stat_sum=sum(p[2] for p in parts), stat_sq=sum(p[3] for p in parts), loglik=sum(p[4] for p in parts),
               nblocks=global_blocks, first_block=first_block)
    return states, tot
