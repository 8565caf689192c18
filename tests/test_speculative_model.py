"""CPU model of the speculative forward filter (csrc/hml_sweep_impl.cuh: spec_entry, k_fwd_replay_prefix<kSpec>,
k_fwd_fixup), in numpy, against the sequential recursion of ForwardBackward.hpp:64-125.

What the device does, restated: every piece of `sub` blocks runs alpha_t = normalise(e_t o (alpha_{t-1} A)) from uniform
pushed through the `warm` blocks in front of it (pi from block 0 where fewer exist); the repair pass restarts every
piece from the STORED last row of its predecessor and rewrites rows until the new row is parallel to the stored one
(every component within 1e-13 relative); a piece that has not met its guess before its last block is a failure, except
the last piece.  The claim the product relies on: if no piece fails, every stored row is the sequential recursion's
(up to the tolerance) — whatever the data; and data on which the filter does not forget produces failures, not wrong
rows.
"""
import numpy as np
import pytest

TOL = 1e-13


def sequential(e, A, pi):
    a, rows = pi.copy(), np.empty_like(e)
    for t in range(e.shape[0]):
        f = (a @ A) * e[t]
        a = f / f.sum()
        rows[t] = a
    return rows


def parallel(x, y):
    l, r = x * y.sum(), y * x.sum()
    return bool(np.all(np.abs(l - r) <= TOL * l + 1e-300))


def speculative(e, A, pi, sub, warm):
    """-> (rows, failures): the two passes as the kernels run them (pass 2 reads only what pass 1 stored, plus rows of
    its own piece that it rewrote itself)."""
    B, K = e.shape
    rows = np.empty_like(e)
    for first in range(0, B, sub):                      # pass 1: guesses
        if first <= warm:
            a, b0 = pi.copy(), 0
        else:
            a, b0 = np.full(K, 1.0 / K), first - warm
        for b in range(b0, first):
            f = (a @ A) * e[b]
            a = f / f.max() if f.max() > 0 else np.full(K, 1.0 / K)
        for b in range(first, min(B, first + sub)):
            f = (a @ A) * e[b]
            a = f / f.sum()
            rows[b] = a
    stored = rows.copy()                                # what pass 2's threads read from their predecessors
    failures = 0
    for first in range(sub, B, sub):                    # pass 2: repair (independent per piece)
        a = stored[first - 1]
        steps = min(B, first + sub) - first
        last_piece = first + steps == B
        for t in range(steps):
            f = (a @ A) * e[first + t]
            a = f / f.sum()
            met = parallel(a, stored[first + t])
            if not met and t + 1 == steps and not last_piece:
                failures += 1
                break
            if met:
                break
            rows[first + t] = a
    return rows, failures


def emissions(rng, B, K, info):
    """emission terms of B blocks: one state fits each block `info` nats better than its neighbours per step away"""
    true = np.repeat(rng.integers(K, size=B // 7 + 1), 7)[:B]
    d = np.abs(np.arange(K)[None, :] - true[:, None]).astype(np.float64)
    e = np.exp(-info * d * rng.uniform(0.5, 1.5, size=(B, 1)))
    return e / e.max(axis=1, keepdims=True)


def sticky(K, leak):
    A = np.full((K, K), leak / (K - 1)) + np.eye(K) * (1.0 - leak - leak / (K - 1))
    return A / A.sum(1, keepdims=True)


@pytest.mark.parametrize("K,info,sub,warm", [(5, 40.0, 8, 4), (3, 10.0, 8, 4), (5, 6.0, 16, 16), (8, 2.5, 32, 64),
                                             (20, 4.0, 32, 8), (2, 2.5, 32, 64)])
def test_rows_are_the_sequential_ones_whenever_no_piece_fails(K, info, sub, warm):
    rng = np.random.default_rng(K * 100 + sub)
    B = 2000 + int(rng.integers(sub))                   # a ragged last piece
    e, A, pi = emissions(rng, B, K, info), sticky(K, 1e-3), rng.dirichlet(np.ones(K))
    rows, failures = speculative(e, A, pi, sub, warm)
    assert failures == 0, "these settings are meant to hold; a failure here means the case needs a longer warm-up"
    want = sequential(e, A, pi)
    assert np.max(np.abs(rows - want) / np.maximum(want, 1e-300)) <= 1e-10   # every component, however small


@pytest.mark.parametrize("K", [2, 5])
def test_a_filter_that_does_not_forget_reports_failures_not_wrong_rows(K):
    rng = np.random.default_rng(K)
    B = 1500
    e = np.ones((B, K))                                  # blocks that say nothing about the state
    A, pi = sticky(K, 1e-7), rng.dirichlet(np.ones(K))   # and transitions that almost never leave it
    _, failures = speculative(e, A, pi, 32, 64)
    assert failures > 0


def test_short_warm_ups_fail_where_long_ones_hold():
    rng = np.random.default_rng(9)
    K, B = 5, 4000
    e, A, pi = emissions(rng, B, K, 2.0), sticky(K, 1e-3), np.full(K, 0.2)
    fails = {lvl: speculative(e, A, pi, *lvl)[1] for lvl in ((8, 4), (16, 16), (32, 64))}
    assert fails[(8, 4)] > 0 and fails[(32, 64)] == 0 and fails[(16, 16)] <= fails[(8, 4)]
    rows, _ = speculative(e, A, pi, 32, 64)
    want = sequential(e, A, pi)
    assert np.max(np.abs(rows - want) / np.maximum(want, 1e-300)) <= 1e-10


def test_exact_start_near_the_beginning_of_the_sequence():
    """a piece with at most `warm` blocks in front of it starts from pi at block 0: its rows are exact before any repair"""
    rng = np.random.default_rng(4)
    K, B = 4, 64
    e, A, pi = emissions(rng, B, K, 0.05), sticky(K, 1e-4), rng.dirichlet(np.ones(K))
    rows, failures = speculative(e, A, pi, 8, 64)
    assert failures == 0 and np.allclose(rows, sequential(e, A, pi), rtol=1e-12, atol=0)
