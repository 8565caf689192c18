"""The speculative forward filter (hml_set_forward_mode; csrc/hml_sweep_impl.cuh: spec_entry, k_fwd_fixup) against
the oracle and against the operator scan.

The filter ForwardBackward.hpp:64-125 is one loop over the blocks.  The speculative pass runs it per chunk of 32 blocks
from a guessed start and repairs the chunk heads afterwards; it must give the reference's rows (1e-9), the
reference's states under replayed uniforms (exactly) and the same log-likelihood as the operator scan — and when the
data gives the filter no reason to forget its start (flat emissions, sticky transitions) the sweep has to notice,
fall back to the operator scan and still be right.
"""
import numpy as np
import pytest

import oracle
from hammlet_b200 import capi
from hammlet_b200.synth import model_guess, piecewise_gaussian
from test_gpu_parity import _check_fb, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    h = capi.Handle(0)
    yield h
    h.set_forward_mode(capi.FORWARD_AUTO)
    h.close()


CASES = [(3000, 3, 100, 0.9, 1.0), (50_000, 5, 500, 1.2, 1.0), (200_000, 4, 300, 0.6, 1.0), (30_000, 8, 300, 0.7, 1.0),
         (60_000, 6, 100, 0.5, 1.0), (40_000, 12, 100, 0.8, 1.0), (70_000, 20, 100, 0.6, 1.0), (30_000, 32, 100, 0.8, 1.0),
         (2_000_000, 5, 500, 0.5, 1.0), (400_000, 20, 6, 0.3, 0.3), (300_000, 5, 8, 0.3, 0.3), (1, 3, 5, 1.0, 1.0),
         (33, 3, 5, 0.1, 1.0), (1025, 2, 1, 0.01, 1.0)]


@pytest.mark.parametrize("T,K,L,thr,spacing", CASES)
def test_speculative_sweep_vs_oracle_and_operator_scan(dev, T, K, L, thr, spacing):
    x = piecewise_gaussian(T, K, L, seed=T % 89 + K, spacing=spacing)
    mu, var, A, pi = model_guess(K, seed=K, spacing=spacing)
    dev.load(x)
    dev.set_forward_mode(capi.FORWARD_OPERATORS)
    a, _ = _check_fb(dev, x, mu, var, A, pi, thr, 1)
    rows_a, states_a = dev.rows(K).copy(), dev.states().copy()
    _, n0, f0 = dev.forward_info()
    dev.set_forward_mode(capi.FORWARD_SPECULATIVE)
    b, _ = _check_fb(dev, x, mu, var, A, pi, thr, 1)
    _, n1, f1 = dev.forward_info()
    # (the first attempt uses pieces of 8 blocks behind 4 warm-up blocks: enough where a block or two pin the state down;
    # levels one sigma apart or blocks of a few observations need more — the sweep then repeats itself through the
    # operator scan and the next one would use longer pieces and warm-ups)
    easy = spacing >= 1.0 and L >= 300 and T <= 200_000
    assert n1 == n0 + 1 and f1 - f0 <= (0 if easy else 1), "the speculative pass should hold on informative data"
    assert np.array_equal(dev.states(), states_a)
    assert rel_err(dev.rows(K), rows_a, scale=1e-300) <= 1e-11   # every component, however small
    assert abs(a["loglik"] - b["loglik"]) <= 1e-12 * abs(a["loglik"])
    for k in ("trans", "counts", "stat_n"):
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("K", [2, 5, 8, 20])
def test_filter_that_does_not_forget_falls_back_to_the_operator_scan(dev, K):
    """Variances so large that no block says anything about the state, transitions that almost never leave it: the
    row at block t still remembers pi, the chunks' guesses never meet the true rows, and the sweep must be repeated
    through the operator scan — with the oracle's result."""
    T = 120_000
    x = piecewise_gaussian(T, K, 40, seed=11 + K)
    mu = np.linspace(-1.0, 1.0, K).astype(np.float32)
    var = np.full(K, 1e6, np.float32)
    A = (np.full((K, K), 1e-7) + np.eye(K) * (1.0 - K * 1e-7)).astype(np.float32)
    A = (A / A.sum(1, keepdims=True)).astype(np.float32)
    pi = (np.arange(1, K + 1) / np.arange(1, K + 1).sum()).astype(np.float32)
    dev.load(x)
    dev.set_forward_mode(capi.FORWARD_SPECULATIVE)
    _, n0, f0 = dev.forward_info()
    out, ref = _check_fb(dev, x, mu, var, A, pi, 0.5, 1)
    _, n1, f1 = dev.forward_info()
    assert out["nblocks"] > 64 and n1 == n0 + 1 and f1 == f0 + 1


def test_auto_mode_backs_off_after_a_failure_and_comes_back(dev):
    K, T = 3, 60_000
    x = piecewise_gaussian(T, K, 40, seed=5)
    flat = (np.linspace(-1, 1, K).astype(np.float32), np.full(K, 1e6, np.float32),
            (np.full((K, K), 1e-7) + np.eye(K) * (1 - 3e-7)).astype(np.float32), np.array([0.2, 0.3, 0.5], np.float32))
    sharp = model_guess(K, seed=K)
    dev.load(x)
    dev.create_blocks(0.5)
    dev.set_forward_mode(capi.FORWARD_AUTO)
    _, n0, f0 = dev.forward_info()
    for i in range(3):                                         # speculative at the three levels (pieces of 8, 16, 32 blocks): all fail
        dev.fb_sweep(*flat, use_self=1, seed=1, sweep=i)
        assert dev.forward_info()[1:] == (n0 + 1 + i, f0 + 1 + i)
    dev.fb_sweep(*flat, use_self=1, seed=1, sweep=4)           # operator scan (one sweep of back-off)
    assert dev.forward_info()[1:] == (n0 + 3, f0 + 3)
    dev.fb_sweep(*flat, use_self=1, seed=1, sweep=5)           # speculative again, fails again
    assert dev.forward_info()[1:] == (n0 + 4, f0 + 4)
    for s in range(3):                                         # three sweeps of back-off
        dev.fb_sweep(*sharp, use_self=1, seed=1, sweep=6 + s)
    assert dev.forward_info()[1:] == (n0 + 4, f0 + 4)
    dev.fb_sweep(*sharp, use_self=1, seed=1, sweep=9)          # speculative, holds
    assert dev.forward_info()[1:] == (n0 + 5, f0 + 4)
    dev.fb_sweep(*sharp, use_self=1, seed=1, sweep=10)         # and stays on
    assert dev.forward_info()[1:] == (n0 + 6, f0 + 4)


def test_longer_warm_ups_take_over_on_weakly_informative_blocks(dev):
    """Levels two sigma apart, blocks of a few observations: the first speculative sweeps do not meet their guesses, the
    following ones warm up over more blocks and hold; every sweep equals the operator scan's, states included."""
    T, K = 300_000, 5
    x = piecewise_gaussian(T, K, 8, seed=16, spacing=0.6)
    mu, var, A, pi = model_guess(K, seed=K, spacing=0.6)
    dev.load(x)
    B = dev.create_blocks(0.3)
    dev.set_forward_mode(capi.FORWARD_OPERATORS)
    want = []
    for i in range(8):
        dev.fb_sweep(mu, var, A, pi, use_self=1, seed=4, sweep=i)
        want.append(dev.states().copy())
    dev.set_forward_mode(capi.FORWARD_AUTO)
    _, n0, f0 = dev.forward_info()
    for i in range(8):
        before = dev.forward_info()
        out = dev.fb_sweep(mu, var, A, pi, use_self=1, seed=4, sweep=i)
        assert out["nblocks"] == B and np.array_equal(dev.states(), want[i]), i
    _, n1, f1 = dev.forward_info()
    assert 1 <= f1 - f0 <= 3 and n1 - n0 >= 6          # pieces of 8 and 16 blocks fail, 32 behind 64 warm-up blocks hold
    assert dev.forward_info()[1:] == (before[1] + 1, before[2])   # the last sweep was speculative and held


def test_philox_states_do_not_depend_on_the_forward_mode(dev):
    T, K = 3_000_000, 5
    x = piecewise_gaussian(T, K, 700, seed=3)
    mu, var, A, pi = model_guess(K, seed=2, stay=0.999)
    dev.load(x)
    got = {}
    for mode in (capi.FORWARD_OPERATORS, capi.FORWARD_SPECULATIVE):
        dev.set_forward_mode(mode)
        out = dev.fb_sweep(mu, var, A, pi, use_self=1, flags=capi.SWEEP_DYNAMIC, threshold=0.6, seed=9, sweep=4)
        got[mode] = (out["nblocks"], dev.states().copy(), out["trans"].copy(), out["stat_sum"].copy())
    a, b = got[capi.FORWARD_OPERATORS], got[capi.FORWARD_SPECULATIVE]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
