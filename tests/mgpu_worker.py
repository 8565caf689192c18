"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), one sequence split into segments.

Every rank also loads the WHOLE sequence into a second, single handle on its own GPU and checks that the
segment-split run reproduces it: weights bit for bit, block lists, sampled states (the Philox counters and
the replayed uniforms are indexed by global block number, so the partition must not change a single draw),
integer statistics exactly and fp64 statistics to 1e-12; the equal-state runs and the state marginals of recorded
sweeps, which each rank accumulates for its own positions, merge to exactly the single handle's lists.  gloo carries
the NCCL unique id.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

from hammlet_b200 import capi, gibbs  # noqa: E402
from hammlet_b200.synth import model_guess, piecewise_gaussian  # noqa: E402


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


def check_same(o, r, what):
    assert o["nblocks"] == r["nblocks"], (what, o["nblocks"], r["nblocks"])
    assert np.array_equal(o["trans"], r["trans"]), (what, o["trans"], r["trans"])
    assert np.array_equal(o["counts"], r["counts"]), what
    assert np.array_equal(o["stat_n"], r["stat_n"]), what
    assert close(o["stat_sum"], r["stat_sum"]) and close(o["stat_sq"], r["stat_sq"]), what
    assert o["fallbacks"] == r["fallbacks"], what


def local_states_match(h, ref):
    info = h.segment_info()
    full = ref.states()
    mine = h.states() if h.nr_blocks() else np.empty(0, np.int16)
    fb = info["first_block"]
    assert np.array_equal(mine, full[fb:fb + mine.size]), "sampled states differ from the single-GPU run"
    return mine.size


def runs_match(h, ref, what):
    """hml_get_segments in segment mode is collective and returns the runs of the whole sequence on every rank."""
    n, q = h.segments()
    rn, rq = ref.segments()
    assert np.array_equal(n, rn) and np.array_equal(q, rq), (what, n.size, rn.size)


def allgather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    uid = [capi.Handle.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    h = capi.Handle(dev)
    h.comm_init(rank, world, uid[0])
    expect = os.environ.get("HML_EXPECT_TRANSPORT")
    if expect:
        assert h.exchange_transport() == expect, (h.exchange_transport(), expect)
    ref = capi.Handle(dev)

    cases = [  # T, K, L, thr, use_self
        (4096 * world, 3, 50, 0.8, 1),
        (4096 * world + 1, 2, 20, 1e30, 1),       # only the forced boundaries: ranks without any block
        (65536 * world, 5, 300, 1e30, 1),
        (300_007, 5, 200, 1.0, 1),
        (1_000_003, 5, 400, 0.4, 0),
        (2_500_000, 20, 100, 0.9, 1),
        (3_000_017, 8, 2000, 1.3, 1),
    ]
    small = bool(os.environ.get("HML_MGPU_SMALL"))   # tools/gpu.sh sanitize_mgpu: the first cases only (memcheck is slow)
    if small:
        cases = cases[:4]
    for ci, (T, K, L, thr, use_self) in enumerate(cases):
        x = piecewise_gaussian(T, K, L, seed=100 + ci)
        mu, var, A, pi = model_guess(K, seed=ci)
        s, n = capi.Handle.segment_plan(T, world, rank)
        h.load_segment(x[s:s + n], T)
        ref.load(x)
        w_ref = ref.weights()
        assert np.array_equal(h.weights().view(np.uint32), w_ref[s:s + n].view(np.uint32)), f"case {ci}: weights differ"
        assert abs(h.sigma_hat() - ref.sigma_hat()) <= 1e-12 * abs(ref.sigma_hat()), f"case {ci}: sigma_hat"

        # ---- block structure
        Bg, Br = h.create_blocks(thr), ref.create_blocks(thr)
        assert Bg == Br, (ci, Bg, Br)
        st_ref, sx_ref, sq_ref = ref.blocks()
        info = h.segment_info()
        fb, nb = info["first_block"], h.nr_blocks()
        assert info["global_blocks"] == Br
        if nb:
            st, sx, sq = h.blocks()
            assert np.array_equal(st, st_ref[fb:fb + nb]), f"case {ci}: block starts differ"
            assert close(sx, sx_ref[fb:fb + nb]) and close(sq, sq_ref[fb:fb + nb]), f"case {ci}: block sums differ"

        # ---- uniform replay on the fixed structure
        u = np.random.default_rng(ci).random(Br)
        o = h.fb_sweep(mu, var, A, pi, use_self=use_self, flags=capi.SWEEP_LOGLIK, replay=u)
        r = ref.fb_sweep(mu, var, A, pi, use_self=use_self, flags=capi.SWEEP_LOGLIK, replay=u)
        check_same(o, r, f"case {ci} replay")
        assert close(o["loglik"], r["loglik"], 1e-11), (ci, o["loglik"], r["loglik"])
        local_states_match(h, ref)
        runs_match(h, ref, f"case {ci} replay")

        # ---- the same sweep with the forward filter forced either way (speculative: chunk 0 of the later ranks is repaired
        # from the last row of the rank before, one all-gather of K + 1 words; a rank with a single chunk falls back)
        for mode in (capi.FORWARD_SPECULATIVE, capi.FORWARD_OPERATORS):
            h.set_forward_mode(mode)
            o = h.fb_sweep(mu, var, A, pi, use_self=use_self, replay=u)
            check_same(o, r, f"case {ci} replay, forward mode {mode}")
            local_states_match(h, ref)
        h.set_forward_mode(capi.FORWARD_AUTO)

        # ---- mixture sweep, static structure, Philox
        o = h.mix_sweep(mu, var, A, pi, seed=5, sweep=3)
        r = ref.mix_sweep(mu, var, A, pi, seed=5, sweep=3)
        check_same(o, r, f"case {ci} mixture")
        local_states_match(h, ref)
        runs_match(h, ref, f"case {ci} mixture")

        # ---- dynamic Philox chain through the Gibbs driver: identical chains sweep by sweep
        if thr < 1e29:
            tau = gibbs.auto_prior(ref, 0.2, 0.9)
            tau_seg = gibbs.auto_prior(h, 0.2, 0.9, allgather=allgather)
            assert np.allclose(tau, tau_seg, rtol=1e-5), (tau, tau_seg)
            sa, sb = gibbs.GibbsState(K, tau, seed=9), gibbs.GibbsState(K, tau, seed=9)
            for state in (sa, sb):
                state.mean, state.var = mu.copy(), var.copy()
                state.A, state.pi = A.copy(), pi.copy()
            # every sweep is recorded: the ranks keep the marginals of their own positions, the merged view must be the
            # single handle's, entry by entry (a run that crosses a rank border is ONE segment: Records.hpp:166-188)
            h.marginals_reset(K)
            ref.marginals_reset(K)
            for i in range(6):
                oa = gibbs.sample_hmm(h, sa, 1, seed=77, sweep0=i, use_self=bool(use_self))
                ob = gibbs.sample_hmm(ref, sb, 1, seed=77, sweep0=i, use_self=bool(use_self))
                check_same(oa, ob, f"case {ci} dynamic sweep {i}")
                local_states_match(h, ref)
                if i % 2 == 0:
                    runs_match(h, ref, f"case {ci} dynamic sweep {i}")
                h.marginals_add()
                ref.marginals_add()
                if i in (0, 3, 5):
                    ms, mc, mi = h.marginals()
                    rs, rc, ri = ref.marginals()
                    assert mi == ri == i + 1 and ms.sum() == T
                    assert np.array_equal(ms, rs) and np.array_equal(mc, rc), f"case {ci} sweep {i}: merged marginals differ"
                # keep the two chains on identical parameters (the fp64 sums may differ in the last bits)
                sb.mean, sb.var, sb.A, sb.pi = sa.mean.copy(), sa.var.copy(), sa.A.copy(), sa.pi.copy()
                sb.rng.bit_generator.state = sa.rng.bit_generator.state
        if rank == 0:
            print(f"case {ci} ok: T={T} K={K} world={world} blocks={Br}", flush=True)

    # ---- multivariate data on a split sequence (-s C P D): maxlet weights = maximum over the dimensions of per-dimension
    # coefficients whose upper levels come from all-gathered tile sums; heads carry one pair of sums per dimension
    from hammlet_b200.synth import piecewise_gaussian_md
    for ci, (T, P, D, thr) in enumerate([(4096 * world + 77, 2, 2, 0.9), (150_001 if small else 900_001, 3, 2, 0.7),
                                         (70_000 if small else 400_000, 2, 3, 1e30)]):
        xm = piecewise_gaussian_md(T, P, D, 300, 40 + ci, quantum_bits=10)
        Kmd = P ** D
        mapping = np.array([[(st // (P ** d)) % P for d in range(D)] for st in range(Kmd)], np.int32)
        mu, var, _, _ = model_guess(P, seed=ci + 3)
        _, _, A, pi = model_guess(Kmd, seed=ci + 3)
        s, n = capi.Handle.segment_plan(T, world, rank)
        h.load_segment(xm[s:s + n], T)
        ref.load(xm)
        assert np.array_equal(h.weights().view(np.uint32), ref.weights()[s:s + n].view(np.uint32)), f"md case {ci}: weights differ"
        assert abs(h.sigma_hat() - ref.sigma_hat()) <= 1e-12 * abs(ref.sigma_hat())
        Bg, Br = h.create_blocks(thr), ref.create_blocks(thr)
        assert Bg == Br
        u = np.random.default_rng(ci).random(Br)
        o = h.fb_sweep(mu, var, A, pi, replay=u, mapping=mapping)
        r = ref.fb_sweep(mu, var, A, pi, replay=u, mapping=mapping)
        check_same(o, r, f"md case {ci} replay")
        local_states_match(h, ref)
        runs_match(h, ref, f"md case {ci} replay")
        if thr < 1e29:
            for i in range(3):
                o = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr * (1 + 0.05 * i), seed=5, sweep=i, mapping=mapping)
                r = ref.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=thr * (1 + 0.05 * i), seed=5, sweep=i, mapping=mapping)
                check_same(o, r, f"md case {ci} dynamic {i}")
                local_states_match(h, ref)
        if rank == 0:
            print(f"md case {ci} ok: T={T} P={P} D={D} world={world} blocks={Br}", flush=True)

    # ---- capacity growth must stay collective: tiny threshold => far more blocks than the initial capacity
    T = 200_000 if small else 2_000_000
    x = piecewise_gaussian(T, 3, 50, seed=3)
    mu, var, A, pi = model_guess(3, seed=3)
    s, n = capi.Handle.segment_plan(T, world, rank)
    h.load_segment(x[s:s + n], T)
    ref.load(x)
    o = h.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.0, seed=1, sweep=0)
    r = ref.fb_sweep(mu, var, A, pi, flags=capi.SWEEP_DYNAMIC, threshold=0.0, seed=1, sweep=0)
    check_same(o, r, "capacity growth")
    assert o["nblocks"] == T
    local_states_match(h, ref)
    if rank == 0:
        print("capacity growth ok", flush=True)

    if rank == 0:
        print("forward filter (mode, speculative sweeps, repeated):", h.forward_info(), flush=True)
    h.close()
    ref.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU WORKER OK", flush=True)


if __name__ == "__main__":
    main()
