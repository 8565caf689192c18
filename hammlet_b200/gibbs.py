"""Python mirror of the reference's Gibbs driver (HMM.hpp:60-125) above the C ABI.

Used by bench.py and the statistical tests.  The device does the per-sweep heavy lifting
(hml_fb_sweep / hml_mix_sweep); this module does the O(K^2) host part exactly where the reference does
it: conjugate Normal-Inverse-Gamma and Dirichlet updates (Conjugate.hpp:120-205) and the parameter
draws (Theta.hpp:203-211, Initial.hpp:35-40, Transitions.hpp:75-79) in real_t = float32.  Draws come
from numpy's generator (the production host, hammlet_b200/host/, uses libstdc++ <random> like the
reference), so chains agree with the reference in distribution, not draw by draw.
"""
import numpy as np

from . import capi

F = np.float32


def auto_prior(handle, s2=0.2, p=0.9, allgather=None):
    """AutoPriors.hpp:18-110 on device-computed block sums: blocks at (float)(sqrt(2 log T) sigma_hat).

    Segment mode (one sequence split over ranks): `allgather(obj) -> [obj of rank 0, ...]` collects the
    per-rank block lists, so that every rank accumulates the block means of the whole sequence in the
    reference's order and arrives at the same hyper-parameters."""
    T = handle.T
    thr = F(np.sqrt(2.0 * np.log(float(T))) * handle.sigma_hat())
    handle.create_blocks(float(thr))
    starts, s, _ = handle.blocks()
    if getattr(handle, "world", 1) > 1:
        if allgather is None:
            raise ValueError("auto_prior on a segment-split sequence needs an allgather callable")
        parts = allgather((starts, s))
        starts = np.concatenate([q[0] for q in parts])
        s = np.concatenate([q[1] for q in parts])
    n = np.diff(np.append(starts.astype(np.int64), T))
    m = (s.astype(F) / n.astype(F)).astype(F)
    # SufficientStatistics.hpp:88-91 accumulates in real_t, sequentially
    msum = np.cumsum(m, dtype=F)[-1]
    msq = np.cumsum(m * m, dtype=F)[-1]
    nb = float(m.size)
    mean = F(float(msum) / nb)
    var = F(float(msq) / nb - float(mean) * float(mean))
    b = F(-np.log(F(p)))
    sb = F(np.sqrt(b))
    M1, M2, M3 = F(0.3361), F(-0.0042), F(-0.0201)
    beta = F(float(F(s2)) * ((2.0 * float(sb)) / (float(M1 * sb) + np.sqrt(2.0) * float(F(M2 * b * F(np.exp(M3 * sb))) + F(1)))
                             + float(b)))
    if not (var > 0 and beta > 0):
        raise ValueError("Data variance provided to autoprior must be positive!")
    return np.array([2.0, beta, mean, beta / var], dtype=F)


class GibbsState:
    """theta, A, pi and their conjugate hyper-parameters for K states (univariate, identity mapping)."""

    def __init__(self, K, tau_theta, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=0):
        self.K = K
        self.rng = np.random.default_rng(seed)
        self.prior_theta = np.tile(np.asarray(tau_theta, dtype=F), (K, 1))
        self.prior_A = np.full((K, K), trans, dtype=F)
        np.fill_diagonal(self.prior_A, F(self_trans))
        self.prior_pi = np.full(K, alpha_pi, dtype=F)
        self.post_theta, self.post_A, self.post_pi = self.prior_theta.copy(), self.prior_A.copy(), self.prior_pi.copy()
        self.mean = np.zeros(K, F)
        self.var = np.ones(K, F)
        self.A = np.full((K, K), 1.0 / K, F)
        self.pi = np.full(K, 1.0 / K, F)
        self.sample_parameters()

    def threshold(self, T):
        """BreakpointArray.hpp:195-199 in fp32: sqrt(2 * log((float)T) * min var)."""
        return float(np.sqrt(F(2) * np.log(F(T)) * self.var.min(), dtype=F))

    def add_observation(self, out):
        """ForwardBackward.hpp:203-211 -> Conjugate.hpp:120-205."""
        n = out["stat_n"].astype(np.float64)
        s, q = out["stat_sum"].astype(F), out["stat_sq"].astype(F)
        a, b, mu0, nu = (self.post_theta[:, i].copy() for i in range(4))
        live = n > 0
        N = np.where(live, n, 1.0)
        xbar = (s / N).astype(F)
        ssn = np.minimum(((s * s) / N).astype(F), q)
        d = (xbar - mu0).astype(F)
        na = (a + N / 2.0).astype(F)
        nbeta = (b + ((q + (N * nu / (N + nu)) * (d * d).astype(F)) - ssn) / 2.0).astype(F)
        nmu = (((nu * mu0).astype(F) + s).astype(F) / (nu + N)).astype(F)
        nnu = (nu + N).astype(F)
        for i, new in enumerate((na, nbeta, nmu, nnu)):
            self.post_theta[:, i] = np.where(live, new, self.post_theta[:, i])
        self.post_A = (self.post_A + out["trans"].astype(F)).astype(F)
        self.post_pi = (self.post_pi + out["counts"].astype(F)).astype(F)

    def sample_parameters(self):
        """Theta::sample, Initial::sample, Transitions::sample, in the reference's order; posteriors reset."""
        r = self.rng
        a, b, mu0, nu = (self.post_theta[:, i].astype(np.float64) for i in range(4))
        var = 1.0 / r.gamma(a, 1.0 / b)
        self.var = var.astype(F)
        self.mean = r.normal(mu0, np.sqrt(self.var.astype(np.float64) / nu)).astype(F)
        g = r.gamma(self.post_pi.astype(np.float64), 1.0).astype(F)
        self.pi = (g / g.sum(dtype=F)).astype(F)
        g = r.gamma(self.post_A.astype(np.float64), 1.0).astype(F)
        self.A = (g / g.sum(axis=1, keepdims=True, dtype=F)).astype(F)
        self.post_theta[:] = self.prior_theta
        self.post_A[:] = self.prior_A
        self.post_pi[:] = self.prior_pi


def sample_hmm(handle, state, iterations, method="F", dynamic=True, use_self=True, seed=0, sweep0=0, record=None,
               thinning=0, flags=0):
    """sampleHMM (HMM.hpp:99-121).  `record(sweep_index, handle)` is called on recorded iterations."""
    T = handle.T
    outs = None
    for i in range(iterations):
        fl = flags | (capi.SWEEP_DYNAMIC if dynamic else 0)
        thr = state.threshold(T) if dynamic else 0.0
        fn = handle.fb_sweep if method == "F" else handle.mix_sweep
        outs = fn(state.mean, state.var, state.A, state.pi, use_self=use_self, flags=fl, threshold=thr, seed=seed,
                  sweep=sweep0 + i)
        state.add_observation(outs)
        state.sample_parameters()
        if record is not None and thinning > 0 and (i + 1) % thinning == 0:
            record(sweep0 + i, handle)
    return outs
