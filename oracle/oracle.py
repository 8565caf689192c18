"""TEST INFRASTRUCTURE — numpy front end of the oracle (oracle/libhammlet_oracle.so).

The oracle is a CPU restatement of the reference's hot path (see hammlet_oracle_impl.h for the
reference file:line each function follows).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product path never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhammlet_oracle.so")


def build():
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)


def _load():
    if not os.path.exists(LIB):
        build()
    return C.CDLL(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """real_t = float (fp64=False, what `hammlet` ships) or double (fp64=True, the fp64 oracle)."""

    def __init__(self, fp64=False):
        self.fp64 = fp64
        self.dt = np.float64 if fp64 else np.float32
        self.ct = C.c_double if fp64 else C.c_float
        self.sfx = "f64" if fp64 else "f32"
        self.L = lib()

    def _f(self, name, restype=None):
        f = getattr(self.L, f"{name}_{self.sfx}")
        f.restype = restype
        return f

    def _a(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=self.dt))

    # ---- load-time transforms
    def maxlet(self, x):
        """x: T values, or a (T, D) array of D-dimensional observations (position-major, as in the input file)."""
        x = self._a(x)
        if x.ndim == 2:
            out = np.empty(x.shape[0], dtype=self.dt)
            self._f("ho_maxlet_md")(_p(x), C.c_size_t(x.shape[0]), C.c_size_t(x.shape[1]), _p(out))
            return out
        out = np.empty_like(x)
        self._f("ho_maxlet")(_p(x), C.c_size_t(x.size), _p(out))
        return out

    def sigma_hat(self, coeffs):
        c = self._a(coeffs)
        return self._f("ho_sigma_hat", C.c_double)(_p(c), C.c_size_t(c.size))

    def breakpoint_weights(self, coeffs, mult=1.0):
        w = self._a(coeffs).copy()
        self._f("ho_breakpoint_weights")(_p(w), C.c_size_t(w.size), self.ct(mult))
        return w

    def weights(self, x, mult=1.0):
        return self.breakpoint_weights(self.maxlet(x), mult)

    def boundaries(self, w, thr):
        w = self._a(w)
        f = self._f("ho_boundaries", C.c_size_t)
        n = f(_p(w), C.c_size_t(w.size), self.ct(thr), None)
        starts = np.empty(n, dtype=np.uint64)
        f(_p(w), C.c_size_t(w.size), self.ct(thr), _p(starts))
        return starts

    def threshold(self, T, var):
        v = self._a(var)
        return self._f("ho_threshold", self.ct)(C.c_size_t(T), _p(v), C.c_int(v.size))

    def integral(self, x):
        x = self._a(x)
        isum = np.empty(x.size + 1, dtype=self.dt)
        isq = np.empty(x.size + 1, dtype=self.dt)
        self._f("ho_integral_build")(_p(x), C.c_size_t(x.size), _p(isum), _p(isq))
        return isum, isq

    def block_stats(self, integral, starts, T):
        isum, isq = integral
        starts = np.asarray(starts, dtype=np.uint64)
        ends = np.append(starts[1:], np.uint64(T))
        f = self._f("ho_block_stats")
        s = np.empty(starts.size, dtype=self.dt)
        q = np.empty(starts.size, dtype=self.dt)
        a, b = self.ct(), self.ct()
        for i in range(starts.size):
            f(_p(isum), _p(isq), C.c_size_t(int(starts[i])), C.c_size_t(int(ends[i])), C.byref(a), C.byref(b))
            s[i] = a.value
            q[i] = b.value
        return (ends - starts).astype(np.uint64), s, q

    # ---- multivariate data (Mapping.hpp:89-117, EFD.hpp:83-93, IntegralArray.hpp:136-212)
    @staticmethod
    def mapping(P, D):
        """mapping[s][d] = emission parameter of state s in dimension d: reversed P-ary digits of s."""
        K = P ** D
        m = np.empty((K, D), dtype=np.int32)
        for s in range(K):
            n = s
            for d in range(D):
                m[s, d] = n % P
                n //= P
        return m

    def integral_md(self, x):
        """Per-dimension integral arrays: the reference's strided cells are the univariate construction on each
        dimension's plane (IntegralArray.hpp:176-182)."""
        x = self._a(x)
        return [self.integral(np.ascontiguousarray(x[:, d])) for d in range(x.shape[1])]

    def block_stats_md(self, integrals, starts, T):
        """-> sizes[B], sum[B, D], sumsq[B, D]"""
        cols = [self.block_stats(ig, starts, T) for ig in integrals]
        return cols[0][0], np.stack([c[1] for c in cols], 1), np.stack([c[2] for c in cols], 1)

    def fb_sweep_md(self, bsize, bsum, bsq, mapping, mean, var, A, pi, use_self, uniforms, want_rows=True):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        mapping = np.ascontiguousarray(mapping, dtype=np.int32)
        K, D = mapping.shape
        B, P = bsize.size, len(mean)
        bsum, bsq, mean, var = self._a(bsum).reshape(B, D), self._a(bsq).reshape(B, D), self._a(mean), self._a(var)
        A, pi = self._a(A).reshape(K, K), self._a(pi)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        rows = np.empty((B + 1, K), dtype=self.dt) if want_rows else None
        states = np.empty(B, dtype=np.int16)
        ssum, ssq = np.empty(P, dtype=self.dt), np.empty(P, dtype=self.dt)
        sn, cnt = np.empty(P, dtype=np.uint64), np.empty(K, dtype=np.uint64)
        trans = np.empty((K, K), dtype=np.uint64)
        ll = C.c_double()
        rc = self._f("ho_fb_sweep_md", C.c_int)(
            C.c_size_t(B), _p(bsize), _p(bsum), _p(bsq), C.c_int(D), C.c_int(P), _p(mapping), C.c_int(K), _p(mean),
            _p(var), _p(A), _p(pi), C.c_int(int(use_self)), _p(u), _p(rows) if want_rows else None, _p(states),
            _p(ssum), _p(ssq), _p(sn), _p(trans), _p(cnt), C.byref(ll))
        return dict(rc=rc, rows=rows, states=states, stat_sum=ssum, stat_sq=ssq, stat_n=sn, trans=trans,
                    counts=cnt, loglik=ll.value)

    def mix_sweep_md(self, bsize, bsum, bsq, mapping, mean, var, uniforms):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        mapping = np.ascontiguousarray(mapping, dtype=np.int32)
        K, D = mapping.shape
        B, P = bsize.size, len(mean)
        bsum, bsq, mean, var = self._a(bsum).reshape(B, D), self._a(bsq).reshape(B, D), self._a(mean), self._a(var)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        states = np.empty(B, dtype=np.int16)
        ssum, ssq = np.empty(P, dtype=self.dt), np.empty(P, dtype=self.dt)
        sn, cnt = np.empty(P, dtype=np.uint64), np.empty(K, dtype=np.uint64)
        trans = np.empty((K, K), dtype=np.uint64)
        rc = self._f("ho_mix_sweep_md", C.c_int)(
            C.c_size_t(B), _p(bsize), _p(bsum), _p(bsq), C.c_int(D), C.c_int(P), _p(mapping), C.c_int(K), _p(mean),
            _p(var), _p(u), _p(states), _p(ssum), _p(ssq), _p(sn), _p(trans), _p(cnt))
        return dict(rc=rc, states=states, stat_sum=ssum, stat_sq=ssq, stat_n=sn, trans=trans, counts=cnt)

    def auto_prior_md(self, bsize, bsum, s2=0.2, p=0.9):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        bsum = self._a(bsum)
        D = bsum.shape[1] if bsum.ndim == 2 else 1
        out = np.empty(4, dtype=self.dt)
        rc = self._f("ho_auto_prior_md", C.c_int)(C.c_size_t(bsize.size), _p(bsize), _p(bsum), C.c_size_t(D),
                                                  self.ct(s2), self.ct(p), _p(out))
        if rc != 0:
            raise ValueError("auto prior rejected its inputs")
        return out

    # ---- sweeps
    def fb_sweep(self, bsize, bsum, bsq, mean, var, A, pi, use_self, uniforms, want_rows=True):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        B, K = bsize.size, len(mean)
        bsum, bsq, mean, var = self._a(bsum), self._a(bsq), self._a(mean), self._a(var)
        A, pi = self._a(A).reshape(K, K), self._a(pi)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        rows = np.empty((B + 1, K), dtype=self.dt) if want_rows else None
        states = np.empty(B, dtype=np.int16)
        ssum, ssq = np.empty(K, dtype=self.dt), np.empty(K, dtype=self.dt)
        sn, cnt = np.empty(K, dtype=np.uint64), np.empty(K, dtype=np.uint64)
        trans = np.empty((K, K), dtype=np.uint64)
        ll = C.c_double()
        rc = self._f("ho_fb_sweep", C.c_int)(
            C.c_size_t(B), _p(bsize), _p(bsum), _p(bsq), C.c_int(K), _p(mean), _p(var), _p(A), _p(pi),
            C.c_int(int(use_self)), _p(u), _p(rows) if want_rows else None, _p(states), _p(ssum), _p(ssq),
            _p(sn), _p(trans), _p(cnt), C.byref(ll))
        return dict(rc=rc, rows=rows, states=states, stat_sum=ssum, stat_sq=ssq, stat_n=sn, trans=trans,
                    counts=cnt, loglik=ll.value)

    def mix_sweep(self, bsize, bsum, bsq, mean, var, uniforms):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        B, K = bsize.size, len(mean)
        bsum, bsq, mean, var = self._a(bsum), self._a(bsq), self._a(mean), self._a(var)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        states = np.empty(B, dtype=np.int16)
        ssum, ssq = np.empty(K, dtype=self.dt), np.empty(K, dtype=self.dt)
        sn, cnt = np.empty(K, dtype=np.uint64), np.empty(K, dtype=np.uint64)
        trans = np.empty((K, K), dtype=np.uint64)
        rc = self._f("ho_mix_sweep", C.c_int)(
            C.c_size_t(B), _p(bsize), _p(bsum), _p(bsq), C.c_int(K), _p(mean), _p(var), _p(u), _p(states),
            _p(ssum), _p(ssq), _p(sn), _p(trans), _p(cnt))
        return dict(rc=rc, states=states, stat_sum=ssum, stat_sq=ssq, stat_n=sn, trans=trans, counts=cnt)

    # ---- conjugate updates
    def nig_update(self, hp, s, q, n):
        hp = self._a(hp).copy()
        rc = self._f("ho_nig_update", C.c_int)(_p(hp), self.ct(s), self.ct(q), C.c_uint64(int(n)))
        return rc, hp

    def dirichlet_update(self, alphas, counts):
        a = self._a(alphas).copy()
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        self._f("ho_dirichlet_update")(_p(a), _p(c), C.c_size_t(a.size))
        return a

    def posterior(self, res, tau_theta, tau_A, tau_pi):
        """Posterior hyper-parameters after one sweep (ForwardBackward.hpp:203-211)."""
        K = len(res["counts"])
        P = len(res["stat_n"])  # emission parameters (== K for univariate data)
        pt = np.tile(self._a(tau_theta), (P, 1))
        for s in range(P):
            if res["stat_n"][s] > 0:
                _, pt[s] = self.nig_update(pt[s], res["stat_sum"][s], res["stat_sq"][s], res["stat_n"][s])
        pa = np.full((K, K), tau_A[0], dtype=self.dt)
        np.fill_diagonal(pa, tau_A[1])
        pa = self.dirichlet_update(pa.ravel(), res["trans"].ravel()).reshape(K, K)
        pp = self.dirichlet_update(np.full(K, tau_pi, dtype=self.dt), res["counts"])
        return pt, pa, pp

    def auto_prior(self, bsize, bsum, s2=0.2, p=0.9):
        bsize = np.ascontiguousarray(bsize, dtype=np.uint64)
        bsum = self._a(bsum)
        out = np.empty(4, dtype=self.dt)
        rc = self._f("ho_auto_prior", C.c_int)(C.c_size_t(bsize.size), _p(bsize), _p(bsum), self.ct(s2),
                                               self.ct(p), _p(out))
        if rc != 0:
            raise ValueError("auto prior rejected its inputs")
        return out

    def auto_prior_threshold(self, T, sigma_hat):
        """AutoPriors.hpp:96: (real_t)(sqrt(2*log((double)T)) * noiseStdev)"""
        return self.dt(np.sqrt(2.0 * np.log(float(T))) * sigma_hat)


# ---------------------------------------------------------------------------- records / marginals


def merge_runs(states, sizes):
    """Records.hpp:155-235: adjacent equal-state blocks merge into segments (size, state)."""
    states = np.asarray(states)
    sizes = np.asarray(sizes, dtype=np.int64)
    if states.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    head = np.ones(states.size, dtype=bool)
    head[1:] = states[1:] != states[:-1]
    idx = np.flatnonzero(head)
    seg_sizes = np.add.reduceat(sizes, idx)
    return seg_sizes, states[idx].astype(np.int64)


class Marginals:
    """StateMarginals.hpp:51-137,268-310 by its observable semantics: the stored segmentation is the
    common refinement of every recorded iteration's segmentation; a line per segment with one count
    per state label 0..max_label; equal neighbours are NOT merged (:17)."""

    def __init__(self, T):
        self.T = int(T)
        self.bounds = {0}
        self.records = []
        self.max_label = -1

    def add(self, seg_sizes, seg_states):
        pos = np.concatenate([[0], np.cumsum(seg_sizes)])
        assert pos[-1] == self.T
        self.bounds.update(int(p) for p in pos[:-1])
        self.records.append((pos[:-1].copy(), np.asarray(seg_states)))
        if len(seg_states):
            self.max_label = max(self.max_label, int(np.max(seg_states)))

    def lines(self):
        b = np.array(sorted(self.bounds), dtype=np.int64)
        sizes = np.diff(np.append(b, self.T))
        S = self.max_label + 1
        counts = np.zeros((b.size, S), dtype=np.int64)
        for starts, st in self.records:
            which = np.searchsorted(starts, b, side="right") - 1
            counts[np.arange(b.size), st[which]] += 1
        return sizes, counts

    def text(self):
        sizes, counts = self.lines()
        return "".join(
            str(int(n)) + "".join("\t" + str(int(c)) for c in row) + "\n" for n, row in zip(sizes, counts))


def sequence_line(seg_sizes, seg_states):
    """Records.hpp:177-179,218-220: `size:state` tokens, TAB-separated, newline-terminated."""
    return "\t".join(f"{int(n)}:{int(s)}" for n, s in zip(seg_sizes, seg_states)) + "\n"
