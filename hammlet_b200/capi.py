"""ctypes binding of the C ABI in include/hammlet_b200.h (libhammlet_b200.so).

This is the Python-side mirror of the boundary: every method is one C call.  There is no CPU
fallback: if the shared library is missing or no CUDA device exists, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhammlet_b200.so")

SWEEP_DYNAMIC, SWEEP_LOGLIK, SWEEP_KEEP_ROWS, SWEEP_FUSED = 1, 2, 4, 8
DETECT_STREAM, DETECT_PYRAMID, DETECT_CANDIDATES = 0, 1, 2
FORWARD_AUTO, FORWARD_OPERATORS, FORWARD_SPECULATIVE = 0, 1, 2
MAX_STATES = 32
MAX_DIMS = 5


def combinations_mapping(P, D):
    """Mapping.hpp:89-117 (`combinations`): state s uses parameter (s // P**d) % P in dimension d; K = P**D states."""
    return np.array([[(s // P ** d) % P for d in range(D)] for s in range(P ** D)], dtype=np.int32)


class HmlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[hammlet_b200 {code}] {msg}")
        self.code = code


class _Model(C.Structure):
    _fields_ = [("K", C.c_int32), ("use_self_transitions", C.c_int32), ("mean", C.c_void_p), ("var", C.c_void_p),
                ("A", C.c_void_p), ("pi", C.c_void_p),
                # multivariate data only: handle's nr_dims, number of emission parameters, mapping[s * nr_dims + d]
                ("nr_dims", C.c_int32), ("nr_params", C.c_int32), ("mapping", C.c_void_p)]


class _SweepOut(C.Structure):
    _fields_ = [("nblocks", C.c_uint64), ("uniform_fallbacks", C.c_uint64), ("loglik", C.c_double),
                ("stat_sum", C.c_void_p), ("stat_sumsq", C.c_void_p), ("stat_n", C.c_void_p), ("trans", C.c_void_p),
                ("counts", C.c_void_p)]


EXPORTS = ["hml_create", "hml_destroy", "hml_last_error", "hml_version", "hml_load_f32", "hml_load_f32_device",
           "hml_load_f32_md", "hml_load_f32_device_md", "hml_nr_dims", "hml_get_block_sums", "hml_marginals_reset",
           "hml_marginals_add", "hml_marginals_info", "hml_marginals_get", "hml_size", "hml_sigma_hat", "hml_get_weights", "hml_get_coeffs", "hml_create_blocks", "hml_nr_blocks",
           "hml_get_blocks", "hml_fb_sweep", "hml_mix_sweep", "hml_get_states", "hml_get_segments", "hml_get_rows",
           "hml_set_timing", "hml_get_timing", "hml_launch_count", "hml_sync", "hml_get_stream",
           "hml_comm_unique_id", "hml_comm_init", "hml_segment_plan", "hml_load_segment_f32",
           "hml_load_segment_f32_device", "hml_segment_info", "hml_exchange_transport", "hml_set_detect_mode",
           "hml_detect_info",
           # include/hammlet_host.h
           "hammlet_auto_prior", "hammlet_chain_create", "hammlet_chain_destroy", "hammlet_chain_error",
           "hammlet_chain_get", "hammlet_chain_set", "hammlet_chain_run", "hammlet_chain_run_recorded",
           "hammlet_chain_save_marginals", "hammlet_chains_run", "hammlet_chain_last_sweep",
           "hml_comm_allgather", "hml_chain_init", "hml_chain_set", "hml_chain_get", "hml_chain_run",
           "hml_chain_phase_ns", "hml_load_segment_f32_md", "hml_load_segment_f32_device_md",
           "hml_set_forward_mode", "hml_forward_info"]
UNIQUE_ID_BYTES = 128

_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HmlError(-1, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.hml_last_error.restype = C.c_char_p
        lib.hml_last_error.argtypes = [C.c_void_p]
        lib.hml_version.restype = C.c_char_p
        lib.hammlet_chain_error.restype = C.c_char_p
        lib.hammlet_chain_error.argtypes = [C.c_void_p]
        lib.hammlet_chain_destroy.argtypes = [C.c_void_p]
        lib.hammlet_chain_destroy.restype = None
        for name in EXPORTS:
            getattr(lib, name)  # every declared symbol must be exported
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Handle:
    """One sequence (or shard) resident on one GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.hml_create(C.byref(self.h), C.c_int(device))
        if rc != 0:
            raise HmlError(rc, self.lib.hml_last_error(None).decode())
        self.T = 0

    def _ck(self, rc):
        if rc != 0:
            raise HmlError(rc, self.lib.hml_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.hml_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- load
    def load(self, x, weight_multiplier=1.0):
        """x: T values, or (T, D) for D-dimensional observations (position-major like the input stream)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 2:
            self._ck(self.lib.hml_load_f32_md(self.h, _ptr(x), C.c_uint64(x.shape[0]), C.c_uint32(x.shape[1]),
                                              C.c_float(weight_multiplier)))
            self.T = x.shape[0]
            return
        self._ck(self.lib.hml_load_f32(self.h, _ptr(x), C.c_uint64(x.size), C.c_float(weight_multiplier)))
        self.T = x.size

    def load_device_md(self, dev_ptr, T, nr_dims, weight_multiplier=1.0):
        self._ck(self.lib.hml_load_f32_device_md(self.h, C.c_void_p(dev_ptr), C.c_uint64(T), C.c_uint32(nr_dims),
                                                 C.c_float(weight_multiplier)))
        self.T = int(T)

    def nr_dims(self):
        d = C.c_uint32()
        self._ck(self.lib.hml_nr_dims(self.h, C.byref(d)))
        return d.value

    def load_device(self, dev_ptr, T, weight_multiplier=1.0):
        self._ck(self.lib.hml_load_f32_device(self.h, C.c_void_p(dev_ptr), C.c_uint64(T), C.c_float(weight_multiplier)))
        self.T = int(T)

    # ---- multi-GPU: one sequence split into contiguous segments (one handle per rank)
    @staticmethod
    def unique_id():
        """NCCL unique id (bytes) created by rank 0; distribute it to all ranks, then comm_init everywhere."""
        lib = load_library()
        buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
        rc = lib.hml_comm_unique_id(buf)
        if rc != 0:
            raise HmlError(rc, lib.hml_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(uid)
        self._ck(self.lib.hml_comm_init(self.h, C.c_int(rank), C.c_int(world), buf))
        self.rank, self.world = rank, world

    @staticmethod
    def segment_plan(T, world, rank):
        lib = load_library()
        s, n = C.c_uint64(), C.c_uint64()
        rc = lib.hml_segment_plan(C.c_uint64(T), C.c_int(world), C.c_int(rank), C.byref(s), C.byref(n))
        if rc != 0:
            raise HmlError(rc, "sequence too short to split: T < 4096 * world")
        return s.value, n.value

    def load_segment(self, x_local, T, weight_multiplier=1.0):
        """x_local: this rank's observations, (len,) or (len, D) for D-dimensional data; T: positions of the whole sequence."""
        x = np.ascontiguousarray(x_local, dtype=np.float32)
        if x.ndim == 2:
            self._ck(self.lib.hml_load_segment_f32_md(self.h, _ptr(x), C.c_uint64(x.shape[0]), C.c_uint64(T),
                                                      C.c_uint32(x.shape[1]), C.c_float(weight_multiplier)))
        else:
            self._ck(self.lib.hml_load_segment_f32(self.h, _ptr(x), C.c_uint64(x.size), C.c_uint64(T), C.c_float(weight_multiplier)))
        self.T = int(T)

    def load_segment_device(self, dev_ptr, n, T, weight_multiplier=1.0):
        self._ck(self.lib.hml_load_segment_f32_device(self.h, C.c_void_p(dev_ptr), C.c_uint64(n), C.c_uint64(T),
                                                      C.c_float(weight_multiplier)))
        self.T = int(T)

    def exchange_transport(self):
        """'peer' (mailboxes written over NVLink), 'nccl' (all-gathers) or 'none' (single handle)."""
        t = C.c_int()
        self._ck(self.lib.hml_exchange_transport(self.h, C.byref(t)))
        return {0: "none", 1: "peer", 2: "nccl"}[t.value]

    def segment_info(self):
        r, w = C.c_int(), C.c_int()
        s, n, fb, gb = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.hml_segment_info(self.h, C.byref(r), C.byref(w), C.byref(s), C.byref(n), C.byref(fb), C.byref(gb)))
        return dict(rank=r.value, world=w.value, seg_start=s.value, seg_len=n.value, first_block=fb.value,
                    global_blocks=gb.value)

    def sigma_hat(self):
        v = C.c_double()
        self._ck(self.lib.hml_sigma_hat(self.h, C.byref(v)))
        return v.value

    def weights(self):
        """fp32 breakpoint weights (segment mode: of the local segment)."""
        w = np.empty(self.segment_info()["seg_len"] if getattr(self, "world", 1) > 1 else self.T, dtype=np.float32)
        self._ck(self.lib.hml_get_weights(self.h, _ptr(w), C.c_uint64(w.size)))
        return w

    def coeffs(self):
        w = np.empty(self.T, dtype=np.float32)
        self._ck(self.lib.hml_get_coeffs(self.h, _ptr(w), C.c_uint64(w.size)))
        return w

    # ---- blocks
    def create_blocks(self, threshold):
        n = C.c_uint64()
        self._ck(self.lib.hml_create_blocks(self.h, C.c_float(threshold), C.byref(n)))
        return n.value

    def set_detect_mode(self, mode):
        """DETECT_STREAM (read every weight), DETECT_PYRAMID (read only sub-blocks that can hold a boundary) or
        DETECT_CANDIDATES (default: one pass over the list of positions that can be boundaries near the threshold)."""
        self._ck(self.lib.hml_set_detect_mode(self.h, C.c_int(mode)))

    def set_forward_mode(self, mode):
        """FORWARD_AUTO (default: speculative, operator scan after a failure), FORWARD_OPERATORS, FORWARD_SPECULATIVE."""
        self._ck(self.lib.hml_set_forward_mode(self.h, C.c_int(mode)))

    def forward_info(self):
        """(mode, speculative sweeps so far, how many of them were repeated through the operator scan)"""
        mode, n, f = C.c_int(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.hml_forward_info(self.h, C.byref(mode), C.byref(n), C.byref(f), None, None))
        return mode.value, n.value, f.value

    def forward_level(self):
        """(blocks per piece, warm-up blocks) the next speculative sweep would use"""
        piece, warm = C.c_int(), C.c_int()
        self._ck(self.lib.hml_forward_info(self.h, None, None, None, C.byref(piece), C.byref(warm)))
        return piece.value, warm.value

    def detect_info(self):
        mode, hot = C.c_int(), C.c_uint64()
        self._ck(self.lib.hml_detect_info(self.h, C.byref(mode), C.byref(hot)))
        return mode.value, hot.value

    def nr_blocks(self):
        n = C.c_uint64()
        self._ck(self.lib.hml_nr_blocks(self.h, C.byref(n)))
        return n.value

    def blocks(self, stats=True):
        B = self.nr_blocks()
        starts = np.empty(B, dtype=np.uint32)
        if stats:
            s, q = np.empty(B, dtype=np.float64), np.empty(B, dtype=np.float64)
            self._ck(self.lib.hml_get_blocks(self.h, _ptr(starts), _ptr(s), _ptr(q), C.c_uint64(B)))
            return starts, s, q
        self._ck(self.lib.hml_get_blocks(self.h, _ptr(starts), None, None, C.c_uint64(B)))
        return starts

    def block_sums(self, dim=0):
        """(sum x, sum x^2) per block of data dimension `dim`."""
        B = self.nr_blocks()
        s, q = np.empty(B, dtype=np.float64), np.empty(B, dtype=np.float64)
        self._ck(self.lib.hml_get_block_sums(self.h, C.c_uint32(dim), _ptr(s), _ptr(q), C.c_uint64(B)))
        return s, q

    # ---- sweeps
    def _sweep(self, fn, mean, var, A, pi, use_self, flags, threshold, seed, sweep, replay, mapping=None):
        P = len(mean)   # emission parameters; with a mapping (K, D) the model has K = mapping.shape[0] states
        mean, var = np.ascontiguousarray(mean, np.float64), np.ascontiguousarray(var, np.float64)
        if mapping is not None:
            mapping = np.ascontiguousarray(mapping, np.int32)
            K, D = mapping.shape
        else:
            K, D = P, 0
        A, pi = np.ascontiguousarray(A, np.float64).reshape(K, K), np.ascontiguousarray(pi, np.float64)
        m = _Model(K, int(use_self), _ptr(mean), _ptr(var), _ptr(A), _ptr(pi), D, P if mapping is not None else 0,
                   _ptr(mapping) if mapping is not None else None)
        ssum, ssq = np.zeros(P), np.zeros(P)
        sn, cnt, tr = np.zeros(P, np.uint64), np.zeros(K, np.uint64), np.zeros((K, K), np.uint64)
        out = _SweepOut(0, 0, 0.0, _ptr(ssum), _ptr(ssq), _ptr(sn), _ptr(tr), _ptr(cnt))
        if replay is not None:
            replay = np.ascontiguousarray(replay, np.float64)
            rp, rn = _ptr(replay), replay.size
        else:
            rp, rn = None, 0
        self._ck(fn(self.h, C.byref(m), C.c_uint32(flags), C.c_float(threshold), C.c_uint64(seed), C.c_uint64(sweep),
                    rp, C.c_uint64(rn), C.byref(out)))
        return dict(nblocks=out.nblocks, fallbacks=out.uniform_fallbacks, loglik=out.loglik, stat_sum=ssum,
                    stat_sq=ssq, stat_n=sn, trans=tr, counts=cnt)

    def fb_sweep(self, mean, var, A, pi, use_self=True, flags=0, threshold=0.0, seed=0, sweep=0, replay=None,
                 mapping=None):
        return self._sweep(self.lib.hml_fb_sweep, mean, var, A, pi, use_self, flags, threshold, seed, sweep, replay,
                           mapping)

    def mix_sweep(self, mean, var, A, pi, use_self=True, flags=0, threshold=0.0, seed=0, sweep=0, replay=None,
                  mapping=None):
        return self._sweep(self.lib.hml_mix_sweep, mean, var, A, pi, use_self, flags, threshold, seed, sweep, replay,
                           mapping)

    # ---- device-resident Gibbs chain (parameters, conjugate updates and draws on the device)
    def chain_init(self, K, prior, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=0, use_self=True):
        pr = (C.c_float * 4)(*[float(v) for v in prior])
        self._ck(self.lib.hml_chain_init(self.h, C.c_int(K), pr, C.c_float(trans), C.c_float(self_trans), C.c_float(alpha_pi),
                                         C.c_uint64(seed), C.c_int(int(use_self))))
        self.chain_K = K

    def chain_set(self, mean=None, var=None, A=None, pi=None):
        a = [None if v is None else np.ascontiguousarray(v, np.float64) for v in (mean, var, A, pi)]
        self._ck(self.lib.hml_chain_set(self.h, *[None if v is None else _ptr(v) for v in a]))

    def chain_get(self):
        K = self.chain_K
        mean, var, A, pi = np.empty(K), np.empty(K), np.empty(K * K), np.empty(K)
        thr, sweeps = C.c_float(), C.c_uint64()
        self._ck(self.lib.hml_chain_get(self.h, _ptr(mean), _ptr(var), _ptr(A), _ptr(pi), C.byref(thr), C.byref(sweeps)))
        return dict(mean=mean, var=var, A=A.reshape(K, K), pi=pi, threshold=thr.value, sweeps=sweeps.value)

    def chain_run(self, nsweeps):
        """-> statistics of the last sweep (as fb_sweep) + 'fused': sweeps that ran inside the persistent kernel."""
        K = self.chain_K
        ssum, ssq = np.zeros(K), np.zeros(K)
        sn, cnt, tr = np.zeros(K, np.uint64), np.zeros(K, np.uint64), np.zeros((K, K), np.uint64)
        out = _SweepOut(0, 0, 0.0, _ptr(ssum), _ptr(ssq), _ptr(sn), _ptr(tr), _ptr(cnt))
        fused = C.c_uint64()
        self._ck(self.lib.hml_chain_run(self.h, C.c_uint64(nsweeps), C.byref(fused), C.byref(out)))
        return dict(nblocks=out.nblocks, stat_sum=ssum, stat_sq=ssq, stat_n=sn, trans=tr, counts=cnt, fused=fused.value)

    def chain_phase_ns(self):
        st = np.zeros(16, np.uint64)
        self._ck(self.lib.hml_chain_phase_ns(self.h, _ptr(st)))
        return st

    # ---- records / debug
    def states(self):
        B = self.nr_blocks()
        s = np.empty(B, dtype=np.int16)
        self._ck(self.lib.hml_get_states(self.h, _ptr(s), C.c_uint64(B)))
        return s

    def segments(self):
        n = C.c_uint64()
        self._ck(self.lib.hml_get_segments(self.h, C.byref(n), None, None, C.c_uint64(0)))
        sizes, st = np.empty(n.value, dtype=np.uint64), np.empty(n.value, dtype=np.int16)
        self._ck(self.lib.hml_get_segments(self.h, C.byref(n), _ptr(sizes), _ptr(st), C.c_uint64(sizes.size)))
        return sizes, st

    # ---- state marginals accumulated on the device
    def marginals_reset(self, K):
        self._ck(self.lib.hml_marginals_reset(self.h, C.c_int(K)))

    def marginals_add(self):
        """The last sweep's state sequence joins the marginals (StateMarginals::addRecord for every run)."""
        self._ck(self.lib.hml_marginals_add(self.h))

    def marginals(self):
        """-> (sizes[n], counts[n, K], iterations): the common refinement of all added segmentations."""
        n, it, K = C.c_uint64(), C.c_uint64(), C.c_int()
        self._ck(self.lib.hml_marginals_info(self.h, C.byref(n), C.byref(it), C.byref(K)))
        sizes = np.empty(n.value, dtype=np.uint64)
        counts = np.empty((n.value, K.value), dtype=np.int32)
        self._ck(self.lib.hml_marginals_get(self.h, _ptr(sizes), _ptr(counts), C.c_uint64(n.value)))
        return sizes, counts, it.value

    def rows(self, K):
        B = self.nr_blocks()
        r = np.empty((B + 1, K), dtype=np.float64)
        self._ck(self.lib.hml_get_rows(self.h, _ptr(r), C.c_uint64(B + 1)))
        return r

    # ---- measurement
    def set_timing(self, on=True):
        self._ck(self.lib.hml_set_timing(self.h, C.c_int(int(on))))

    def timing(self):
        n = C.c_int()
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        self._ck(self.lib.hml_get_timing(self.h, C.byref(n), names, ms, C.c_int(32)))
        return [(names[i].decode(), float(ms[i])) for i in range(min(n.value, 32))]

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self.lib.hml_launch_count(self.h, C.byref(n)))
        return n.value

    def sync(self):
        self._ck(self.lib.hml_sync(self.h))

    def stream(self):
        p = C.c_void_p()
        self._ck(self.lib.hml_get_stream(self.h, C.byref(p)))
        return p.value


class Chain:
    """The C++ host side's Gibbs chain (include/hammlet_host.h): theta, A, pi, conjugates and the shared
    mt19937 of main.cpp, driven by sampleHMM (HMM.hpp:99-121) in C++ — no Python in the sweep loop."""

    def __init__(self, handle, K, prior, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=0):
        self.lib, self.handle, self.K = handle.lib, handle, K
        self.c = C.c_void_p()
        pr = (C.c_float * 4)(*[float(v) for v in prior])
        rc = self.lib.hammlet_chain_create(C.byref(self.c), handle.h, C.c_int(K), pr, C.c_float(trans),
                                           C.c_float(self_trans), C.c_float(alpha_pi), C.c_uint32(seed))
        if rc != 0:
            raise HmlError(rc, self.lib.hammlet_chain_error(None).decode())

    @staticmethod
    def auto_prior(handle, s2=0.2, p=0.9):
        out = (C.c_float * 4)()
        rc = handle.lib.hammlet_auto_prior(handle.h, C.c_float(s2), C.c_float(p), out)
        if rc != 0:
            raise HmlError(rc, handle.lib.hammlet_chain_error(None).decode())
        return np.array(list(out), dtype=np.float32)

    def _ck(self, rc):
        if rc != 0:
            raise HmlError(rc, self.lib.hammlet_chain_error(self.c).decode())

    def get(self):
        K = self.K
        mean, var, A, pi = (np.empty(n, np.float32) for n in (K, K, K * K, K))
        self._ck(self.lib.hammlet_chain_get(self.c, _ptr(mean), _ptr(var), _ptr(A), _ptr(pi)))
        return mean, var, A.reshape(K, K), pi

    def set(self, mean, var, A, pi):
        a = [np.ascontiguousarray(v, np.float32) for v in (mean, var, A, pi)]
        self._ck(self.lib.hammlet_chain_set(self.c, *[_ptr(v) for v in a]))

    def run(self, iterations, method="F", dynamic=True, use_self=True):
        nb = C.c_uint64()
        self._ck(self.lib.hammlet_chain_run(self.c, C.c_char(method.encode()), C.c_uint64(iterations), C.c_int(int(dynamic)),
                                            C.c_int(int(use_self)), C.byref(nb)))
        return nb.value

    def last_sweep(self):
        """Integer statistics of the chain's most recent sweep -> dict(nblocks, counts[K], trans[K, K], stat_n[K])."""
        K = self.K
        nb = C.c_uint64()
        counts, trans, stat_n = np.zeros(K, np.uint64), np.zeros(K * K, np.uint64), np.zeros(K, np.uint64)
        self._ck(self.lib.hammlet_chain_last_sweep(self.c, C.byref(nb), _ptr(counts), _ptr(trans), _ptr(stat_n)))
        return {"nblocks": nb.value, "counts": counts, "trans": trans.reshape(K, K), "stat_n": stat_n}

    def run_recorded(self, iterations, thinning=1, method="F", dynamic=True, use_self=True):
        """sampleHMM with recording: every `thinning`-th sweep joins the state marginals.  -> (blocks of the last
        sweep, marginal segments so far)"""
        nb, ns = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.hammlet_chain_run_recorded(self.c, C.c_char(method.encode()), C.c_uint64(iterations),
                                                     C.c_uint64(thinning), C.c_int(int(dynamic)), C.c_int(int(use_self)),
                                                     C.byref(nb), C.byref(ns)))
        return nb.value, ns.value

    def save_marginals(self, path):
        self._ck(self.lib.hammlet_chain_save_marginals(self.c, path.encode()))

    def close(self):
        if self.c:
            self.lib.hammlet_chain_destroy(self.c)
            self.c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_chains(chains, iterations, threads=4, method="F", dynamic=True, use_self=True):
    """hammlet_chains_run: independent chains (one per sequence), `threads` at a time on C++ host threads."""
    if not chains:
        return
    lib = chains[0].lib
    arr = (C.c_void_p * len(chains))(*[c.c for c in chains])
    rc = lib.hammlet_chains_run(arr, C.c_int(len(chains)), C.c_int(threads), C.c_char(method.encode()),
                                C.c_uint64(iterations), C.c_int(int(dynamic)), C.c_int(int(use_self)))
    if rc != 0:
        msgs = [c.lib.hammlet_chain_error(c.c).decode() for c in chains]
        raise HmlError(rc, "; ".join(m for m in msgs if m) or "a chain failed")


def philox_uniform(seed, sweep, stream, index):
    """Host restatement of hml::Philox::uniform (hml_common.cuh) for tests."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask = 0xFFFFFFFF
    c = [index & mask, (index >> 32) & mask, sweep & mask, ((sweep >> 32) & mask) ^ ((stream << 24) & mask)]
    k0, k1 = seed & mask, (seed >> 32) & mask
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> 32) ^ c[3] ^ k1) & mask, p0 & mask]
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    bits = (c[0] << 32) | c[1]
    return (bits >> 11) * (1.0 / 9007199254740992.0)
