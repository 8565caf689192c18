"""Synthetic piecewise-constant Gaussian signals of the shapes BASELINE.json names.

levels mu_k = (k - (K-1)/2) * spacing, noise sigma, segment lengths Geometric(1/L), segment state
uniform on the K-1 states different from the previous one.  `quantum` rounds values to a multiple of
2**-quantum_bits so that the same real numbers are exact in fp32, fp64 and short decimal text.
"""
import numpy as np


def piecewise_gaussian(T, K, L, seed, spacing=1.0, sigma=0.3, quantum_bits=None, return_states=False):
    rng = np.random.default_rng(seed)
    nseg = int(T / L * 2) + 10
    lens = rng.geometric(1.0 / L, size=nseg)
    while lens.sum() < T:
        lens = np.concatenate([lens, rng.geometric(1.0 / L, size=nseg)])
    st = np.empty(lens.size, dtype=np.int64)
    st[0] = rng.integers(K)
    jump = rng.integers(K - 1, size=lens.size) if K > 1 else np.zeros(lens.size, dtype=np.int64)
    for i in range(1, lens.size):
        st[i] = jump[i] if jump[i] < st[i - 1] else jump[i] + 1
    states = np.repeat(st, lens)[:T]
    mu = (np.arange(K) - (K - 1) / 2.0) * spacing
    x = mu[states] + sigma * rng.standard_normal(T)
    if quantum_bits is None:
        x = np.round(x, 5).astype(np.float32)
    else:
        q = float(1 << quantum_bits)
        x = (np.round(x * q) / q).astype(np.float32)
    return (x, states) if return_states else x


def piecewise_gaussian_md(T, P, D, L, seed, spacing=1.0, sigma=0.3, quantum_bits=None):
    """D-dimensional observations, (T, D) position-major like the reference's input stream (wavelet.hpp:131-136):
    common change points, every dimension at one of P shared levels (the `-s C P D` model, K = P**D states)."""
    rng = np.random.default_rng(seed)
    seg = np.cumsum(rng.random(T) < 1.0 / L)
    levels = rng.integers(0, P, size=(int(seg[-1]) + 1, D))
    mu = (np.arange(P) - (P - 1) / 2.0) * spacing
    x = mu[levels[seg]] + sigma * rng.standard_normal((T, D))
    if quantum_bits is None:
        return np.round(x, 5).astype(np.float32)
    q = float(1 << quantum_bits)
    return (np.round(x * q) / q).astype(np.float32)


def model_guess_md(P, D, seed, spacing=1.0, sigma=0.3, stay=0.9):
    """(mean[P], var[P]) of the shared emission parameters and (A, pi) over the K = P**D states."""
    mu, var, _, _ = model_guess(P, seed, spacing, sigma, stay)
    _, _, A, pi = model_guess(P ** D, seed + 7, spacing, sigma, stay)
    return mu, var, A, pi


def model_guess(K, seed, spacing=1.0, sigma=0.3, stay=0.9):
    """A plausible (theta, A, pi) near the generating model, as float32 like the reference's real_t."""
    rng = np.random.default_rng(seed + 1000)
    mu = ((np.arange(K) - (K - 1) / 2.0) * spacing + 0.05 * rng.standard_normal(K)).astype(np.float32)
    var = (sigma * sigma * (1 + 0.3 * rng.random(K))).astype(np.float32)
    A = rng.dirichlet(np.full(K, 0.5), size=K) * (1 - stay) + np.eye(K) * stay
    A = (A / A.sum(1, keepdims=True)).astype(np.float32)
    pi = rng.dirichlet(np.ones(K)).astype(np.float32)
    return mu, var, A, pi


# hg38 chromosome lengths (chr1..chr22, X, Y) for the C3 workload shape
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def lpt_assign(lengths, nbins):
    """Longest-processing-time bin packing of sequences onto GPUs."""
    order = np.argsort(lengths)[::-1]
    load = [0] * nbins
    bins = [[] for _ in range(nbins)]
    for i in order:
        b = int(np.argmin(load))
        bins[b].append(int(i))
        load[b] += int(lengths[i])
    return bins
