"""TEST INFRASTRUCTURE — regenerates tests/golden/*.npz from the reference itself.

Runs oracle/_ref/ref_probe* (the reference's own classes compiled from /root/reference/src, see
oracle/Makefile) on seeded inputs and stores inputs + reference outputs as small fixtures, so that
`pytest -m "not gpu"` can pin the oracle on machines where /root/reference does not exist.
Usage (in the build container):  make -C oracle ref && python oracle/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refprobe  # noqa: E402
from hammlet_b200.synth import model_guess, model_guess_md, piecewise_gaussian, piecewise_gaussian_md  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
QB = 10  # inputs are multiples of 2**-10: exact in fp32, fp64 and decimal text


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quant(x):
    q = np.round(np.asarray(x, dtype=np.float64) * (1 << QB)).astype(np.int16)
    assert np.array_equal(q.astype(np.float32) / (1 << QB), x)
    return q


def weights_case(T, seed):
    x = piecewise_gaussian(T, 3, 50, seed, quantum_bits=QB)
    d = {"xq": quant(x), "T": T}
    for fp64, tag in ((False, "32"), (True, "64")):
        r = refprobe.run("weights", x, fp64=fp64)
        dt = np.float64 if fp64 else np.float32
        c, w = r["coeffs"].astype(dt), r["weights"].astype(dt)
        if T <= 2000:
            d["coeffs" + tag], d["weights" + tag] = c, w
        else:  # digest + every position where the weight is infinite + a strided sample
            d["coeffs_sha" + tag], d["weights_sha" + tag] = digest(c), digest(w)
            d["inf_pos" + tag] = np.flatnonzero(np.isinf(w))
            d["sample" + tag] = w[::257]
        d["sigma_hat" + tag] = r["sigma_hat"]
    return d


def sweep_case(T, K, L, seed, thr, use_self, method, nsweeps=1, dynamic=0):
    x = piecewise_gaussian(T, K, L, seed, quantum_bits=QB)
    mu, var, A, pi = model_guess(K, seed)
    tau_theta, tau_A, tau_pi = [2.0, 0.5, 0.125, 1.5], [0.5, 0.75], [0.5]
    d = dict(xq=quant(x), T=T, K=K, thr=np.float32(thr), use_self=use_self, method=method, mu=mu, var=var, A=A,
             pi=pi, tau_theta=np.array(tau_theta), tau_A=np.array(tau_A), tau_pi=np.array(tau_pi), seed=seed,
             nsweeps=nsweeps, dynamic=dynamic)
    for fp64, tag in ((False, "32"), (True, "64")):
        r = refprobe.run("sweep", x, fp64=fp64, trellis=True, K=K, seed=seed, theta=np.stack([mu, var], 1).ravel(),
                         A=A.ravel(), pi=pi, thr=thr, self=use_self, method=method, tau_theta=tau_theta,
                         tau_A=tau_A, tau_pi=tau_pi, nsweeps=nsweeps, dynamic=dynamic)
        r0 = refprobe.run("sweep", x, fp64=fp64, trellis=False, K=K, seed=seed,
                          theta=np.stack([mu, var], 1).ravel(), A=A.ravel(), pi=pi, thr=thr, self=use_self,
                          method=method, tau_theta=tau_theta, tau_A=tau_A, tau_pi=tau_pi, nsweeps=nsweeps,
                          dynamic=dynamic)
        # the instrumented Trellis double must not change what the reference samples
        for k in ("all_states", "post_theta", "post_A", "post_pi", "drawn", "all_uniforms"):
            assert np.array_equal(r[k], r0[k]), k
        assert r["files"] == r0["files"]
        dt = np.float64 if fp64 else np.float32
        d["starts"] = r["starts"]
        d["sum" + tag], d["sumsq" + tag] = r["sum"].astype(dt), r["sumsq"].astype(dt)
        d["uniforms" + tag] = r["uniforms"]
        if method == "F":
            d["rows" + tag] = r["rows"].astype(dt).reshape(-1, K)
            d["states" + tag] = r["states"].astype(np.int16)
        for k in ("post_theta", "post_A", "post_pi", "drawn"):
            d[k + tag] = r[k].astype(dt)
        d["all_states" + tag] = r["all_states"].astype(np.int16)
        d["all_uniforms" + tag] = r["all_uniforms"]
        for k, v in r["files"].items():
            d["file_" + k + tag] = np.array(v)
    return d


def md_case(T, P, D, L, seed, thr, use_self, method, nsweeps=1, dynamic=0):
    """Multivariate data (`-s C P D`): T positions x D dimensions, P shared emission parameters, K = P**D states."""
    x = piecewise_gaussian_md(T, P, D, L, seed, quantum_bits=QB)
    K = P ** D
    mu, var, A, pi = model_guess_md(P, D, seed)
    tau_theta, tau_A, tau_pi = [2.0, 0.5, 0.125, 1.5], [0.5, 0.75], [0.5]
    d = dict(xq=quant(x), T=T, P=P, D=D, K=K, thr=np.float32(thr), use_self=use_self, method=method, mu=mu, var=var,
             A=A, pi=pi, tau_theta=np.array(tau_theta), tau_A=np.array(tau_A), tau_pi=np.array(tau_pi), seed=seed,
             nsweeps=nsweeps, dynamic=dynamic)
    for fp64, tag in ((False, "32"), (True, "64")):
        kw = dict(dims=D, K=P, seed=seed, theta=np.stack([mu, var], 1).ravel(), A=A.ravel(), pi=pi, thr=thr,
                  self=use_self, method=method, tau_theta=tau_theta, tau_A=tau_A, tau_pi=tau_pi, nsweeps=nsweeps,
                  dynamic=dynamic)
        r = refprobe.run("sweep", x, fp64=fp64, trellis=True, **kw)
        r0 = refprobe.run("sweep", x, fp64=fp64, trellis=False, **kw)
        for k in ("all_states", "post_theta", "post_A", "post_pi", "drawn", "all_uniforms"):
            assert np.array_equal(r[k], r0[k]), k
        assert r["files"] == r0["files"]
        dt = np.float64 if fp64 else np.float32
        c, w = r["coeffs"].astype(dt), r["weights"].astype(dt)
        d["coeffs_sha" + tag], d["weights_sha" + tag] = digest(c), digest(w)
        d["sigma_hat" + tag] = r["sigma_hat"]
        d["starts"] = r["starts"]
        d["sum" + tag], d["sumsq" + tag] = r["sum"].astype(dt).reshape(-1, D), r["sumsq"].astype(dt).reshape(-1, D)
        d["uniforms" + tag] = r["uniforms"]
        if method == "F":
            d["rows" + tag] = r["rows"].astype(dt).reshape(-1, K)
            d["states" + tag] = r["states"].astype(np.int16)
        for k in ("post_theta", "post_A", "post_pi", "drawn"):
            d[k + tag] = r[k].astype(dt)
        d["all_states" + tag] = r["all_states"].astype(np.int16)
        d["all_uniforms" + tag] = r["all_uniforms"]
        for k, v in r["files"].items():
            d["file_" + k + tag] = np.array(v)
        ra = refprobe.run("autoprior", x, fp64=fp64, dims=D, s2=0.2, p=0.9)
        d["autoprior" + tag] = ra["autoprior"].astype(dt)
        d["ap_starts" + tag] = ra["ap_starts"].astype(np.uint32)
    return d


def blocks_case(T, seed, thrs):
    x = piecewise_gaussian(T, 3, 200, seed, quantum_bits=QB)
    d = {"xq": quant(x), "T": T, "thrs": np.array(thrs, dtype=np.float32)}
    for fp64, tag in ((False, "32"), (True, "64")):
        r = refprobe.run("blocks", x, fp64=fp64, thr=thrs)
        dt = np.float64 if fp64 else np.float32
        for i in range(len(thrs)):
            d[f"starts{i}_{tag}"] = r[f"t{i}_starts"].astype(np.uint32)
            d[f"sum{i}_{tag}"] = r[f"t{i}_sum"].astype(dt)
            d[f"sumsq{i}_{tag}"] = r[f"t{i}_sumsq"].astype(dt)
        r = refprobe.run("autoprior", x, fp64=fp64, s2=0.2, p=0.9)
        d["autoprior" + tag] = r["autoprior"].astype(dt)
        d["ap_starts" + tag] = r["ap_starts"].astype(np.uint32)
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    only_md = len(sys.argv) > 1 and sys.argv[1] == "md"
    if not only_md:  # `python oracle/make_golden.py md` regenerates the multivariate fixtures only
        for T in (1, 2, 3, 8, 11, 16, 1000, 65534, 65535, 65536, 65537, 131071):
            np.savez_compressed(os.path.join(OUT, f"weights_T{T}.npz"), **weights_case(T, seed=T))
        np.savez_compressed(os.path.join(OUT, "blocks_T140000.npz"), **blocks_case(140000, 5, [0.6, 1.0, 1.5, 3.0]))
        cases = [("fb_T3000_K3", (3000, 3, 100, 1, 0.9, 1, "F")),
                 ("fb_T50000_K5", (50000, 5, 500, 2, 1.2, 1, "F")),
                 ("fb_T12000_K5_noself_lowthr", (12000, 5, 500, 3, 0.4, 0, "F")),
                 ("fb_T20000_K2", (20000, 2, 50, 4, 0.8, 1, "F")),
                 ("fb_T30000_K8", (30000, 8, 300, 6, 0.7, 1, "F")),
                 ("fb_T30000_K20", (30000, 20, 100, 8, 0.8, 1, "F")),
                 ("mix_T50000_K5", (50000, 5, 500, 2, 1.2, 1, "M")),
                 ("fb_T20000_K3_dyn5", (20000, 3, 200, 9, 0.0, 1, "F", 5, 1)),
                 ("mix_T20000_K3_dyn5", (20000, 3, 200, 9, 0.0, 1, "M", 5, 1))]
        for name, args in cases:
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **sweep_case(*args))
    md_cases = [("md_fb_T20000_P2_D2", (20000, 2, 2, 300, 11, 1.0, 1, "F")),
                ("md_fb_T15000_P3_D2", (15000, 3, 2, 200, 12, 0.9, 1, "F")),
                ("md_fb_T70000_P2_D3_noself", (70000, 2, 3, 400, 13, 1.1, 0, "F")),
                ("md_mix_T20000_P2_D2", (20000, 2, 2, 300, 11, 1.0, 1, "M")),
                ("md_fb_T10000_P2_D2_dyn4", (10000, 2, 2, 200, 14, 0.0, 1, "F", 4, 1))]
    for name, args in md_cases:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **md_case(*args))
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures written:", len(os.listdir(OUT)), "files,", total // 1024, "KiB")


if __name__ == "__main__":
    main()
