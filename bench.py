#!/usr/bin/env python
"""bench.py — Gibbs sweeps/sec of the FBG hot path on the BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun launches N ranks for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's own CPU path, rank 0)

Workload (config.workload): BASELINE.json configs[3] — a single 1e9-observation synthetic piecewise-constant
Gaussian sequence, K = 5, dynamic wavelet blocks, FBG, self transitions on, no recording.  It fits one B200, so it
is the N = 1 workload too.  A step is one full Gibbs sweep (HMM.hpp:99-121): threshold from the current theta,
boundaries from the candidate list (the positions whose weight can reach the threshold: 8 bytes per candidate instead
of 4 bytes per observation), block statistics from the integral arrays, emission terms, forward
filter, backward sampling and reductions on the device; conjugate updates and parameter draws in the C++ host chain.
At N > 1 the SAME sequence is split into N contiguous segments, one per GPU, and the scan carries travel over NVLink
peer memory: the total work is fixed, i.e. strong scaling (`--mode independent` runs one sequence per GPU instead:
weak scaling, no collective).

One JSON line on stdout (rank 0).  `value` is timed with CUDA events on the stream the kernels run on; `e2e` is wall
clock through the same public call with host buffers in and out; `roofline` describes the per-sweep kernel with the
largest live CUDA-event time (algorithmic bytes of DESIGN.md §4 over its mean launch duration; `traffic` from the
ncu capture under profiles/, null when that capture predates the kernel sources of this tree); `invariants_ok` says
that the last timed sweep's counts sum to T; `stream_detect` / `pyramid_detect` time the weight-reading formulations
of boundary detection on the same data; `cpu_baseline` is the reference's own sampleHMM on the box's host cores on a
bounded sample.  The reference arm runs that same code on the first --ref-sample observations and scales by the
sample share (`extrapolated`, `sample_T` say so at the top level; profiles/r2_reference_linearity*.json pins the
extrapolation against a measured full-size run).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gibbs sweeps/sec (1e9 obs, K=5)"
UNIT = "sweeps/s"
SIGMA, SPACING = 0.3, 1.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def gen_chunk(torch, n, K, L, gen, carry_state, device):
    """n observations of the piecewise-constant recipe (levels spaced 1.0, sigma 0.3, segment lengths
    Geometric(1/L), segment level uniform on K), continuing from the previous chunk's level."""
    change = torch.rand(n, generator=gen, device=device) < (1.0 / L)
    seg = torch.cumsum(change.to(torch.int32), 0, dtype=torch.int32)
    nseg = int(seg[-1].item()) + 1
    levels = torch.randint(0, K, (nseg,), generator=gen, device=device, dtype=torch.int32)
    levels[0] = carry_state
    st = levels[seg.long()]
    x = (st.to(torch.float32) - (K - 1) / 2.0) * SPACING + SIGMA * torch.randn(n, generator=gen, device=device)
    return x, int(st[-1].item())


def generate(torch, T, K, L, seed, device, limit=None, keep=None):
    """The first min(T, limit) observations of the sequence, or — keep=(start, n) — only that slice of it (every
    rank of a segment-split run walks the same generator stream and keeps its own segment)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_total = T if limit is None else min(T, limit)
    k0, kn = keep if keep is not None else (0, n_total)
    n_total = min(n_total, k0 + kn)
    x = torch.empty(kn, dtype=torch.float32, device=device)
    chunk, done, carry = 1 << 26, 0, seed % K
    while done < n_total:
        n = min(chunk, T - done)         # chunk sizes depend on T only, so every slice sees the same stream
        xc, carry = gen_chunk(torch, n, K, L, gen, carry, device)
        lo, hi = max(done, k0), min(done + n, k0 + kn)
        if lo < hi:
            x[lo - k0:hi - k0] = xc[lo - done:hi - done]
        done += n
    return x


class ClockSampler:
    """SM clock and clock-event reasons of one GPU, polled through NVML from a thread for as long as the timed
    regions run (the sweeps are ctypes calls, so the GIL is free).  `nvidia-smi -lms` needs ~100 ms per row and
    misses a 20 ms timed region; NVML answers in well under a millisecond."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                         pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
        except Exception as e:  # noqa: BLE001 - reported in the JSON line
            self.nv, self.err = None, f"NVML unavailable: {e}"

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                power = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                self.rows.append((sm, reasons, power))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(0.002)

    def stop(self):
        if self.nv is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "sampler not started"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4) if r[1] & self.bits[i]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "power_w_max": max((r[2] for r in self.rows), default=None),
                "how": "NVML polled every ~2 ms from a thread across the timed regions (device-clock, end-to-end and "
                       "per-stage passes: the same sweeps)"}


def measured_peak():
    """HBM copy bandwidth in GB/s from the driver-written MEASURED_PEAKS.json (the kernels here are timed inside a long
    step, so a `sustained` figure is preferred over a `burst` one when the file distinguishes them), else the
    fallback of B200_PROFILING.md.  Any unexpected layout of the file falls back too, and says so."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        return 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(p) as f:
            doc = json.load(f)

        def number(v):
            if isinstance(v, (int, float)) and not isinstance(v, bool):
                return float(v)
            if isinstance(v, dict):
                for key in ("sustained", "sustained_gbs", "value", "gbs", "burst", "burst_gbs"):
                    if key in v:
                        r = number(v[key])
                        if r:
                            return r
            return None

        def find(d, path=""):
            if not isinstance(d, dict):
                return None
            for key in ("hbm_gbs_sustained", "hbm_sustained_gbs", "hbm_gbs", "hbm_gbps", "hbm"):
                if key in d:
                    r = number(d[key])
                    if r:
                        return r, path + key
            for k, v in d.items():
                if "hbm" in str(k).lower():
                    r = number(v)
                    if r:
                        return r, path + str(k)
            for k, v in d.items():
                r = find(v, path + str(k) + ".")
                if r:
                    return r
            return None

        hit = find(doc)
        if hit and 1000.0 < hit[0] < 20000.0:
            return hit[0], f"measured (MEASURED_PEAKS.json {hit[1]})"
        return 6650.0, "fallback (B200_PROFILING.md; MEASURED_PEAKS.json holds no HBM GB/s figure this script recognises)"
    except (OSError, ValueError, TypeError) as e:
        return 6650.0, f"fallback (B200_PROFILING.md; MEASURED_PEAKS.json unreadable: {e})"


def reference_sweeps(x_sample, K, burn, timed, reps, method="F"):
    """The reference's own sampleHMM (compiled from its sources into oracle/_ref/ref_probe) on a host sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refprobe
    r = refprobe.run("bench", x_sample, K=K, seed=1, burn=burn, timed=timed, reps=reps, method=method, raw32=True)
    secs = r["bench_secs"]
    return timed / float(np.min(secs)), int(r["bench_blocks"][0]), [float(s) for s in secs]


def kernel_sources_hash():
    """sha256 over the CUDA sources of the tree (hammlet_b200/csrc, sorted by name): what an ncu capture belongs to."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "hammlet_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=float, default=1e9)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--L", type=int, default=5000)
    ap.add_argument("--sample", type=float, default=3e7, help="observations of the cpu_baseline sample (our arm)")
    ap.add_argument("--ref-sample", type=float, default=1e8,
                    help="observations the reference arm sweeps (>= T: the whole workload, ~15 min and ~25 GB at 1e9)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="segments", choices=["segments", "independent"],
                    help="N > 1: one sequence split into contiguous segments with NCCL carry exchange (strong scaling, "
                         "BASELINE configs[3]) or one independent sequence per GPU (weak scaling, no collective)")
    args = ap.parse_args()
    T, K, L = int(args.T), args.K, args.L
    steps, warmup = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"single {T:.0e}-observation piecewise-constant Gaussian sequence, K={K}, mean segment {L}, " \
               f"sigma {SIGMA}, level spacing {SPACING}, dynamic blocks, FBG, no recording"

    import torch

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        Ts = int(min(args.ref_sample, T))
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        x = generate(torch, T, K, L, seed=4, device=dev, limit=Ts).cpu().numpy()
        burn = 100
        sps, nb, secs = reference_sweeps(x, K, burn=burn, timed=steps, reps=max(1, warmup // 3))
        scaled = sps * Ts / T
        pinned = None
        try:  # the one-off full-size run that pins the extrapolation (tools/reference_linearity.py)
            with open(os.path.join(ROOT, "profiles", "r2_reference_linearity.json")) as f:
                doc = json.load(f)
            pinned = {"measured_full_config_sweeps_per_s": doc.get("measured_full_config_sweeps_per_s"),
                      "extrapolated_over_measured": doc.get("extrapolated_over_measured"), "host": doc.get("host"),
                      "file": "profiles/r2_reference_linearity.json"}
        except (OSError, ValueError):
            pass
        line = {
            "impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1000.0 / scaled, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "extrapolated": Ts < T, "sample_T": Ts, "same_config": Ts == T, "measured_on_sample_sweeps_per_s": sps,
            "linearity": pinned,
            "config": {"workload": workload, "states": K, "observations": T},
            "cpu_baseline": {"value": scaled, "unit": UNIT, "cores": 1, "kind": "reference",
                             "sample": f"first {Ts} observations of the workload through the reference's own sampleHMM "
                                       f"(oracle/_ref/ref_probe, g++ -O3, 1 thread — the reference is single-threaded; "
                                       f"host: {cpu_model()}, {os.cpu_count()} cores); {burn} burn-in sweeps, best of "
                                       f"{len(secs)} x {steps} timed sweeps = {sps:.2f} sweeps/s at {nb} blocks"
                                       + (f"; scaled by {Ts}/{T} because the reference's sweep cost is linear in the "
                                          f"number of blocks (SURVEY.md §6.2; measured against a full-size run in "
                                          f"profiles/r2_reference_linearity.json)" if Ts < T else "; the whole workload")},
            "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), file=json_out, flush=True)
        return

    # ------------------------------------------------------------------ our arm
    from hammlet_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hammlet_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)

    segments = world > 1 and args.mode == "segments"
    t0 = time.time()
    h = capi.Handle(local_rank)
    if segments:
        # one sequence, contiguous segments: every rank keeps its slice of the same generator stream
        uid = [capi.Handle.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        h.comm_init(rank, world, uid[0])
        seg_start, seg_len = capi.Handle.segment_plan(T, world, rank)
        x = generate(torch, T, K, L, seed=4, device=device, keep=(seg_start, seg_len))
        torch.cuda.synchronize()
        sample_host = x[:int(min(args.sample, seg_len))].cpu().numpy() if (rank == 0 and not args.no_cpu_baseline) else None
        h.load_segment_device(x.data_ptr(), seg_len, T)
        T_local = seg_len
    else:
        x = generate(torch, T, K, L, seed=4 + rank, device=device)
        torch.cuda.synchronize()
        sample_host = x[:int(min(args.sample, T))].cpu().numpy() if (rank == 0 and not args.no_cpu_baseline) else None
        h.load_device(x.data_ptr(), T)
        T_local = T
    del x
    torch.cuda.empty_cache()
    t_load = time.time() - t0
    log(f"[rank {rank}] generated + loaded {T_local} of T={T} in {t_load:.1f}s, sigma_hat={h.sigma_hat():.4f}")

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # The chain is the C++ host side (include/hammlet_host.h): sampleHMM, conjugate updates and parameter draws
    # in C++ with libstdc++ <random> like the reference; Python only starts and stops the clock.
    # (auto priors of a split sequence: the C++ side gathers the ranks' block lists, Emissions.hpp Blocks::fetch)
    tau = capi.Chain.auto_prior(h, 0.2, 0.9)
    # segment mode: every rank draws the same parameters from the same (all-gathered) statistics
    chain = capi.Chain(h, K, tau, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=100 if segments else 100 + rank)
    # start near the generating model so that the warm-up sweeps reach the stationary compression ratio quickly
    chain.set(((np.arange(K) - (K - 1) / 2.0) * SPACING).astype(np.float32), np.full(K, SIGMA * SIGMA, np.float32),
              (np.full((K, K), 0.0002 / (K - 1)) + np.eye(K) * (0.9998 - 0.0002 / (K - 1))).astype(np.float32),
              np.full(K, 1.0 / K, np.float32))
    stream = torch.cuda.ExternalStream(h.stream(), device=device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    chain.run(warmup)

    # ---- timed region 1: device clock (CUDA events on the library's stream), K steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    chain.run(steps)
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = h.launch_count() - launches0
    # invariants of the last timed sweep (ForwardBackward.hpp:177-200): every observation is counted once in the
    # occupancy, once as a transition target (the phantom 0 -> q0 included) and once in a parameter's term count
    last = chain.last_sweep()
    _, cands_now = h.detect_info()
    inv = {"T": T, "sum_trans": int(last["trans"].sum()), "sum_counts": int(last["counts"].sum()),
           "sum_stat_n": int(last["stat_n"].sum()), "nblocks": int(last["nblocks"]), "candidates_this_rank": int(cands_now)}

    # ---- timed region 2: end to end (wall clock around the public call; host buffers in and out every sweep)
    barrier()
    w0 = time.perf_counter()
    chain.run(steps)
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0

    # ---- the second number SURVEY.md §8d asks for: thinning 1 — every sweep is recorded into the state marginals
    # (Records -> device-resident marginals).  A split sequence keeps the marginals of each rank's positions on that
    # rank (the ranks trade 8 bytes per recorded sweep so that runs continue across borders) and merges them when asked.
    rec_steps = max(10, steps // 5)
    chain.run_recorded(max(3, warmup // 4), thinning=1)   # warm-up: first use allocates the run / marginal buffers
    barrier()
    r0 = time.perf_counter()
    _, nseg = chain.run_recorded(rec_steps, thinning=1)
    torch.cuda.synchronize()
    rec_wall = time.perf_counter() - r0
    runs = int(h.segments()[0].size)
    if dist is not None:
        t = torch.tensor([rec_wall], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec_wall = float(t[0].item())
    recorded = {"value": rec_steps / rec_wall * (world if (world > 1 and not segments) else 1), "unit": UNIT, "thinning": 1,
                "steps": rec_steps, "runs_last_sweep": runs, "marginal_segments": int(nseg),
                "how": "wall clock through hammlet_chain_run_recorded: sweep + equal-state runs formed on the device + "
                       "merge into the device-resident state marginals (hml_marginals_add: StateMarginals::addRecord "
                       "as three small kernels; nothing but a 4-byte count returns to the host per sweep)"}

    # how the forward filter ran so far (warm-up, the timed regions, the recorded sweeps): speculative sweeps, how many had
    # to be repeated through the operator scan, and the (piece, warm-up) level it settled on
    def forward_doc():
        _, nspec, nrep = h.forward_info()
        piece, warm = h.forward_level()
        return {"speculative_sweeps": nspec, "repeated": nrep, "piece_blocks": piece, "warmup_blocks": warm}
    forward_timed = forward_doc()

    # ---- region 3 (not part of `value`): the same steps with per-stage CUDA events, for the roofline and stage table
    h.set_timing(True)
    stage_ms, nblocks = {}, []
    for i in range(steps):
        nblocks.append(chain.run(1))
        for name, ms in h.timing():
            stage_ms.setdefault(name, []).append(ms)
    h.set_timing(False)
    forward_staged = forward_doc()
    clocks = sampler.stop()

    # ---- the streaming formulation of boundary detection (4 B/observation), timed on the same data
    # (collective in segment mode, so every rank runs it)
    _, cands = h.detect_info()           # candidates the last candidate pass looked at (this rank)
    h.set_timing(True)
    _, var_now, _, _ = chain.get()
    thr_now = float(np.sqrt(np.float32(2) * np.log(np.float32(T)) * var_now.min(), dtype=np.float32))
    h.set_detect_mode(capi.DETECT_PYRAMID)
    tp = []
    for _ in range(8):
        h.create_blocks(thr_now)
        tm = dict(h.timing())
        tp.append(tm["detect_hot"] + tm["detect_scatter"])
    t_pyramid = float(np.mean(tp[3:]))
    _, hot = h.detect_info()             # sub-blocks of 32 weights the pyramid pass had to read
    h.set_detect_mode(capi.DETECT_STREAM)
    ts = []
    for _ in range(8):
        h.create_blocks(thr_now)
        ts.append(dict(h.timing())["detect_flags"])
    t_stream = float(np.mean(ts[3:]))
    h.set_detect_mode(capi.DETECT_CANDIDATES)
    h.set_timing(False)

    if dist is not None:
        t = torch.tensor([dev_ms, wall * 1000.0], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0].item()), float(t[1].item()) / 1000.0
        invs = allgather(inv)
    else:
        invs = [inv]
    if segments:   # one sequence: every rank reports the global statistics, and they must be the same ones
        same = all(v == invs[0] or {k: v[k] for k in v if k != "candidates_this_rank"} ==
                   {k: invs[0][k] for k in invs[0] if k != "candidates_this_rank"} for v in invs)
        cand_total = sum(v["candidates_this_rank"] for v in invs)
        invariants_ok = bool(same and inv["sum_trans"] == T and inv["sum_counts"] == T and inv["sum_stat_n"] == T
                             and 0 < inv["nblocks"] <= cand_total)
        inv_doc = dict(invs[0], ranks_agree=same, candidates_total=cand_total)
    else:          # one sequence per rank: each checks its own
        invariants_ok = all(v["sum_trans"] == T and v["sum_counts"] == T and v["sum_stat_n"] == T
                            and 0 < v["nblocks"] <= v["candidates_this_rank"] for v in invs)
        inv_doc = dict(invs[0], ranks_checked=len(invs), candidates_total=invs[0]["candidates_this_rank"])
    inv_doc.pop("candidates_this_rank", None)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms / steps
    jobs = 1 if segments else world      # sequences swept per step by the whole job
    value = jobs * steps / (dev_ms / 1000.0)
    e2e = jobs * steps / wall
    B = float(np.mean(nblocks))          # blocks of one whole sequence
    Bl = B * T_local / T                 # ... of which on this rank (about)
    peak, peak_src = measured_peak()
    busy = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    # algorithmic bytes per launch of the kernels that can dominate a sweep
    alg = {
        "detect_cand": 4.0 * cands + 4.0 * Bl + 4.0 * Bl,                              # candidate weights, positions of the hits, starts
        "detect_hot": T_local / 16.0 + 128.0 * hot + 8.0 * hot,                        # bf16 pyramid + hot sub-blocks + triples
        "detect_flags": 4.0 * T_local + 4.0 * Bl,                                      # every weight + starts
        "block_emit": Bl * (4 + 16 + 4 + 16 + 16 * K),                                 # starts, integral gathers, N, sums, e, sp
        "fwd_chunks": Bl * 8 * K * (1 + K / 32.0),
        "fwd_replay": Bl * 16 * K,
        "fwd_spec": Bl * 8 * K * (2 + 4 / 32.0),                                       # e (+ 4 warm-up blocks per chunk of 32), rows
        "bwd_maps": Bl * (8 * K + 4 + 8),                                              # rows, N, the K -> K map of the block
    }
    kernels = {k: v for k, v in busy.items() if k in alg}
    top = max(kernels, key=kernels.get)
    det = busy[top]
    alg_bytes = alg[top]
    achieved = alg_bytes / (det * 1e-3) / 1e9
    sb = 4.0 * T_local + 4.0 * Bl
    stream = {"kernel": "k_detect_flags", "achieved": sb / (t_stream * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
              "frac": sb / (t_stream * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": sb, "avg_launch_ms": t_stream,
              "note": "hml_set_detect_mode(HML_DETECT_STREAM): reads every weight; not used by the timed sweeps"}
    pb = T_local / 16.0 + 136.0 * hot
    pyramid = {"kernel": "k_detect_hot + k_scatter_hot", "achieved": pb / (t_pyramid * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
               "frac": pb / (t_pyramid * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": pb, "avg_launch_ms": t_pyramid,
               "hot_subblocks": hot,
               "note": "hml_set_detect_mode(HML_DETECT_PYRAMID): bf16 max pyramid + the sub-blocks that can hold a boundary; "
                       "builds the candidate list, not used by the timed sweeps otherwise"}
    h2d = 8 * (2 * K + K * K + K)        # mean, var, A, pi as doubles (kernel parameters built from host buffers)
    d2h = 8 * (2 + K + K * K + 2 + 2 * K + 1) * (world if segments else 1)
    if world == 1:
        wl = workload
    elif segments:
        wl = f"{workload}; split into {world} contiguous segments, one per GPU, scan carries exchanged over NVLink peer memory (NCCL fallback)"
    else:
        wl = f"{world} independent sequences, one per GPU, each: " + workload

    # DRAM bytes of one launch of the roofline kernel from an `ncu --set full` capture of this workload (profiles/)
    # The capture is stamped with a hash of the kernel sources it was taken from (tools/ncu_traffic.py): a capture that
    # predates the current kernels is reported as stale and `traffic` stays null.
    traffic, traffic_note = None, "no capture under profiles/ for this kernel and workload"
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            doc = json.load(f)
        rec = doc.get("k_" + top)
        if rec and int(rec["T"]) == T_local and int(rec["K"]) == K:
            if doc.get("_kernel_sources_sha256") == kernel_sources_hash():
                traffic = float(rec["dram_bytes_per_launch"])
                traffic_note = f"ncu --set full capture of git {doc.get('_git_head', '?')} (same kernel sources as this tree)"
            else:
                traffic_note = f"stale: the capture of git {doc.get('_git_head', '?')} predates the kernel sources of this tree"
    except (OSError, ValueError, KeyError):
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        # N ranks split ONE sequence (total work fixed): strong scaling, also the label of the N = 1 point of that curve
        "scaling": "weak" if (world > 1 and not segments) or args.mode == "independent" else "strong",
        "invariants_ok": invariants_ok, "invariants": inv_doc,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "mode": "segments" if segments else ("independent" if world > 1 else "single"),
                   "states": K, "observations": T, "observations_per_gpu": T_local, "blocks_per_sweep": B,
                   "compression_ratio": T / B,
                   "l2_policy": f"per GPU and sweep the kernels touch {8.0 * cands / 1e6:.0f} MB of candidates, "
                                f"{Bl * 32 / 1e6:.0f} MB of scattered integral-array entries (out of {16.0 * T_local / 1e9:.1f} GB) and "
                                f"{Bl * (24 + 24 * K + 10) / 1e6:.0f} MB of per-block arrays, each written by one kernel and read by a "
                                f"later one: larger than the 126 MB L2 together, no flush between sweeps",
                   "detect_mode": "candidates", "candidates_per_sweep": cands,
                   "carry_exchange": h.exchange_transport(),
                   "load_seconds": t_load},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_" + top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": det,
                     "sweep_bytes": 4.0 * T + B * (4 + 16 + 16 * K + 2),
                     "sweep_frac_of_hbm_roofline": (4.0 * T + B * (4 + 16 + 16 * K + 2)) / (world if segments else 1)
                                                   / peak / 1e9 / (ms_per_step * 1e-3)},
        "recorded": recorded,
        "forward_filter": {"mode": "auto (speculative pieces + repair pass; operator scan after a failure)",
                           "after_timed_regions": forward_timed, "after_stage_pass": forward_staged},
        "stream_detect": stream,
        "pyramid_detect": pyramid,
        "stage_ms": busy,
        "device_busy_ms_per_step": float(sum(busy.values())),
    }
    if sample_host is not None:
        try:
            Ts = sample_host.size
            sps, nb, secs = reference_sweeps(sample_host, K, burn=100, timed=200, reps=2)
            line["cpu_baseline"] = {
                "value": sps * Ts / T, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"first {Ts} observations through the reference's own sampleHMM (oracle/_ref/ref_probe, 1 thread; "
                          f"host {cpu_model()}, {os.cpu_count()} cores): {sps:.2f} sweeps/s at {nb} blocks after 100 burn-in "
                          f"sweeps (best of 2 x 200 timed sweeps), scaled by {Ts}/{T} (sweep cost linear in #blocks)"}
        except Exception as e:  # the checker is optional for the number, never for the product
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"unavailable: {e}"}
    print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
