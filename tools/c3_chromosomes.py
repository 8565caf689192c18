#!/usr/bin/env python
"""BASELINE.json configs[2]: a WGS depth-of-coverage-shaped genome as 24 independent chromosome sequences (hg38
lengths, 3.09e9 bins in total, K = 5, mean segment 50 000), sharded over the GPUs of one box by longest-processing-time
bin packing (no collective; every sequence is its own chain with its own parameters and RNG stream).

  python tools/c3_chromosomes.py [--scale 1.0] [--sweeps 200] [--streams 4] [--out gpurun_out/c3.json]
  python -m torch.distributed.run --nproc-per-node N ... tools/c3_chromosomes.py    (one rank per GPU)

Two timings per GPU: the sequences swept one after the other, and `--streams` chains at a time (hammlet_chains_run:
C++ host threads, every handle owns a CUDA stream), which lets the latency-bound kernels of one chain (tile scans,
map scans: one CTA or one cluster) overlap the wide kernels of another.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hammlet_b200.synth import HG38, lpt_assign  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the chromosome lengths (quick runs)")
    ap.add_argument("--sweeps", type=int, default=200)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c3.json"))
    args = ap.parse_args()
    import torch
    from hammlet_b200 import capi
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    K, L = 5, 50000
    lengths = [max(4096, int(n * args.scale)) for n in HG38]
    mine = lpt_assign(lengths, world)[rank]
    bench.SPACING = 1.0
    t0 = time.time()
    chains, handles = [], []
    for i in mine:
        x = bench.generate(torch, lengths[i], K, L, seed=100 + i, device=device)
        h = capi.Handle(local)
        h.load_device(x.data_ptr(), lengths[i])
        del x
        tau = capi.Chain.auto_prior(h, 0.2, 0.9)
        c = capi.Chain(h, K, tau, trans=0.5, self_trans=0.5, alpha_pi=0.5, seed=100 + i)
        c.set(((np.arange(K) - (K - 1) / 2.0)).astype(np.float32), np.full(K, 0.09, np.float32),
              (np.full((K, K), 0.00002 / (K - 1)) + np.eye(K) * (0.99998 - 0.00002 / (K - 1))).astype(np.float32),
              np.full(K, 1.0 / K, np.float32))
        c.run(10)
        handles.append(h)
        chains.append(c)
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_load = time.time() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- one sequence after the other
    barrier()
    w0 = time.perf_counter()
    blocks = [c.run(args.sweeps) for c in chains]
    torch.cuda.synchronize()
    t_seq = time.perf_counter() - w0

    # ---- `streams` chains at a time: hammlet_chains_run (C++ host threads, longest sequence first)
    barrier()
    w0 = time.perf_counter()
    capi.run_chains(chains, args.sweeps, threads=max(1, args.streams))
    torch.cuda.synchronize()
    t_par = time.perf_counter() - w0

    res = torch.tensor([t_seq, t_par, t_load], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    t_seq, t_par, t_load = (float(v) for v in res)
    if rank == 0:
        total = float(sum(lengths))
        line = {
            "config": f"C3: 24 chromosome-length sequences ({total:.3e} bins, scale {args.scale}), K=5, mean segment {L}, "
                      f"{args.sweeps} dynamic FBG sweeps each, LPT-sharded over {world} GPU(s), no collective",
            "n_gpus": world, "sweeps": args.sweeps, "sequences": len(lengths), "sequences_rank0": len(mine),
            "seconds_sequential": t_seq, "seconds_concurrent": t_par, "streams": args.streams,
            "genome_sweeps_per_s_sequential": args.sweeps / t_seq, "genome_sweeps_per_s_concurrent": args.sweeps / t_par,
            "bin_sweeps_per_s_concurrent": total * args.sweeps / t_par,
            "load_seconds": t_load, "blocks_last_sweep_rank0": [int(b) for b in blocks],
            "timing": "wall clock around the chains, barrier + device synchronize on both sides, max over ranks",
        }
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(json.dumps(line) + "\n")
    for c in chains:
        c.close()
    for h in handles:
        h.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
