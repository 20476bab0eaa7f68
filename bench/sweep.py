#!/usr/bin/env python
"""Parameter sweeps of one ALS fit, JSON lines instead of the reference's matplotlib plots
(reference: src/scripts/experiments/benchmark_cupy.py:28-36 — base point and the four axes N, p, d, r — and BASELINE.json
configs[4], the n_assets sweep d = 10..500 of src/scripts/experiments/n_assets.py:200-235).

    python bench/sweep.py --vary d --values 10,20,50,100,200,500 [--samples 1048576] [--rank 4] [--basis 4] [--sweeps 10]
    python bench/sweep.py --vary N --values 1000,2000,5000,10000,20000 --assets 10 --rank 5 --basis 5 --sweeps 20   # benchmark_cupy axes
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 ... bench/sweep.py ...    # samples are per GPU

Synthetic HJB-law samples x ~ N(0, 2 I) from the library's Philox path kernel, y = log(1/2 + 1/2 |x|^2).  One JSON line per
point: device time of the fit (CUDA events, max over ranks), samples*sweeps/s over all ranks, algorithmic flops per
sample*sweep and the fraction of the FP64 peak they amount to, and which solver path the micro-steps took.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), os.path.join(ROOT, "bsde-tensor-networks_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

FP64_PEAK_TFLOPS = 37.0      # bench/micro/fp64_peak.cu, profiles/fp64_peak_r01.txt


def flops_per_sample_sweep(d, r, p):
    """n(n+1)+2n per micro-step (n = unknowns of the core) plus 2 r_l p r_r per interface update."""
    ranks = [1] + [r] * (d - 1) + [1]
    total = 0
    order = list(range(d - 1)) + list(range(d - 1, 0, -1))
    for j in order:
        n = ranks[j] * p * ranks[j + 1]
        total += n * (n + 1) + 2 * n + 2 * n
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vary", choices=["d", "N", "p", "r"], default="d")
    ap.add_argument("--values", default="10,20,50,100,200,500")
    ap.add_argument("--assets", type=int, default=100)
    ap.add_argument("--samples", type=int, default=1 << 20)
    ap.add_argument("--rank", type=int, default=4)
    ap.add_argument("--basis", type=int, default=4)
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    be = get_backend()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        be.init_distributed()

    for v in [int(x) for x in args.values.split(",")]:
        d, N, p, r = args.assets, args.samples, args.basis, args.rank
        if args.vary == "d":
            d = v
        elif args.vary == "N":
            N = v
        elif args.vary == "p":
            p = v
        else:
            r = v
        sig = np.array([np.sqrt(2.0)])
        x, _ = be.paths_simulate(1, sig, np.zeros(d), N, 1, 1.0, seed=2024, sample_offset=rank * N, keep_xi=False)
        xT = x[1].contiguous()
        y = be.terminal(1, sig, xT)
        del x
        inp = DeviceInputs(BASIS_POLY, p, xT)
        ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
        rng = np.random.RandomState(7)
        init = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
        try:
            info = be.fit_als(inp, y, args.sweeps, ranks, 0, init.copy())          # warm-up: plans, graph
        except Exception as exc:                                                     # a limit of the build: say so, go on
            if rank == 0:
                print(json.dumps({"vary": args.vary, "d": d, "samples_per_gpu": N, "basis": p, "rank": r, "unsupported": str(exc)}), flush=True)
            del xT, y, inp
            continue
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.repeats):
            info = be.fit_als(inp, y, args.sweeps, ranks, 0, init.copy())
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.repeats], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        value = world * N * args.sweeps / (ms * 1e-3)
        fl = flops_per_sample_sweep(d, r, p)
        if rank == 0:
            print(json.dumps({"vary": args.vary, "d": d, "samples_per_gpu": N, "basis": p, "rank": r, "sweeps": args.sweeps, "n_gpus": world,
                              "ms_per_fit": ms, "samples_sweeps_per_s": value, "algorithmic_flops_per_sample_sweep": fl,
                              "algorithmic_tflops_per_gpu": value / world * fl / 1e12,
                              "frac_of_fp64_peak": value / world * fl / 1e12 / FP64_PEAK_TFLOPS,
                              "solves": {"cholesky": int(info[0]), "eigen": int(info[1]), "lu": int(info[2])}}), flush=True)
        del xT, y, inp
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
