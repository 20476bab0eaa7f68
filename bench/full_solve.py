#!/usr/bin/env python
"""Full backward BSDE solve, BASELINE.json configs[2] ("C3"): HJB, d = 100, N = 2^20 paths, 50 time steps, TT rank 4,
4 monomials per asset, explicit scheme, ALS with 10 sweeps per fit — 51 fits, everything resident on the GPU
(reference: src/scripts/pipeline.py:66-141 with the literals of the config).

    python bench/full_solve.py [--samples 1048576] [--assets 100] [--steps 50] [--rank 4] [--degree 4] [--sweeps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 ... bench/full_solve.py

Under torchrun the N samples are split into G contiguous shards (strong scaling; Philox is indexed by the global
sample id, so the sample set is the same for every G) and the per-core moments are all-reduced every micro-step.
Prints one JSON line: wall time of path simulation and of the backward solve (device time between CUDA events,
max over ranks), samples*sweeps/s over the solve, Y0 and its distance from the Monte-Carlo reference price of the
HJB problem, -log E[exp(-g(X0 + sqrt(2 T) xi))] (bsde.py:98-102), evaluated here with 2^22 draws on the device.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), os.path.join(ROOT, "bsde-tensor-networks_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402


def hjb_reference_price(be, d, sigma, T, n_draws=1 << 22, seed=99):
    """-log mean(1 / (1/2 + 1/2 |X0 + sigma sqrt(T) xi|^2)) with X0 = 0: one Euler step of the library's own path kernel
    over [0, T] gives exactly X0 + sigma sqrt(T) xi; g through k_terminal; the mean of exp(-g) with torch (diagnostic)."""
    import torch

    acc, cnt = 0.0, 0
    chunk = 1 << 19
    for c in range(n_draws // chunk):
        x, _ = be.paths_simulate(1, np.array([sigma]), np.zeros(d), chunk, 1, T, seed=seed, sample_offset=c * chunk, keep_xi=False)
        g = be.terminal(1, np.array([sigma]), x[1].contiguous())
        acc += float(torch.exp(-g).sum())
        cnt += chunk
    return -np.log(acc / cnt)


def run(samples=1 << 20, assets=100, steps=50, rank=4, degree=4, sweeps=10, seed=0, verbose=False, model_name="hjb"):
    import torch
    import torch.distributed as dist

    from bsde_solver._backend import get_backend
    from bsde_solver._shard import shard_range
    from bsde_solver.bsde import HJB, BlackScholes
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.optimisation.als import ALS
    from bsde_solver.pipeline import predicted_price, run_explicit, simulate_paths

    world = dist.get_world_size() if dist.is_initialized() else 1
    rk = dist.get_rank() if dist.is_initialized() else 0
    be = get_backend()
    np.random.seed(seed)                          # initial cores come from the global numpy stream (als.py:37): same on every rank
    T = 1.0
    if model_name == "hjb":                       # BASELINE.json configs[2]: X0 = 0, sigma = sqrt(2)   (pipeline_explicit.py:33)
        sigma = float(np.sqrt(2.0))
        X0 = np.zeros(assets)
        model = HJB(np.tile(X0, (1, 1)), T / steps, T, sigma=sigma)
    else:                                         # the reference's Black-Scholes literals: X0 = (1, 0.5, ...), sigma = 0.4, r = 0.05
        sigma, rate = 0.4, 0.05                   # (src/scripts/pipeline.py:52-59)
        X0 = np.tile([1.0, 0.5], (assets + 1) // 2)[:assets]
        model = BlackScholes(np.tile(X0, (1, 1)), T / steps, T, rate, sigma)
    basis = PolynomialBasis(degree)
    ranks = (1,) + (rank,) * (assets - 1) + (1,)
    lo, hi = shard_range(samples, rk, world)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def elapsed_max(e0, e1):
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    sync()
    w0 = time.perf_counter()
    ev[0].record()
    X, _ = simulate_paths(model, X0, steps, hi - lo, paths="philox", seed=seed + 1000, sample_offset=lo, keep_noise=False)
    ev[1].record()
    Y, Vs, step_times = run_explicit(model, basis, X, sweeps, ranks, optimizer=ALS, verbose=verbose)
    ev[2].record()
    sync()
    wall = time.perf_counter() - w0
    t_paths, t_solve = elapsed_max(ev[0], ev[1]), elapsed_max(ev[1], ev[2])
    y0 = predicted_price(Y)
    out = None
    if rk == 0:
        if model_name == "hjb":
            ref, ref_kind = hjb_reference_price(be, assets, sigma, T), "Monte-Carlo closed form, 2^22 draws (bsde.py:98-102)"
        else:
            ref, ref_kind = float(np.exp((rate + sigma**2) * T) * np.sum(X0**2)), "analytic exp((r + sigma^2) T) |X0|^2 (bsde.py:54-55)"
        fits = steps + 1
        out = {"workload": f"full explicit BSDE solve: {'HJB' if model_name == 'hjb' else 'Black-Scholes'} d={assets}, N={samples} paths, {steps} steps, rank {rank}, {degree} monomials/asset, ALS {sweeps} sweeps/fit, {fits} fits",
               "n_gpus": world, "scaling": "strong (N split over the GPUs)", "paths_s": t_paths, "solve_s": t_solve, "total_s": t_paths + t_solve,
               "wall_s_host_clock": wall, "samples_sweeps_per_s": samples * sweeps * fits / t_solve,
               "Y0": y0, "Y0_reference": ref, "Y0_reference_kind": ref_kind, "Y0_rel_err": abs(y0 - ref) / abs(ref),
               "step_time_s_min_median_max": [float(np.min(step_times)), float(np.median(step_times)), float(np.max(step_times))],
               "paths_bytes_per_gpu": int(X.numel() * 8), "dtype": "f64"}
    del X, Y
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 20)
    ap.add_argument("--assets", type=int, default=100)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--rank", type=int, default=4)
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--model", default="hjb", choices=["hjb", "bs"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from bsde_solver._backend import get_backend

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        get_backend().init_distributed()
    out = run(args.samples, args.assets, args.steps, args.rank, args.degree, args.sweeps, args.seed, args.verbose, args.model)
    if out is not None:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
