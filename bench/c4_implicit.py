#!/usr/bin/env python
"""BASELINE.json configs[3] ("C4"): implicit scheme with SALSA rank adaptation, HJB d = 100, N = 2^20 paths, fixed-point
iterations per time step (reference: src/scripts/pipeline_explicit.py:74-112 in the intended form of
src/scripts/experiments/n_assets.py:155-191; SALSA calls with max_rank = degree, n_iter = 20, do_reg = False at n = 0).

    python bench/c4_implicit.py [--samples 1048576] [--assets 100] [--steps 2] [--K 3] [--sweeps 20] [--rank 2] [--degree 4]

Times the terminal SALSA fit and `steps` backward time steps of K fixed-point iterations each, everything resident on the
GPU (run_implicit), and prints one JSON line: seconds per fixed-point iteration (gradient + target + SALSA fit + evaluation),
samples*sweeps/s of the SALSA fits (one SALSA sweep = d micro-steps, als.py:235), bond ranks reached, fit residual of the
terminal fit.  The train starts at `rank` and may grow to max_rank = degree (rank adaptation with its host read-backs in
the early sweeps).  The full config (100 time steps x 30 iterations) is this number times 3000."""
import argparse
import json
import os
import sys
import time
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), os.path.join(ROOT, "bsde-tensor-networks_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 20)
    ap.add_argument("--assets", type=int, default=100)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--sweeps", type=int, default=20)
    ap.add_argument("--rank", type=int, default=2)
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--total-steps", type=int, default=100, help="time steps of the full config (dt = T / total_steps)")
    args = ap.parse_args()
    import torch

    from bsde_solver._backend import get_backend
    from bsde_solver.bsde import HJB
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.optimisation import als as A
    from bsde_solver.pipeline import run_implicit, simulate_paths

    be = get_backend()
    d, N = args.assets, args.samples
    np.random.seed(0)
    sigma = float(np.sqrt(2.0))
    # the last `steps` time steps of the full grid: paths simulated over [0, T] with total_steps steps, the tail kept
    T = 1.0
    model = HJB(np.zeros((1, d)), T / args.total_steps, T, sigma=sigma)
    X, xi = simulate_paths(model, np.zeros(d), args.total_steps, N, paths="philox", seed=7)
    X = X[args.total_steps - args.steps:].contiguous()
    xi = xi[args.total_steps - args.steps:].contiguous()
    sub = HJB(np.zeros((1, d)), T / args.total_steps, args.steps * T / args.total_steps, sigma=sigma)     # dt as in the full grid
    ranks = (1,) + (args.rank,) * (d - 1) + (1,)
    opt = partial(A.SALSA, max_rank=args.degree)
    fits = []
    orig = A.get_backend

    torch.cuda.synchronize()
    l0 = be.launch_count()
    t0 = time.perf_counter()
    Y, V, step_times = run_implicit(sub, PolynomialBasis(args.degree), X, xi, args.sweeps, args.K, ranks, optimizer=opt)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    launches = be.launch_count() - l0
    n_fits = 1 + args.steps * args.K
    inp = PolynomialBasis(args.degree).device_inputs(X[0])
    out = {"workload": f"C4: implicit scheme, HJB d={d}, N=2^{int(np.log2(N))} paths, SALSA n_iter={args.sweeps}, max_rank={args.degree} from rank {args.rank}, "
                       f"{args.steps} time steps x K={args.K} fixed-point iterations (+ the terminal fit)",
           "seconds_total": total, "fits": n_fits, "seconds_per_fixed_point_iteration": float(np.mean(step_times)) / args.K,
           "samples_sweeps_per_s": N * args.sweeps * n_fits / total, "sweep_definition": "one SALSA sweep = d micro-steps (als.py:235)",
           "us_per_micro_step": 1e6 * total / (n_fits * args.sweeps * d),
           "final_ranks_max": int(max(c.shape[0] for c in V.cores.values())), "rank_limited": bool(A.last_fit_info.get("rank_limited", False)),
           "solves_last_fit": {k: A.last_fit_info.get(k) for k in ("cholesky", "jacobi", "lu")},
           "gpu_launches": int(launches), "Y_mean_per_step": [float(v) for v in Y.mean(dim=1).cpu().numpy()],
           "extrapolated_full_config_hours": (float(np.mean(step_times)) / args.K) * 30 * args.total_steps / 3600.0,
           "dtype": "f64"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
