#!/usr/bin/env python
"""Explicit against implicit scheme over the number of assets — the experiment of
src/scripts/experiments/n_assets.py:200-235 (Black-Scholes, X0 = (0.5, 1, ...), sigma = 0.4, r = 0.05, SALSA with
max_rank = degree, Z = sigma grad V in both schemes, do_reg = False at n = 0), JSON lines instead of a matplotlib plot.

    python bench/explicit_vs_implicit.py [--assets 4,8,10,16,20] [--batch 2000] [--steps 100] [--sweeps 50] [--K 10] [--rank 3] [--degree 3]

Defaults are the reference's literals.  Everything is resident on the GPU (run_explicit / run_implicit); per point: Y0 of
both schemes, the analytic price exp((r + sigma^2) T) |X0|^2 (bsde.py:54-55), relative errors and wall times."""
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
    ap.add_argument("--assets", default="4,8,10,16,20")
    ap.add_argument("--batch", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--rank", type=int, default=3)
    ap.add_argument("--degree", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch

    from bsde_solver.bsde import BlackScholes
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.optimisation.als import SALSA
    from bsde_solver.pipeline import predicted_price, run_explicit, run_implicit, simulate_paths

    for d in [int(v) for v in args.assets.split(",")]:
        np.random.seed(args.seed)
        x0 = np.tile([0.5, 1.0], (d + 1) // 2)[:d]
        model = BlackScholes(x0[None, :], 1.0 / args.steps, 1.0, 0.05, 0.4)
        X, xi = simulate_paths(model, x0, args.steps, args.batch, paths="numpy")
        ranks = (1,) + (args.rank,) * (d - 1) + (1,)
        opt = partial(SALSA, max_rank=args.degree)
        basis = PolynomialBasis(args.degree)
        price = float(np.exp((0.05 + 0.4**2) * 1.0) * np.sum(x0**2))
        out = {"assets": d, "batch": args.batch, "steps": args.steps, "sweeps": args.sweeps, "K": args.K, "rank": args.rank,
               "degree": args.degree, "price_analytic": price}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Y, _, _ = run_explicit(model, basis, X, args.sweeps, ranks, optimizer=opt, apply_sigma=True, salsa_no_reg_at_0=True)
        y0 = predicted_price(Y)
        out.update(explicit_Y0=y0, explicit_rel_err=abs(y0 - price) / price, explicit_seconds=time.perf_counter() - t0)
        t0 = time.perf_counter()
        Y, _, _ = run_implicit(model, basis, X, xi, args.sweeps, args.K, ranks, optimizer=opt)
        y0 = predicted_price(Y)
        out.update(implicit_Y0=y0, implicit_rel_err=abs(y0 - price) / price, implicit_seconds=time.perf_counter() - t0)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
