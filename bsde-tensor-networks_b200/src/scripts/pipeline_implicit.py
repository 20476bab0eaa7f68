"""Implicit backward scheme with SALSA — `python -m src.scripts.pipeline_implicit`
(reference: src/scripts/pipeline_explicit.py, which despite its name is the implicit scheme; same default literals:
HJB with sigma = sqrt(2), X0 = 0, 4 assets, batch 2000, 100 steps, 20 SALSA sweeps, 30 fixed-point iterations, rank 3,
3 monomials).  Everything runs on the GPU (bsde_solver.pipeline.run_implicit).

Environment overrides: BSDE_ASSETS, BSDE_BATCH, BSDE_STEPS, BSDE_ITER, BSDE_ITER_IMPLICIT, BSDE_RANK, BSDE_DEGREE,
BSDE_MODE (SALSA|ALS), BSDE_PATHS (numpy|philox), BSDE_SEED.
"""
import os
import time
from functools import partial

from bsde_solver import xp
from bsde_solver.bsde import HJB
from bsde_solver.core.calculus.basis import PolynomialBasis
from bsde_solver.core.optimisation.als import ALS, SALSA
from bsde_solver.pipeline import predicted_price, run_implicit, simulate_paths


def main():
    batch_size = int(os.environ.get("BSDE_BATCH", 2000))
    T = 1
    N = int(os.environ.get("BSDE_STEPS", 100))
    num_assets = int(os.environ.get("BSDE_ASSETS", 4))
    dt = T / N
    n_iter = int(os.environ.get("BSDE_ITER", 20))
    n_iter_implicit = int(os.environ.get("BSDE_ITER_IMPLICIT", 30))
    rank = int(os.environ.get("BSDE_RANK", 3))
    degree = int(os.environ.get("BSDE_DEGREE", 3))
    ranks = (1,) + (rank,) * (num_assets - 1) + (1,)
    basis = PolynomialBasis(degree)
    seed = os.environ.get("BSDE_SEED")
    if seed is not None:
        xp.random.seed(int(seed))
    optimizer = ALS if os.environ.get("BSDE_MODE", "SALSA") == "ALS" else partial(SALSA, max_rank=degree)

    X0 = xp.zeros(num_assets)
    model = HJB(X0=xp.tile(X0, (1, 1)), delta_t=dt, T=T, sigma=xp.sqrt(2))
    print(f"{num_assets} assets | {N} steps | {batch_size} batch size | {n_iter} iterations | {degree} degree | {rank} rank")
    X, noise = simulate_paths(model, X0, N, batch_size, paths=os.environ.get("BSDE_PATHS", "numpy"), seed=int(seed or 0))
    print("Start optimisation")
    start_time = time.perf_counter()
    Y, V, step_times = run_implicit(model, basis, X, noise, n_iter, n_iter_implicit, ranks, optimizer=optimizer, verbose=True)
    print("End")
    print(f"Time: {time.perf_counter() - start_time:.2f}s")
    print(f"Mean step time: {xp.mean(step_times):.2f}s")
    print()
    # ground truth at t = 0: every sample sits at X0, so a handful of rows gives the same mean as the whole batch
    # (the reference evaluates all of them: N * d * 1000 normals for HJB, bsde.py:98-102)
    price = xp.mean(model.price(X[0][:, :64].contiguous(), 0))
    predicted = predicted_price(Y)
    print("Price at 0", price)
    print("Predicted Price:", predicted)
    print("Relative error:", f"{xp.abs(price - predicted) / xp.abs(price) * 100:.2f}%")
    return predicted, price


if __name__ == "__main__":
    main()
