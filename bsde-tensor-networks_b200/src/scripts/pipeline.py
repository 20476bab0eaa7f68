"""Explicit backward scheme — `python -m src.scripts.pipeline` / `python -m src.pipeline`
(reference: src/scripts/pipeline.py; same default literals: BlackScholes, 4 assets, batch 2000, 50 steps, 10 sweeps,
rank 2, 3 monomials, SALSA with max_rank 5).  Everything runs on the GPU (bsde_solver.pipeline.run_explicit).

Overrides for the larger configurations, as environment variables: BSDE_MODE (ALS|SALSA), BSDE_MODEL (blackscholes|hjb),
BSDE_ASSETS, BSDE_BATCH, BSDE_STEPS, BSDE_ITER, BSDE_RANK, BSDE_DEGREE, BSDE_BASIS (polynomial|legendre),
BSDE_PATHS (numpy|philox), BSDE_SEED.
"""
import os
import time
from functools import partial

from bsde_solver import xp as np
from bsde_solver.bsde import HJB, BlackScholes
from bsde_solver.core.calculus.basis import LegendreBasis, PolynomialBasis
from bsde_solver.core.optimisation.als import ALS, SALSA
from bsde_solver.pipeline import predicted_price, run_explicit, simulate_paths
from bsde_solver.utils import flatten


def main():
    mode = os.environ.get("BSDE_MODE", "SALSA")
    if mode == "ALS":
        optimizer = ALS
    elif mode == "SALSA":
        optimizer = partial(SALSA, max_rank=5, optimizer="lstsq")
    else:
        raise ValueError(f"unknown mode {mode}")

    batch_size = int(os.environ.get("BSDE_BATCH", 2000))
    T = 1
    N = int(os.environ.get("BSDE_STEPS", 50))
    num_assets = int(os.environ.get("BSDE_ASSETS", 4))
    dt = T / N
    n_iter = int(os.environ.get("BSDE_ITER", 10))
    rank = int(os.environ.get("BSDE_RANK", 2))
    degree = int(os.environ.get("BSDE_DEGREE", 3))
    ranks = (1,) + (rank,) * (num_assets - 1) + (1,)
    basis = LegendreBasis(degree) if os.environ.get("BSDE_BASIS", "polynomial") == "legendre" else PolynomialBasis(degree)
    seed = os.environ.get("BSDE_SEED")
    if seed is not None:
        np.random.seed(int(seed))

    if os.environ.get("BSDE_MODEL", "blackscholes") == "hjb":
        X0 = np.zeros(num_assets)
        model = HJB(np.tile(X0, (1, 1)), dt, T, sigma=np.sqrt(2))
    else:
        X0 = np.array(flatten([(1, 0.5) for _ in range(num_assets // 2)]) + [1.0] * (num_assets % 2))
        model = BlackScholes(X0, dt, T, 0.05, 0.4)
    configurations = f"{num_assets} assets | {N} steps | {batch_size} batch size | {n_iter} iterations | {degree} degree | {rank} rank"
    print(configurations)

    X, noise = simulate_paths(model, X0, N, batch_size, paths=os.environ.get("BSDE_PATHS", "numpy"), seed=int(seed or 0),
                              keep_noise=False)      # the explicit scheme does not use the increments
    print("Start optimisation")
    start_time = time.perf_counter()
    Y, V, step_times = run_explicit(model, basis, X, n_iter, ranks, optimizer=optimizer, verbose=True,
                                    salsa_no_reg_at_0=(mode == "SALSA"))
    print("End")
    print(f"Time: {time.perf_counter() - start_time:.2f}s")
    print(f"Mean step time: {np.mean(step_times):.2f}s")
    print()
    # ground truth at t = 0: every sample sits at X0, so a handful of rows gives the same mean as the whole batch
    # (the reference evaluates all of them: N * d * 1000 normals for HJB, bsde.py:98-102)
    price = np.mean(model.price(X[0][:, :64].contiguous(), 0))
    predicted = predicted_price(Y)
    print("Price at 0", price)
    print("Predicted Price:", predicted)
    print("Relative error:", f"{np.abs(price - predicted) / price * 100:.2f}%")
    return predicted, price


if __name__ == "__main__":
    main()
