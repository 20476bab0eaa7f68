"""TEST / BENCH INFRASTRUCTURE ONLY (oracle/): times the UNMODIFIED reference ALS
(bsde_solver.core.optimisation.als.ALS, /root/reference/src/bsde_solver/core/optimisation/als.py:26-96) as installed by
`pip install --no-index --no-deps --target baseline/_ref /root/reference`, on top of oracle/shim's stand-ins
for the absent third-party packages (opt_einsum, matplotlib).  Used by `bench.py --impl reference` only; never by the
product path.  The reference's solve.py:4 imports `src.bsde_solver...`, so the installed package is also aliased as
`src.bsde_solver` (the reference itself is not touched).

    python oracle/run_reference.py --d 100 --N 1024 --rank 4 --degree 4 --sweeps 1 --steps 2 --warmup 1   -> one JSON line
"""
import argparse
import importlib
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "bsde_solver", "core", "optimisation", "als.py"))


def import_reference():
    """The reference package, first on sys.path; refuses to run if this repo's own `bsde_solver` is already imported."""
    if "bsde_solver" in sys.modules and not os.path.abspath(sys.modules["bsde_solver"].__file__).startswith(REF):
        raise RuntimeError("another bsde_solver is already imported in this process")
    os.environ.setdefault("PYTHONHASHSEED", "0")
    for p in (REF, os.path.join(HERE, "shim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    pkg = importlib.import_module("bsde_solver")
    src = types.ModuleType("src")
    src.__path__ = []
    sys.modules.setdefault("src", src)
    sys.modules.setdefault("src.bsde_solver", pkg)
    for sub in ("core", "core.tensor", "core.tensor.tensor_network", "core.tensor.tensor_core", "core.tensor.tensor_train"):
        sys.modules.setdefault("src.bsde_solver." + sub, importlib.import_module("bsde_solver." + sub))
    return pkg


def time_als(d, N, rank, degree, sweeps, steps, warmup, seed=0):
    """ALS(phis, y, n_iter=sweeps, ranks) of the reference on the C3 workload's law: x ~ N(0, 2 I), y = log(1/2 + |x|^2 / 2),
    monomial basis values as the reference's PolynomialBasis produces them.  Returns seconds per step (list)."""
    import numpy as np

    import_reference()
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.optimisation.als import ALS
    from bsde_solver.core.tensor.tensor_core import TensorCore

    rng = np.random.RandomState(seed)
    x = rng.randn(N, d) * np.sqrt(2.0)
    y = np.log(0.5 + 0.5 * np.sum(x ** 2, axis=1))
    basis = PolynomialBasis(degree)
    phis = [TensorCore(basis.eval(x[:, i]), name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i in range(d)]
    ranks = (1,) + (rank,) * (d - 1) + (1,)
    np.random.seed(seed)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ALS(phis, y, n_iter=sweeps, ranks=ranks)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--rank", type=int, default=4)
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref not installed"}))
        sys.exit(0)
    t = time_als(a.d, a.N, a.rank, a.degree, a.sweeps, a.steps, a.warmup)
    print(json.dumps({"seconds_per_step": t, "samples_sweeps_per_s": a.N * a.sweeps * len(t) / sum(t)}))
