"""Per-step diagnostics of the explicit backward solve on the GPU: step time, which solver path every micro-step took
(Cholesky fast path / eigen-decomposition), range of Z, target and Y.

    python bench/diag_solve.py N d steps [hjb|bs] [rank] [p]

Used for DESIGN.md §9 (the reference algorithm diverges on HJB d = 100; Black-Scholes d = 10 converges to the analytic price)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT): sys.path.insert(0, p)
import numpy as np, torch
from bsde_solver._backend import get_backend
from bsde_solver.bsde import HJB, BlackScholes
from bsde_solver.core.calculus.basis import PolynomialBasis
from bsde_solver.core.calculus.derivative import multi_derivative
from bsde_solver.core.optimisation import als as A
from bsde_solver.pipeline import simulate_paths
from bsde_solver.utils import fast_contract
N, d, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
r = p = 4
be = get_backend(); np.random.seed(0)
MODEL = sys.argv[4] if len(sys.argv) > 4 else "hjb"
r = int(sys.argv[5]) if len(sys.argv) > 5 else 4
p = int(sys.argv[6]) if len(sys.argv) > 6 else 4
if MODEL == "hjb":
    X0 = np.zeros(d); model = HJB(np.zeros((1, d)), 1.0 / steps, 1.0, sigma=float(np.sqrt(2.0)))
else:
    X0 = np.array([1.0, 0.5] * (d // 2)); model = BlackScholes(X0[None, :], 1.0 / steps, 1.0, 0.05, 0.4)
    print("analytic price", float(np.exp((0.05 + 0.16) * 1.0) * np.sum(X0**2)))
basis = PolynomialBasis(p); ranks = (1,) + (r,) * (d - 1) + (1,)
X, _ = simulate_paths(model, X0, steps, N, paths="philox", seed=1000, keep_noise=False)
Y = be.empty(steps + 1, N); Y[steps] = model.g(X[steps])
def stat(name, v):
    v = v.flatten()
    return f"{name}: mean {float(v.mean()):.6g} min {float(v.min()):.4g} max {float(v.max()):.4g} nonfinite {int((~torch.isfinite(v)).sum())}"
inp = basis.device_inputs(X[steps])
torch.cuda.synchronize(); t0 = time.perf_counter()
V = A.ALS(inp, Y[steps], n_iter=10, ranks=ranks)
torch.cuda.synchronize(); print(f"step {steps} {time.perf_counter()-t0:.3f}s", dict(A.last_fit_info), stat("fit", fast_contract(V, inp)), stat("y", Y[steps]))
dt = 1.0 / steps
for n in range(steps - 1, -1, -1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Z = multi_derivative(V, basis.device_inputs(X[n + 1]))
    target = be.target(model.model_id, model.params(), X[n + 1], Y[n + 1], Y[n + 1], Z, None, dt, False)
    inp = basis.device_inputs(X[n])
    V = A.ALS(inp, target, n_iter=10, ranks=ranks, init_tt=V)
    Y[n] = fast_contract(V, inp)
    torch.cuda.synchronize()
    cmax = max(float(np.abs(V.cores[f"core_{i}"].view(np.ndarray)).max()) if hasattr(V.cores[f"core_{i}"], "view") else 0 for i in range(d))
    print(f"step {n} {time.perf_counter()-t0:.3f}s", dict(A.last_fit_info), stat("Z", Z), stat("target", target), stat("Y", Y[n]), f"max|core| {cmax:.3g}", flush=True)
print("Y0", float(Y[0].mean()))
