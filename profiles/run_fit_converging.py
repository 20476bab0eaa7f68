"""Driver for ncu captures of the truncated solve path: one ALS fit that converges (Black-Scholes law, y = |x|^2, d = 8,
r = p = 4, N = 2^20), so that most micro-steps of the later sweeps are numerically rank-deficient and k_micro_solve_fast
falls through into vh_lstsq.  Usage: python profiles/run_fit_converging.py [sweeps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
be = get_backend()
d, p, r, N = 8, 4, 4, 1 << 20
par = np.array([0.05, 0.4])
x, _ = be.paths_simulate(0, par, np.tile([1.0, 0.5], d // 2), N, 10, 1.0, seed=1, keep_xi=False)
xT = x[10].contiguous()
y = be.terminal(0, par, xT)
ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
rng = np.random.RandomState(0)
cores = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
be.set_option("use_graph", 0)
info = be.fit_als(DeviceInputs(BASIS_POLY, p, xT), y, sweeps, ranks, 0, cores)
print("fit done", info.tolist(), "launches", be.launch_count())
