"""Driver for ncu captures and timing of the drop-in convention's hot kernels at the C3 core shape (r = p = 4, N = 2^20, d = 8):
Legendre basis evaluated in the kernels (BASIS_LEGENDRE) or explicit basis values (BASIS_PHI, what ALS(phis, y) with anonymous
host arrays turns into).  Usage: python profiles/run_fit_legendre.py [legendre|phi] [sweeps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np

from bsde_solver._backend import BASIS_LEGENDRE, BASIS_PHI, BASIS_POLY, DeviceInputs, get_backend
from bsde_solver.core.calculus.basis import LegendreBasis

mode = sys.argv[1] if len(sys.argv) > 1 else "legendre"
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
be = get_backend()
d, p, r, N = 8, 4, 4, 1 << 20
x, _ = be.paths_simulate(1, np.array([0.5]), np.zeros(d), N, 1, 1.0, seed=1, keep_xi=False)
xT = x[1].contiguous()
y = be.terminal(1, np.array([0.5]), xT)
coef = LegendreBasis(p)._coef()
if mode == "legendre":
    inp = DeviceInputs(BASIS_LEGENDRE, p, xT, coef=coef)
else:
    import torch
    phi = torch.stack([be.basis_eval(BASIS_LEGENDRE, p, xT[j], coef, 0) for j in range(d)]).contiguous()   # [d][p][N] explicit values
    inp = DeviceInputs(BASIS_PHI, p, phi)
ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
rng = np.random.RandomState(0)
cores = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
be.set_option("use_graph", 0)
be.fit_als(inp, y, 1, ranks, 0, cores.copy())
be.set_option("profile", 1)
info = be.fit_als(inp, y, sweeps, ranks, 0, cores.copy())
stats = {k: (be.get_stat(f"{k}_ms"), int(be.get_stat(f"{k}_n"))) for k in ("assembly", "reduce", "solve", "interface")}
be.set_option("profile", 0)
print(mode, "fit done", info.tolist(), {k: (round(1e3 * v[0] / max(1, v[1]), 2), v[1]) for k, v in stats.items()})
