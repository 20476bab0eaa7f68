"""Small driver for ncu captures: one ALS fit at the C3 core shape (r = p = 4, N = 2^20) on a short chain (d = 8),
so that a profiler run stays short.  Usage: python profiles/run_fit.py [sweeps] [generic]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
be = get_backend()
if len(sys.argv) > 2 and sys.argv[2] == "generic":
    be.set_option("asm_generic", 1)
d, p, r, N = 8, 4, 4, 1 << 20
x, _ = be.paths_simulate(1, np.array([np.sqrt(2.0)]), np.zeros(d), N, 1, 1.0, seed=1, keep_xi=False)
xT = x[1].contiguous()
y = be.terminal(1, np.array([np.sqrt(2.0)]), xT)
ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
rng = np.random.RandomState(0)
cores = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
be.set_option("use_graph", 0)
for kv in sys.argv[3:]:          # further library options, name=value
    name, _, val = kv.partition("=")
    be.set_option(name, int(val))
info = be.fit_als(DeviceInputs(BASIS_POLY, p, xT), y, sweeps, ranks, 0, cores)
print("fit done", info.tolist(), "launches", be.launch_count())
