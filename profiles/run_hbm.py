"""Driver for ncu captures of the HBM-bound kernels at the C3 shape (d = 100, N = 2^20, r = p = 4): path simulation (Philox),
terminal condition, target formation, TT evaluation, TT gradient (right-interface pass + k_grad_step), and the unfused
interface kernels (first ALS sweep).  Usage: python profiles/run_hbm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

be = get_backend()
d, p, r, N, steps = 100, 4, 4, 1 << 20, 4
sig = np.array([np.sqrt(2.0)])
x, xi = be.paths_simulate(1, sig, np.zeros(d), N, steps, 1.0, seed=1, keep_xi=True)       # k_paths: writes 2 x (steps+1) x d x N doubles
xT = x[steps].contiguous()
y = be.terminal(1, sig, xT)                                                                   # k_terminal
ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
rng = np.random.RandomState(0)
cores = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 0.5 for i in range(d)])
inp = DeviceInputs(BASIS_POLY, p, xT)
v = be.tt_eval(inp, ranks, cores)                                                             # k_tt_eval
z = be.tt_grad(inp, ranks, cores)                                                             # k_interface_right x 99, k_grad_step x 100
t = be.target(1, sig, xT, v, v, z, xi[steps], 0.02, True)                                     # k_target (implicit form: with the Z.xi term)
be.sync()
print("hbm driver done", float(t.mean()))
