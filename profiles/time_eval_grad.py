"""Event-timed TT evaluation (k_tt_eval) and TT gradient (k_interface_right x (d-1) + k_grad_step x d) at the C3 shape
(d = 100, N = 2^20, r = p = 4).  Usage: python profiles/time_eval_grad.py   (BSDE_B200_LIB selects the library build)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

be = get_backend()
for d, p, r, N in ((100, 4, 4, 1 << 20), (20, 4, 4, 1 << 20), (10, 3, 2, 1 << 16)):
    x, _ = be.paths_simulate(1, np.array([np.sqrt(2.0)]), np.zeros(d), N, 1, 1.0, seed=1, keep_xi=False)
    xT = x[1].contiguous()
    ranks = np.array([1] + [r] * (d - 1) + [1], dtype=np.int32)
    rng = np.random.RandomState(0)
    cores = np.concatenate([(rng.rand(ranks[i] * p * ranks[i + 1]) - 0.5) * 0.5 for i in range(d)])
    inp = DeviceInputs(BASIS_POLY, p, xT)
    res = {}
    for name, fn in (("tt_eval", lambda: be.tt_eval(inp, ranks, cores)), ("tt_grad", lambda: be.tt_grad(inp, ranks, cores))):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 10, float(out.double().sum()))
    flops = 2 * sum(int(ranks[i]) * p * int(ranks[i + 1]) for i in range(d)) * N
    print(f"d={d} N={N} r={r} p={p}: tt_eval {res['tt_eval'][0]:.3f} ms ({flops / res['tt_eval'][0] / 1e9:.1f} TFLOP/s, checksum {res['tt_eval'][1]:.12e}), "
          f"tt_grad {res['tt_grad'][0]:.3f} ms (checksum {res['tt_grad'][1]:.12e})")
