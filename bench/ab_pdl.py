"""Same-process A/B of option "pdl" (programmatic dependent launch of the three kernels of a micro-step inside the captured
sweeps): the C3 fit (HJB law, d = 100, N = 2^20, r = p = 4, 10 sweeps) and the converging fit of bench.py (Black-Scholes law,
d = 20), each timed with the option off, on, off, on; the cores of every fit are compared bit for bit with the first one.
Usage: python bench/ab_pdl.py [fits per leg]     (BSDE_B200_LIB selects the library build)"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch

from bsde_solver._backend import BASIS_POLY, DeviceInputs, get_backend

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
be = get_backend()
P = R = 4
SWEEPS = 10


def workload(kind):
    if kind == "c3":
        d, N = 100, 1 << 20
        x, _ = be.paths_simulate(1, np.array([np.sqrt(2.0)]), np.zeros(d), N, 1, 1.0, seed=2024, keep_xi=False)
        xT = x[1].contiguous()
        y = be.terminal(1, np.array([np.sqrt(2.0)]), xT)
        rng = np.random.RandomState(7)
    else:
        d, N = 20, 1 << 20
        x, _ = be.paths_simulate(0, np.array([0.05, 0.4]), np.tile([1.0, 0.5], d // 2), N, 50, 1.0, seed=4242, keep_xi=False)
        xT = x[50].contiguous()
        y = be.terminal(0, np.array([0.05, 0.4]), xT)
        rng = np.random.RandomState(11)
    del x
    ranks = np.array([1] + [R] * (d - 1) + [1], dtype=np.int32)
    init = np.concatenate([(rng.rand(ranks[i] * P * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
    return DeviceInputs(BASIS_POLY, P, xT), y, ranks, init, d, N


for kind in ("c3", "converging"):
    inp, y, ranks, init, d, N = workload(kind)
    first = None
    for pdl in (0, 1, 0, 1):
        be.set_option("pdl", pdl)
        for _ in range(2):
            cores = init.copy()
            be.fit_als(inp, y, SWEEPS, ranks, 0, cores)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            cores = init.copy()
            info = be.fit_als(inp, y, SWEEPS, ranks, 0, cores)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        digest = hashlib.sha256(cores.tobytes()).hexdigest()[:16]
        if first is None:
            first = digest
        print(json.dumps({"workload": kind, "d": d, "N": N, "pdl": pdl, "ms_per_fit": round(ms, 3),
                          "us_per_micro_step": round(1e3 * ms / (2 * (d - 1) * SWEEPS), 2),
                          "M_samples_sweeps_per_s": round(N * SWEEPS / ms / 1e3, 2), "solves": [int(v) for v in info[:4]],
                          "cores_sha256": digest, "bit_identical_to_first": digest == first,
                          "sweep_graph_edges": [int(be.get_stat("graph_edges")), int(be.get_stat("graph_pdl_edges"))]}), flush=True)
    be.set_option("pdl", 0)
    del inp, y
