"""Same-process A/B of run-time options on the C3 fit (HJB law, d = 100, N = 2^20, r = p = 4, 10 sweeps): every argument is
one option set "name=value,name=value" (bsde_ctx_set_option), timed in turn for `--rounds` rounds; the cores of every fit
are hashed.  Usage: python bench/ab_opts.py [--fits K] [--rounds R] "pdl=0" "pdl=1,red_groups=6" ...
(BSDE_B200_LIB selects the library build; options keep their value until another set changes them.)"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--fits", type=int, default=4)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("sets", nargs="+")
args = ap.parse_args()
be = get_backend()
P = R = 4
SWEEPS = 10
d, N = 100, 1 << 20
x, _ = be.paths_simulate(1, np.array([np.sqrt(2.0)]), np.zeros(d), N, 1, 1.0, seed=2024, keep_xi=False)
xT = x[1].contiguous()
y = be.terminal(1, np.array([np.sqrt(2.0)]), xT)
del x
rng = np.random.RandomState(7)
ranks = np.array([1] + [R] * (d - 1) + [1], dtype=np.int32)
init = np.concatenate([(rng.rand(ranks[i] * P * ranks[i + 1]) - 0.5) * 2 for i in range(d)])
inp = DeviceInputs(BASIS_POLY, P, xT)
for rnd in range(args.rounds):
    for spec in args.sets:
        for kv in spec.split(","):
            name, _, val = kv.partition("=")
            be.set_option(name, int(val))
        for _ in range(2):
            cores = init.copy()
            be.fit_als(inp, y, SWEEPS, ranks, 0, cores)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.fits):
            cores = init.copy()
            be.fit_als(inp, y, SWEEPS, ranks, 0, cores)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.fits
        print(json.dumps({"lib": os.path.basename(os.environ.get("BSDE_B200_LIB", "libbsde_b200.so")), "options": spec, "round": rnd,
                          "ms_per_fit": round(ms, 3), "us_per_micro_step": round(1e3 * ms / (2 * (d - 1) * SWEEPS), 2),
                          "M_samples_sweeps_per_s": round(N * SWEEPS / ms / 1e3, 2),
                          "cores_sha256": hashlib.sha256(cores.tobytes()).hexdigest()[:16]}), flush=True)
