"""Eigen path of 'lstsq' (k_micro_solve) on graded, numerically rank-deficient Gram matrices of Kronecker / monomial
features — the systems that reach it in real fits: accuracy against numpy.linalg.lstsq (fitted values, solution, numerical
rank) and time per solve, two-sided Jacobi (default) against the one-sided SVD (option svd_onesided).  Run on a GPU box:

    python bench/micro/eigen_path_check.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "bsde-tensor-networks_b200", "src"), ROOT): sys.path.insert(0, p)
import numpy as np, torch
from bsde_solver._backend import get_backend
be = get_backend()
rng = np.random.RandomState(0)
def design(n_l, p, n_r, N, spread):
    L = rng.randn(N, n_l) * np.exp(spread * rng.randn(N, 1)); R = rng.randn(N, n_r) * np.exp(spread * rng.randn(N, 1))
    x = 1.0 + 0.3 * rng.randn(N)
    phi = np.stack([x**k for k in range(p)], axis=1)
    V = np.einsum("sa,sm,sb->samb", L, phi, R).reshape(N, -1)
    return V
cases = []
for (rl, p, rr, spread) in [(4, 4, 4, 0.0), (4, 4, 4, 1.0), (4, 4, 4, 3.0), (2, 3, 2, 1.0), (4, 4, 1, 2.0), (5, 4, 5, 1.0), (3, 3, 3, 0.5)]:
    V = design(rl, p, rr, 4096, spread)
    # make columns nearly dependent: duplicate directions with tiny noise
    V[:, -1] = V[:, 0] + 1e-9 * V[:, 1]
    y = np.log(1 + np.sum(V[:, :3] ** 2, axis=1))
    cases.append((f"{rl}x{p}x{rr} spread {spread}", V, y))
for name, V, y in cases:
    A = V.T @ V; A = 0.5 * (A + A.T); b = V.T @ y
    n = A.shape[0]
    ref, _, rank_ref, sv = np.linalg.lstsq(A, b, rcond=None)
    out = {}
    for mode in (0, 1):
        be.set_option("svd_onesided", mode)
        x, info = be.solve(A, b, 0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): x, info = be.solve(A, b, 0)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        fit_err = np.linalg.norm(V @ x - V @ ref) / np.linalg.norm(V @ ref)
        out[mode] = (dt * 1e6, info.tolist(), fit_err, np.linalg.norm(x - ref) / np.linalg.norm(ref))
    print(f"{name:24s} n={n:3d} cond={sv[0]/max(sv[-1],1e-300):.1e} rank_ref={rank_ref} | two-sided {out[0][0]:8.0f} us info {out[0][1]} fit-diff {out[0][2]:.1e} x-diff {out[0][3]:.1e} | one-sided {out[1][0]:8.0f} us info {out[1][1]} fit-diff {out[1][2]:.1e} x-diff {out[1][3]:.1e}", flush=True)
be.set_option("svd_onesided", 0)
