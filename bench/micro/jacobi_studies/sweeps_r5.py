import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(_HERE)))
sys.path.insert(0, os.path.join(_ROOT, "oracle")); sys.path.insert(0, _HERE)
import sys, pickle, numpy as np
from warm_start_sim import pivchol, jacobi, solve, eps
for tag, A, b, rk in pickle.load(open(os.path.join(_ROOT, "scratch", "corpus.pkl"), "rb"))[::3]:
    n = A.shape[0]
    nrm = 2.0 ** -np.floor(np.log2(np.diag(A).max())); A = A * nrm; b = b * nrm
    L, piv, _ = pivchol(A)
    G = L.copy(); W = np.eye(L.shape[1])
    s, rots = jacobi(G, W, n, np.diag(A).max())
    lam = np.sort((G * G).sum(0))[::-1]
    x, k = solve(G, b, n)
    G2 = L.copy(); s2, rots2 = jacobi(G2, W, n, np.diag(A).max(), lenient=True); x2, k2 = solve(G2, b, n)
    fit = lambda u, v: np.sqrt(abs((u - v) @ A @ (u - v)) / (v @ A @ v))
    xr = np.linalg.lstsq(A, b, rcond=None)[0]
    print("   lenient: sweeps", s2, "rank", k2, "fit-err vs strict %.1e" % fit(x2, x), "strict vs lstsq %.1e" % fit(x, xr), "lenient vs lstsq %.1e" % fit(x2, xr))
    print(tag, "n", n, "r", L.shape[1], "rank", k, "ref", rk, "sweeps", s, rots, "lam[rank-2:rank+3]/cut", (lam[max(k-2,0):k+3] / (eps * n * lam[0])).round(2))
