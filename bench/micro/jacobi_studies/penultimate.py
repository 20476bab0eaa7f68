"""max cos^2 per sweep of the one-sided Jacobi on truncated systems of a converging fit: can the verification sweep go?"""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(_HERE)))
sys.path.insert(0, os.path.join(_ROOT, "oracle")); sys.path.insert(0, _HERE)
import sys, numpy as np
import tt_oracle as O
from warm_start_sim import pivchol, rr_pairs, eps

def jacobi_trace(G, n, dmax0):
    r = G.shape[1]; m = (r + 1) & ~1
    negl = 1e-2 * eps * n * dmax0; sub = 0.25 * eps * n * dmax0
    out = []
    for sweep in range(30):
        mx = 0.0; need = False
        for ps in rr_pairs(m):
            for p, q in ps:
                if q >= r: continue
                al = G[:, p] @ G[:, p]; be = G[:, q] @ G[:, q]; ga = G[:, p] @ G[:, q]
                if al <= negl and be <= negl: continue
                c2 = ga * ga / (al * be)
                if not (al <= sub and be <= sub): mx = max(mx, c2)
                if c2 > (1e-4 if (al <= sub and be <= sub) else 1e-14): need = True
                if c2 > eps * eps:
                    w = be - al; g = 2 * ga; hyp = np.hypot(w, g); a = abs(w) + hyp
                    rn = 1 / np.hypot(a, g); c = a * rn; s = np.copysign(g * rn, g * w if w != 0 else g)
                    gp, gq = G[:, p].copy(), G[:, q].copy()
                    G[:, p] = c * gp - s * gq; G[:, q] = s * gp + c * gq
        out.append(mx)
        if not need: break
    return out

d, N, r, p, sweeps = 6, 1 << 14, 4, 4, 6
rng = np.random.RandomState(0)
x = np.exp(0.2 * rng.randn(N, d) - 0.02); y = (x * x).sum(1)
phis = [O.poly_eval(x[:, i], p) for i in range(d)]
np.random.seed(1); tr = []
O.als(phis, y, n_iter=sweeps, ranks=[1] + [r] * (d - 1) + [1], trace=tr)
for it, t in enumerate(tr):
    A, b = t["A"], t["b"]; n = A.shape[0]
    A = 0.5 * (A + A.T); A = A * 2.0 ** -np.floor(np.log2(np.diag(A).max()))
    ev = np.linalg.eigvalsh(A)
    if ev.min() > 100 * eps * n * ev.max(): continue
    L, _, _ = pivchol(A)
    tr_ = jacobi_trace(L.copy(), n, np.diag(A).max())
    print(it, n, L.shape[1], len(tr_), " ".join(f"{v:.0e}" for v in tr_))
