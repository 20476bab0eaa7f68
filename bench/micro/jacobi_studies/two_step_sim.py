"""numpy study: second Cholesky step before the one-sided Jacobi (L^T L = L2 L2^T, Jacobi on the columns of L2, G = L G2 Sigma^-1)."""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(_HERE)))
sys.path.insert(0, os.path.join(_ROOT, "oracle")); sys.path.insert(0, _HERE)
import sys, numpy as np
import tt_oracle as O
from warm_start_sim import pivchol, jacobi, solve, eps

def main():
    d, N, r, p, sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 6, 1 << 14, 4, 4, 6
    rng = np.random.RandomState(0)
    x = np.exp(0.2 * rng.randn(N, d) - 0.02); y = (x * x).sum(1)
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ranks = [1] + [r] * (d - 1) + [1]
    np.random.seed(1); tr = []
    cores = O.als(phis, y, n_iter=sweeps, ranks=ranks, trace=tr)
    t1 = t2 = t3 = cnt = 0
    for it, t in enumerate(tr):
        A, b, j = t["A"], t["b"], t["j"]; n = A.shape[0]
        A = 0.5 * (A + A.T); nrm = 2.0 ** -np.floor(np.log2(np.diag(A).max())); A = A * nrm; b = b * nrm
        ev = np.linalg.eigvalsh(A); rank_np = int((ev > eps * n * ev.max()).sum())
        if rank_np == n and ev.min() > 100 * eps * n * ev.max(): continue
        xr = np.linalg.lstsq(A, b, rcond=None)[0]
        L, piv, _ = pivchol(A)
        G = L.copy(); W = np.eye(L.shape[1]); s1, r1 = jacobi(G, W, n, np.diag(A).max()); x1, k1 = solve(G, b, n)
        B = L.T @ L; L2, _, _ = pivchol(B)
        G2 = L2.copy(); W2 = np.eye(L2.shape[1]); s2, r2 = jacobi(G2, W2, n, np.diag(B).max())
        sig2 = (G2 * G2).sum(0)                      # eigenvalues of L^T L = of A
        Gp = (L @ G2) / np.sqrt(sig2)                # orthogonal columns with norm sigma
        x2, k2 = solve(Gp, b, n)
        # three steps
        B3 = L2.T @ L2; L3, _, _ = pivchol(B3); G3 = L3.copy(); W3 = np.eye(L3.shape[1]); s3, r3 = jacobi(G3, W3, n, np.diag(B3).max())
        fe = lambda xx: np.linalg.norm(A @ xx - b) / np.linalg.norm(b)
        print(f"visit {it:3d} core {j} n={n} r={L.shape[1]}/{L2.shape[1]}/{L3.shape[1]} rank np {rank_np} 1-step {k1} 2-step {k2} | sweeps 1-step {s1} 2-step {s2} 3-step {s3} {r2}| dx1 {np.linalg.norm(x1-xr)/np.linalg.norm(xr):.1e} dx2 {np.linalg.norm(x2-xr)/np.linalg.norm(xr):.1e} res {fe(xr):.1e} {fe(x1):.1e} {fe(x2):.1e}")
        t1 += s1; t2 += s2; t3 += s3; cnt += 1
    print("systems", cnt, "sweeps 1-step", t1, "2-step", t2, "3-step", t3)
main()
