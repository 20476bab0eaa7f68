"""numpy study: does a warm start (W of the previous solve of the same core, pivot order replayed) cut the sweeps of the
one-sided Jacobi phase of vh_lstsq on the systems of a converging ALS fit?  CPU only."""
import os, sys
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(_HERE)))
sys.path.insert(0, os.path.join(_ROOT, "oracle")); sys.path.insert(0, _HERE)
import sys, numpy as np
import tt_oracle as O
eps = 2.220446049250313e-16

def pivchol(A, order=None):
    """diagonally pivoted Cholesky, rows unpermuted; returns L (n x r), the pivot order, and whether `order` was followed"""
    n = A.shape[0]; S = A.copy(); d = np.diag(S).copy(); dmax0 = d.max(); tol = eps / 16 * dmax0
    sel = np.zeros(n, bool); cols = []; piv = []; followed = order is not None
    for k in range(n):
        dm = np.where(sel, -np.inf, d)
        jk = int(np.argmax(dm))
        if followed:
            if k < len(order) and not sel[order[k]] and dm[order[k]] >= 0.125 * dm[jk] and dm[order[k]] > tol: jk = order[k]
            else: followed = False
        if not dm[jk] > tol: break
        l = S[:, jk] / np.sqrt(dm[jk]); l[sel] = 0.0
        S -= np.outer(l, l); d = np.diag(S).copy(); sel[jk] = True
        cols.append(l); piv.append(jk)
    if order is not None and len(piv) != len(order): followed = False
    return np.array(cols).T, piv, followed

def rr_pairs(m):
    out = []
    for step in range(m - 1):
        ps = []
        for k in range(m // 2):
            if k == 0: p, q = m - 1, step
            else: p, q = (step + k) % (m - 1), (step - k + m - 1) % (m - 1)
            ps.append((min(p, q), max(p, q)))
        out.append(ps)
    return out

def jacobi(G, W, n, dmax0, lenient=False):
    r = G.shape[1]; m = (r + 1) & ~1
    negl = 1e-2 * eps * n * dmax0
    sweeps = 0; rots_per_sweep = []
    for sweep in range(30):
        need = False; rots = 0
        for ps in rr_pairs(m):
            for p, q in ps:
                if q >= r: continue
                al = G[:, p] @ G[:, p]; be = G[:, q] @ G[:, q]; ga = G[:, p] @ G[:, q]
                if al <= negl and be <= negl: continue
                sub = lenient and al <= 0.25 * eps * n * dmax0 and be <= 0.25 * eps * n * dmax0
                if ga * ga > (1e-4 if sub else 1e-14) * al * be: need = True
                if ga * ga > eps * eps * al * be:
                    w = be - al; g = 2 * ga; hyp = np.hypot(w, g); a = abs(w) + hyp
                    rn = 1 / np.hypot(a, g); c = a * rn; s = np.copysign(g * rn, g * w if w != 0 else g)
                    gp, gq = G[:, p].copy(), G[:, q].copy()
                    G[:, p] = c * gp - s * gq; G[:, q] = s * gp + c * gq
                    wp, wq = W[:, p].copy(), W[:, q].copy()
                    W[:, p] = c * wp - s * wq; W[:, q] = s * wp + c * wq
                    if ga * ga > 1e-14 * al * be: rots += 1
        sweeps += 1; rots_per_sweep.append(rots)
        if not need: break
    return sweeps, rots_per_sweep

def solve(G, b, n):
    lam = (G * G).sum(0); cut = eps * n * lam.max()
    keep = lam > cut
    coef = np.where(keep, (G.T @ b) / np.where(keep, lam * lam, 1), 0.0)
    return G @ coef, int(keep.sum())

def main():
    d, N, r, p, sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 1 << 14, 4, 4, 8
    rng = np.random.RandomState(0)
    x = np.exp(0.2 * rng.randn(N, d) - 0.02)
    y = (x * x).sum(1)
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ranks = [1] + [r] * (d - 1) + [1]
    np.random.seed(1)
    tr = []
    cores = O.als(phis, y, n_iter=sweeps, ranks=ranks, trace=tr)
    print("residual", np.linalg.norm(O.tt_eval(cores, phis) - y) / np.linalg.norm(y))
    state = {}
    tot_cold = tot_warm = cnt = 0
    for it, t in enumerate(tr):
        A, b, j = t["A"], t["b"], t["j"]; n = A.shape[0]
        A = 0.5 * (A + A.T); nrm = 2.0 ** -np.floor(np.log2(np.diag(A).max())); A = A * nrm; b = b * nrm
        ev = np.linalg.eigvalsh(A); rank_np = int((ev > eps * n * ev.max()).sum())
        if rank_np == n and ev.min() > 100 * eps * n * ev.max(): continue      # Cholesky fast path
        L, piv, _ = pivchol(A)
        G = L.copy(); W = np.eye(L.shape[1])
        sc, rc = jacobi(G, W, n, np.diag(A).max())
        xc, rkc = solve(G, b, n)
        line = f"visit {it:4d} core {j} n={n} r={L.shape[1]} rank np {rank_np} ours {rkc} cold sweeps {sc} {rc}"
        if j in state:
            Wp, pivp = state[j]
            L2, piv2, followed = pivchol(A, pivp)
            if followed and L2.shape[1] == Wp.shape[0]:
                G2 = L2 @ Wp; W2 = Wp.copy()
                sw, rw = jacobi(G2, W2, n, np.diag(A).max())
                xw, rkw = solve(G2, b, n)
                line += f" | warm sweeps {sw} {rw} rank {rkw} dx {np.linalg.norm(xw-xc)/np.linalg.norm(xc):.1e}"
                tot_cold += sc; tot_warm += sw; cnt += 1
                state[j] = (W2, piv2)
            else:
                line += " | warm n/a (pivot order / rank changed)"
                state[j] = (W, piv)
        else:
            state[j] = (W, piv)
        print(line)
    print("truncated systems with a warm start:", cnt, "cold sweeps", tot_cold, "warm sweeps", tot_warm)
if __name__ == "__main__":
    main()
