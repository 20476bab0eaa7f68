"""TEST INFRASTRUCTURE ONLY — CPU (numpy, float64) restatement of the reference hot path.

This module is the *checker* for the CUDA product path.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; nothing under
`bsde-tensor-networks_b200/` does (the product path fails loudly without its CUDA library).

Parity status: **pinned against the reference itself run in the build container** — the
reference publishes no golden vectors or asserting tests, so
`oracle/make_golden.py` imports the unmodified reference from /root/reference on top of
`oracle/shim/` (stand-ins for the absent third-party `opt_einsum` / `matplotlib`), records its
inputs/outputs under fixed numpy seeds into `tests/golden/*.npz`, and `tests/test_oracle_golden.py`
checks every function below against those fixtures.

Each function cites the reference lines it restates (paths relative to /root/reference/src/bsde_solver).
All arithmetic is IEEE float64.  The ALS here caches the left/right interfaces (O(N d) per sweep)
instead of re-contracting the whole chain per micro-step as `retraction_operator` does
(core/optimisation/als.py:8-23); the quantities it produces per micro-step (design matrix V,
normal matrix A, right-hand side, solution, QR factors) are the same mathematical objects.
"""
from __future__ import annotations

import numpy as np
from scipy.special import legendre as _scipy_legendre

# --------------------------------------------------------------------------------------
# bases — core/calculus/basis.py:15-44
# --------------------------------------------------------------------------------------


def poly_eval(x, p):
    """PolynomialBasis.eval (basis.py:20-21): [x**i for i < p] -> (N, p)."""
    return np.array([x**i for i in range(p)]).T


def poly_grad(x, p):
    """PolynomialBasis.grad (basis.py:23-24)."""
    return np.array([np.zeros(x.shape)] + [i * x ** (i - 1) for i in range(1, p)]).T


def poly_dgrad(x, p):
    """PolynomialBasis.dgrad (basis.py:26-27)."""
    return np.array([np.zeros(x.shape), np.zeros(x.shape)] + [i * (i - 1) * x ** (i - 2) for i in range(2, p)]).T


def legendre_eval(x, p):
    """LegendreBasis.eval (basis.py:37-38): polyval of scipy's legendre(i) coefficients."""
    return np.array([np.polyval(_scipy_legendre(i), x) for i in range(p)]).T


def legendre_grad(x, p):
    """LegendreBasis.grad (basis.py:40-41)."""
    return np.array([np.polyval(_scipy_legendre(i).deriv(), x) for i in range(p)]).T


def legendre_dgrad(x, p):
    """LegendreBasis.dgrad (basis.py:43-44)."""
    return np.array([np.polyval(_scipy_legendre(i).deriv().deriv(), x) for i in range(p)]).T


def legendre_coefficients(p):
    """Monomial coefficients (ascending powers) of P_0..P_{p-1} and of their first two derivatives,
    exactly the numbers scipy hands to np.polyval in basis.py:34-44.  Shape (3, p, p)."""
    out = np.zeros((3, p, p))
    for i in range(p):
        for k, poly in enumerate((_scipy_legendre(i), _scipy_legendre(i).deriv(), _scipy_legendre(i).deriv().deriv())):
            c = np.atleast_1d(np.asarray(poly.coeffs, dtype=np.float64))[::-1]
            out[k, i, : len(c)] = c[:p]
    return out


BASES = {
    "polynomial": (poly_eval, poly_grad, poly_dgrad),
    "legendre": (legendre_eval, legendre_grad, legendre_dgrad),
}

# --------------------------------------------------------------------------------------
# tensor train container helpers — core/tensor/tensor_train.py, tensor_core.py
# --------------------------------------------------------------------------------------


def randomize_cores(shape, ranks):
    """TensorTrain.randomize (tensor_train.py:119-122) -> TensorCore.randomize (tensor_core.py:117-118):
    cores filled in order with (rand(*shape) - 0.5) * 2 from the *global* numpy stream."""
    return [(np.random.rand(ranks[i], shape[i], ranks[i + 1]) - 0.5) * 2 for i in range(len(shape))]


def left_unfold(c):
    """tensor_train.py:10."""
    return c.reshape(c.shape[0] * c.shape[1], c.shape[2])


def right_unfold(c):
    """tensor_train.py:13."""
    return c.reshape(c.shape[0], c.shape[1] * c.shape[2])


def orthonormalize(cores, mode="right", start=0, stop=-1):
    """TensorTrain.orthonormalize (tensor_train.py:34-87).  In place on the list `cores`."""
    d = len(cores)
    if stop == -1:
        stop = d - 1
    if mode == "left":
        for k in range(start, stop):
            L = left_unfold(cores[k])
            R = right_unfold(cores[k + 1])
            V, U = np.linalg.qr(L)
            W = U @ R
            cores[k] = V.reshape(cores[k].shape)
            cores[k + 1] = W.reshape(cores[k + 1].shape)
    elif mode == "right":
        for k in range(stop, start, -1):
            L = left_unfold(cores[k - 1])
            R = right_unfold(cores[k])
            V, U = np.linalg.qr(R.T)
            W = L @ U.T
            cores[k - 1] = W.reshape(cores[k - 1].shape)
            cores[k] = V.T.reshape(cores[k].shape)
    else:
        raise ValueError("orthogonality must be either 'left' or 'right'")
    return cores


# --------------------------------------------------------------------------------------
# per-sample contractions — utils.py:19-27, derivative.py:43-85, als.py:8-23
# --------------------------------------------------------------------------------------


def core_times_phi(core, phi):
    """G_i[s,a,b] = sum_m C_i[a,m,b] phi_i[s,m]   (als.py:14-17, utils.py:24)."""
    return np.einsum("amb,sm->sab", core, phi, optimize=True)


def left_interfaces(cores, phis, upto=None):
    """L_j[s,:] = (G_0 ... G_{j-1})[s] as (N, r_j); L_0 = ones((N,1)).  Returns list for j = 0..upto."""
    d = len(cores)
    upto = d - 1 if upto is None else upto
    N = phis[0].shape[0]
    out = [np.ones((N, 1))]
    for j in range(upto):
        out.append(np.einsum("sa,sab->sb", out[-1], core_times_phi(cores[j], phis[j]), optimize=True))
    return out


def right_interfaces(cores, phis, downto=0):
    """R_j[s,:] = (G_{j+1} ... G_{d-1})[s] as (N, r_{j+1}); R_{d-1} = ones((N,1)).  dict j -> array."""
    d = len(cores)
    N = phis[0].shape[0]
    out = {d - 1: np.ones((N, 1))}
    for j in range(d - 2, downto - 1, -1):
        out[j] = np.einsum("sab,sb->sa", core_times_phi(cores[j + 1], phis[j + 1]), out[j + 1], optimize=True)
    return out


def design_matrix(Lj, phij, Rj):
    """V[s,(a,m,b)] = L_j[s,a] phi_j[s,m] R_j[s,b] in C order (r_j, m, r_{j+1}) — the batch-unfolded
    retraction operator of als.py:21-22,53."""
    N = phij.shape[0]
    return np.einsum("sa,sm,sb->samb", Lj, phij, Rj, optimize=True).reshape(N, -1)


def tt_eval(cores, phis):
    """fast_contract (utils.py:19-27): V(x_s) for every sample -> (N,)."""
    L = left_interfaces(cores, phis, upto=len(cores) - 1)[-1]
    G = core_times_phi(cores[-1], phis[-1])
    return np.einsum("sa,sab->sb", L, G, optimize=True)[:, 0]


def tt_grad(cores, phis, dphis):
    """multi_derivative (derivative.py:43-85): d/dx_i V(x_s) = left_i[s] (C_i . dphi_i[s]) right_i[s] -> (N, d)."""
    d = len(cores)
    N = phis[0].shape[0]
    Ls = left_interfaces(cores, phis, upto=d - 1)
    Rs = right_interfaces(cores, phis, downto=0)
    out = np.zeros((N, d))
    for i in range(d):
        dG = core_times_phi(cores[i], dphis[i])
        out[:, i] = np.einsum("sa,sab,sb->s", Ls[i], dG, Rs[i], optimize=True)
    return out


def tt_hessian(cores, phis, dphis, ddphis):
    """hessian (core/calculus/hessian.py:6-26) with batch=True: H[i, j, s] = the train contracted with the basis values of
    every asset except i and j, which get the first-derivative values (i != j) or asset i the second-derivative values."""
    d = len(cores)
    N = phis[0].shape[0]
    G = [core_times_phi(cores[k], phis[k]) for k in range(d)]         # (N, r, r')
    dG = [core_times_phi(cores[k], dphis[k]) for k in range(d)]
    ddG = [core_times_phi(cores[k], ddphis[k]) for k in range(d)]
    H = np.zeros((d, d, N))
    for i in range(d):
        for j in range(i, d):
            v = np.ones((N, 1))
            for k in range(d):
                M = ddG[k] if (k == i and k == j) else (dG[k] if k in (i, j) else G[k])
                v = np.einsum("sa,sab->sb", v, M)
            H[i, j] = H[j, i] = v[:, 0]
    return H


# --------------------------------------------------------------------------------------
# small dense solve — core/optimisation/solve.py:6-19
# --------------------------------------------------------------------------------------


def solver(A, b, method="lu"):
    """solve.py:6-19.  'lstsq' = LAPACK gelsd min-norm with rcond = eps * n (numpy rcond=None)."""
    from scipy import linalg

    if method == "lu":
        return linalg.lu_solve(linalg.lu_factor(A), b)
    if method == "lstsq":
        return np.linalg.lstsq(A, b, rcond=None)[0]
    if method == "qr":
        q, r = np.linalg.qr(A)
        return np.linalg.solve(r, q.T @ b)
    if method == "solve":
        return np.linalg.solve(A, b)
    raise ValueError(f"Unknown method {method}")


# --------------------------------------------------------------------------------------
# ALS — core/optimisation/als.py:26-96
# --------------------------------------------------------------------------------------


def als(phis, y, n_iter=10, ranks=None, optimizer="lstsq", cores=None, randomize=True, trace=None):
    """ALS (als.py:26-96).  `cores` plays the role of init_tt (only its shapes survive: als.py:36-37
    re-randomises it).  With randomize=False the given cores are used as the post-randomize state
    (how the CUDA path is fed identical initial cores).  `trace`, if a list, receives one dict per
    micro-step with A, b, x (the solve) for kernel-level parity."""
    d = len(phis)
    y = np.asarray(y, dtype=np.float64)
    shape = tuple(phi.shape[1] for phi in phis)
    if cores is None:
        cores = [np.zeros((ranks[i], shape[i], ranks[i + 1])) for i in range(d)]
    else:
        cores = [np.array(c, dtype=np.float64) for c in cores]
        ranks = [c.shape[0] for c in cores] + [cores[-1].shape[2]]
    if randomize:
        cores = randomize_cores(shape, ranks)  # als.py:37
    orthonormalize(cores, "right", start=1)  # als.py:39

    Rs = right_interfaces(cores, phis, downto=0)
    Ls = [np.ones((y.shape[0], 1))] + [None] * (d - 1)

    def micro(j):  # als.py:44-59
        V = design_matrix(Ls[j], phis[j], Rs[j])
        A = V.T @ V
        b = V.T @ y
        x = solver(A, b, optimizer)
        if trace is not None:
            trace.append({"j": j, "A": A, "b": b, "x": x.copy()})
        return x.reshape(cores[j].shape)

    for _ in range(n_iter):
        for j in range(d - 1):  # left half sweep, als.py:63-77
            X = micro(j)
            Q, S = np.linalg.qr(left_unfold(X))
            W = S @ right_unfold(cores[j + 1])
            cores[j] = Q.reshape(cores[j].shape)
            cores[j + 1] = W.reshape(cores[j + 1].shape)
            Ls[j + 1] = np.einsum("sa,sab->sb", Ls[j], core_times_phi(cores[j], phis[j]), optimize=True)
        for j in range(d - 1, 0, -1):  # right half sweep, als.py:80-94
            X = micro(j)
            Q, S = np.linalg.qr(right_unfold(X).T)
            W = left_unfold(cores[j - 1]) @ S.T
            cores[j - 1] = W.reshape(cores[j - 1].shape)
            cores[j] = Q.T.reshape(cores[j].shape)
            Rs[j - 1] = np.einsum("sab,sb->sa", core_times_phi(cores[j], phis[j]), Rs[j], optimize=True)
    return cores


# --------------------------------------------------------------------------------------
# SALSA — core/optimisation/als.py:99-323 (literal, quirks included)
# --------------------------------------------------------------------------------------


def _inv_sv(S, min_sv):
    """als.py:260,277: diag entries nan_to_num(1 / max(S, min_sv), posinf=1e-6, neginf=1e-6)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.nan_to_num(1 / np.maximum(S, min_sv), posinf=1e-6, neginf=1e-6)


def salsa_rank_adaptation(U, S, G, min_sv, shape_prev, shape_curr):
    """rank_adaptation (als.py:185-211).  U: left_unfold of core_{j-1} factor, G = Vt @ R."""
    S = np.append(S, 0.01 * min_sv)
    U = U.reshape(shape_prev)
    U_comp = np.ones((U.shape[0], U.shape[1]))
    U_comp -= np.einsum("abc,dec,de->ab", U, U, U_comp)
    U_comp -= np.einsum("abc,dec,de->ab", U, U, U_comp)
    U_norm = np.linalg.norm(U_comp)
    if U_norm != 0:
        U_comp /= U_norm
        U = np.concatenate((U, U_comp[:, :, None]), axis=2)
    V = G.reshape(shape_curr)
    V_comp = np.ones((V.shape[1], V.shape[2]))
    V_comp -= np.einsum("abc,ade,de->bc", V, V, V_comp)
    V_comp -= np.einsum("abc,ade,de->bc", V, V, V_comp)
    V_norm = np.linalg.norm(V_comp)
    if V_norm != 0:
        V_comp /= V_norm
        V = np.concatenate((V, V_comp[None, :, :]), axis=0)
    W = np.diag(S) @ V.reshape((V.shape[0], -1))
    return U, S, W, U.shape, V.shape


def salsa_regulariser(n, r_l, p, r_r, j, d, gamma, theta, eta):
    """reg_term of als.py:143-166 including the flat-reshape layout of the theta term."""
    reg = np.zeros((n, n))
    if gamma is not None and theta is not None:
        if j != 0:
            reg += eta**2 * np.einsum("ab,bc,ij,kl->aikcjl", gamma, gamma, np.eye(p), np.eye(r_r)).reshape(n, n)
        if j != d - 1:
            reg += eta**2 * np.einsum("ab,bc,ij,kl->aikcjl", theta, theta, np.eye(p), np.eye(r_l)).reshape(n, n)
    return reg


def salsa_move_left_bond(core_prev, core_curr, min_sv, expand, max_rank_bond, do_reg):
    """als.py:240-260 for one j != 0: SVD of left_unfold(core_{j-1}), optional +1 rank.
    Returns (core_prev, core_curr, gamma_diag or None, adapted: bool)."""
    L = left_unfold(core_prev)
    R = right_unfold(core_curr)
    U, S, Vt = np.linalg.svd(L, full_matrices=False)
    W = S * Vt @ R  # als.py:248: (S*Vt) scales the COLUMNS of Vt
    shape_prev, shape_curr = core_prev.shape, core_curr.shape
    adapted = False
    if expand and S[-1] > min_sv and S.shape[0] < max_rank_bond:
        G = Vt @ R
        U, S, W, shape_prev, shape_curr = salsa_rank_adaptation(U, S, G, min_sv, shape_prev, shape_curr)
        adapted = True
    core_prev = np.asarray(U).reshape(shape_prev)
    core_curr = W.reshape(shape_curr)
    gamma = _inv_sv(S, min_sv) if do_reg else None
    return core_prev, core_curr, gamma, adapted


def salsa_move_right_bond(core_curr, core_next, min_sv, do_reg):
    """als.py:264-277 for one j != d-1."""
    L = left_unfold(core_curr)
    R = right_unfold(core_next)
    U, S, Vt = np.linalg.svd(L, full_matrices=False)
    W = U * S
    Z = Vt @ R
    theta = _inv_sv(S, min_sv) if do_reg else None
    return W.reshape(core_curr.shape), Z.reshape(core_next.shape), theta


def salsa(phis, y, n_iter=10, ranks=None, omega=1.0, freq_omega=1.05, min_sv=0.2, cores=None,
          max_rank=10, optimizer="lstsq", do_reg=True, randomize=True, init_ranks=None, trace=None):
    """SALSA (als.py:99-323).  `init_ranks` is the reference's stale `tt.ranks` attribute (the ranks
    the TensorTrain object was *constructed* with, als.py:122-123) — defaults to the shapes of `cores`."""
    d = len(phis)
    y = np.asarray(y, dtype=np.float64)
    N = y.shape[0]
    shape = tuple(phi.shape[1] for phi in phis)
    if cores is None:
        cores = [np.zeros((ranks[i], shape[i], ranks[i + 1])) for i in range(d)]
        cur_ranks = list(ranks)
    else:
        cores = [np.array(c, dtype=np.float64) for c in cores]
        cur_ranks = [c.shape[0] for c in cores] + [cores[-1].shape[2]]
    if init_ranks is None:
        init_ranks = list(ranks) if ranks is not None else cur_ranks
    if randomize:
        cores = randomize_cores(shape, cur_ranks)  # als.py:116

    eta = 1e-6  # als.py:121
    max_ranks = [min(init_ranks[i] * shape[i], max_rank) for i in range(d - 1)]  # als.py:122
    max_ranks[-1] = min(init_ranks[-1] * shape[-1], max_rank)  # als.py:123

    expand_rank, has_adapted, warmup, last_adapt = False, False, 2, 0  # als.py:213-216
    V = None
    for it in range(n_iter):
        if it >= last_adapt + warmup:
            expand_rank = True  # als.py:224
        if has_adapted:
            expand_rank, has_adapted, last_adapt = False, False, it  # als.py:225
        orthonormalize(cores, "right", start=1)  # als.py:234
        Ls = [np.ones((N, 1))]
        # cores j+2.. are untouched until micro-step j+1, so R_{j+1} of the sweep-start cores can be cached
        Rcache = right_interfaces(cores, phis, downto=0)
        for j in range(d):  # als.py:235
            gamma = theta = None
            if j != 0:
                cores[j - 1], cores[j], gamma, adapted = salsa_move_left_bond(
                    cores[j - 1], cores[j], min_sv, expand_rank, max_ranks[j - 1], do_reg)
                has_adapted = has_adapted or adapted
                # left interface up to core j-1 (its final value for this sweep)
                Lprev = Ls[j - 1]
                Ls.append(np.einsum("sa,sab->sb", Lprev, core_times_phi(cores[j - 1], phis[j - 1]), optimize=True))
            if j != d - 1:
                cores[j], cores[j + 1], theta = salsa_move_right_bond(cores[j], cores[j + 1], min_sv, do_reg)
            # micro_optimization, als.py:125-183 — right interface from the *current* core_{j+1}
            if j != d - 1:
                Rj = np.einsum("sab,sb->sa", core_times_phi(cores[j + 1], phis[j + 1]), Rcache[j + 1], optimize=True)
            else:
                Rj = Rcache[d - 1]
            V = design_matrix(Ls[j], phis[j], Rj)
            r_l, p, r_r = cores[j].shape
            n = r_l * p * r_r
            A = V.T @ V
            g = np.diag(gamma) if gamma is not None else None
            t = np.diag(theta) if theta is not None else None
            A = A + salsa_regulariser(n, r_l, p, r_r, j, d, g, t, eta)
            if j == 0 and it == 0:  # als.py:168-172
                ev = V @ cores[j].reshape(-1)
                eta = min(1 / N * np.linalg.norm(ev - y) ** 2, 10000)
            A = A + eta * np.eye(n)
            b = V.T @ y
            x = solver(A, b, optimizer)
            if trace is not None:
                trace.append({"it": it, "j": j, "A": A, "b": b, "x": x.copy(), "eta": eta,
                              "ranks": [c.shape[0] for c in cores] + [1]})
            cores[j] = x.reshape(cores[j].shape)
        # als.py:288-298
        ev = V @ cores[d - 1].reshape(-1)
        eta_tmp = eta
        eta = 1 / N * np.linalg.norm(ev - y) ** 2
        rel_eta = abs(eta_tmp - eta) / eta_tmp
        omega = min(omega / freq_omega, max(eta, np.sqrt(eta)))
        omega = max(omega, rel_eta)
        min_sv = max(min(0.2 * omega, 0.2 * rel_eta), 0)
        if trace is not None:
            trace.append({"it": it, "sweep_end": True, "eta": eta, "omega": omega, "min_sv": min_sv,
                          "expand_rank": expand_rank, "has_adapted": has_adapted, "last_adapt": last_adapt})
    return cores


# --------------------------------------------------------------------------------------
# models — bsde.py:34-55, 80-102 ; paths — stochastic/path.py:6-34
# --------------------------------------------------------------------------------------


class BlackScholes:
    """bsde.py:34-55.  sigma(x) = sigma0 diag(x) (dense (N,d,d) in the reference; the diagonal closed
    form is bit-identical because the off-diagonals are exact zeros)."""
    name = "blackscholes"

    def __init__(self, T, r, sigma):
        self.T, self.r, self.sigma_ = T, r, sigma

    def sigma_diag(self, x):
        return self.sigma_ * (x * 1.0)

    def h(self, x, t, y, z):
        return -self.r * (y - np.sum(x * z, axis=1))

    def g(self, x):
        return np.linalg.norm(x, axis=1) ** 2

    def price(self, X, t):
        return np.exp((self.r + self.sigma_**2) * (self.T - t)) * np.mean(self.g(X))


class HJB:
    """bsde.py:80-102.  sigma = sigma0 * I."""
    name = "hjb"

    def __init__(self, T, sigma):
        self.T, self.sigma_ = T, sigma

    def sigma_diag(self, x):
        return self.sigma_ * np.ones_like(x)

    def h(self, x, t, y, z):
        return -0.5 * np.sum(z**2, axis=1)

    def g(self, x):
        return np.log(0.5 + 0.5 * np.sum(x**2, axis=1))


def generate_trajectories(X0, n_steps, model, xi=None):
    """path.py:6-34 with b = 0 and diagonal sigma: x_n = x_{n-1} + sqrt(dt) * (sigma_diag(x_{n-1}) * xi_n).
    xi ~ randn(N, n_steps+1, d) from the global stream unless supplied (xi[:,0] unused)."""
    N, d = X0.shape
    dt = model.T / n_steps
    if xi is None:
        xi = np.random.randn(N, n_steps + 1, d)
    x = np.zeros((N, n_steps + 1, d))
    x[:, 0] = X0
    for n in range(1, n_steps + 1):
        x[:, n] = x[:, n - 1] + np.sqrt(dt) * (model.sigma_diag(x[:, n - 1]) * xi[:, n])
    return x, xi


# --------------------------------------------------------------------------------------
# backward schemes — src/scripts/pipeline.py:84-141, experiments/n_assets.py:62-198
# --------------------------------------------------------------------------------------


def explicit_scheme(X, model, basis="polynomial", p=3, ranks=None, n_iter=10, optimizer_name="ALS",
                    apply_sigma=False, salsa_kwargs=None, record=None):
    """Explicit backward loop of src/scripts/pipeline.py:84-141 (Z = grad V, sigma not applied) or of
    experiments/n_assets.py:62-126 (apply_sigma=True).  Returns Y (N, steps+1) and the final cores."""
    ev, gr, _ = BASES[basis]
    N, T1, d = X.shape
    n_steps = T1 - 1
    dt = model.T / n_steps
    salsa_kwargs = dict(salsa_kwargs or {})

    def fit(phis, target, cores, n):
        if optimizer_name == "ALS":
            return als(phis, target, n_iter=n_iter, ranks=ranks, cores=cores)
        kw = dict(salsa_kwargs)
        if n == 0:
            kw["do_reg"] = False  # pipeline.py:135-136
        return salsa(phis, target, n_iter=n_iter, ranks=ranks, cores=cores, init_ranks=list(ranks), **kw)

    Y = np.zeros((N, n_steps + 1))
    Y[:, -1] = model.g(X[:, -1])
    phis = [ev(X[:, -1, i], p) for i in range(d)]
    cores = fit(phis, Y[:, -1], None, n_steps)
    for n in range(n_steps - 1, -1, -1):
        phis1 = [ev(X[:, n + 1, i], p) for i in range(d)]
        dphis1 = [gr(X[:, n + 1, i], p) for i in range(d)]
        Z = tt_grad(cores, phis1, dphis1)
        if apply_sigma:
            Z = model.sigma_diag(X[:, n + 1]) * Z
        h = model.h(X[:, n + 1], (n + 1) * dt, Y[:, n + 1], Z)
        target = h * dt + Y[:, n + 1]
        phis = [ev(X[:, n, i], p) for i in range(d)]
        cores = fit(phis, target, cores, n)
        Y[:, n] = tt_eval(cores, phis)
        if record is not None:
            record.append({"n": n, "target": target, "Z": Z})
    return Y, cores


def implicit_scheme(X, xi, model, basis="polynomial", p=3, ranks=None, n_iter=20, n_iter_implicit=10,
                    optimizer_name="SALSA", salsa_kwargs=None):
    """Implicit backward loop (fixed-point iterations) of src/scripts/pipeline_explicit.py:74-112 in the
    intended form of experiments/n_assets.py:155-191 (`V_nk =` on both branches)."""
    ev, gr, _ = BASES[basis]
    N, T1, d = X.shape
    n_steps = T1 - 1
    dt = model.T / n_steps
    salsa_kwargs = dict(salsa_kwargs or {})

    def fit(phis, target, cores, n):
        if optimizer_name == "ALS":
            return als(phis, target, n_iter=n_iter, ranks=ranks, cores=cores)
        kw = dict(salsa_kwargs)
        if n == 0:
            kw["do_reg"] = False
        return salsa(phis, target, n_iter=n_iter, ranks=ranks, cores=cores, init_ranks=list(ranks), **kw)

    Y = np.zeros((N, n_steps + 1))
    Y[:, -1] = model.g(X[:, -1])
    phis = [ev(X[:, -1, i], p) for i in range(d)]
    cores = fit(phis, Y[:, -1], None, n_steps)
    for n in range(n_steps - 1, -1, -1):
        phis = [ev(X[:, n, i], p) for i in range(d)]
        dphis = [gr(X[:, n, i], p) for i in range(d)]
        Y_nk = Y[:, n + 1]
        for _ in range(n_iter_implicit):
            Z = model.sigma_diag(X[:, n]) * tt_grad(cores, phis, dphis)
            h = model.h(X[:, n], n * dt, Y_nk, Z)
            target = h * dt + Y[:, n + 1] - np.sqrt(dt) * np.einsum("ij,ij->i", Z, xi[:, n + 1])
            cores = fit(phis, target, cores, n)
            Y_nk = tt_eval(cores, phis)
        Y[:, n] = Y_nk
    return Y, cores


# --------------------------------------------------------------------------------------
# Philox4x32-10 counter-based increments — no reference counterpart: restates the stream of the CUDA path
# kernel (bsde-tensor-networks_b200/csrc/model.cu: philox_normal_pair) so that tests can pin it bit for bit.
# --------------------------------------------------------------------------------------


def _philox4x32_10(c, k0, k1):
    """c: (..., 4) uint32 counters, k0/k1 scalar keys -> (..., 4) uint32 (Salmon et al., SC'11)."""
    c = np.array(c, dtype=np.uint64)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[..., 1] ^ k0, lo1, hi0 ^ c[..., 3] ^ k1, lo0], axis=-1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint64)


def philox_increments(seed, n_samples, d, steps, sample_offset=0):
    """xi [steps+1][d][N] exactly as k_paths draws it.  One Philox block per (sample, asset, odd step n) with
    counter = (sample lo, sample hi, asset, n >> 1) and key = (seed lo, seed hi) yields two uniforms of 53 bits;
    Box-Muller's cosine branch is the increment of step n, its sine branch the increment of step n + 1.
    xi[0] = 0 (never used, path.py:34)."""
    xi = np.zeros((steps + 1, d, n_samples))
    s = np.arange(n_samples, dtype=np.uint64) + np.uint64(sample_offset)
    lo, hi = s & np.uint64(0xFFFFFFFF), s >> np.uint64(32)
    for n in range(1, steps + 1, 2):
        for j in range(d):
            ctr = np.stack([lo, hi, np.full_like(s, j), np.full_like(s, n >> 1)], axis=-1)
            r = _philox4x32_10(ctr, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
            a = (r[:, 1] << np.uint64(32)) | r[:, 0]
            b = (r[:, 3] << np.uint64(32)) | r[:, 2]
            u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
            u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            rad = np.sqrt(-2.0 * np.log(u1))
            xi[n, j] = rad * np.cos(2 * np.pi * u2)
            if n + 1 <= steps:
                xi[n + 1, j] = rad * np.sin(2 * np.pi * u2)
    return xi
