"""GPU parity of the individual kernels (through the C ABI) against the numpy oracle and the golden fixtures
recorded from the reference (oracle/make_golden.py).  Tolerances are relative, float64."""
import numpy as np
import pytest

from util import cores_from, flat, ranks_of, rel, unflat

pytestmark = pytest.mark.gpu


def _inputs(backend, x, kind, p, coef=None):
    """x (N, d) host -> DeviceInputs with X [d][N]."""
    from bsde_solver._backend import DeviceInputs

    return DeviceInputs(kind, p, backend.to_device(np.ascontiguousarray(x.T)), coef=coef)


def _phi_inputs(backend, phis, dphis=None):
    from bsde_solver._inputs import as_inputs

    return as_inputs(phis, dphis)


# ---------------------------------------------------------------------------------------------- basis
@pytest.mark.parametrize("p", [3, 4, 6])
def test_polynomial_basis_matches_reference(backend, golden, p):
    from bsde_solver.core.calculus.basis import PolynomialBasis

    g = golden("basis.npz")
    b = PolynomialBasis(p)
    for name, fn in (("eval", b.eval), ("grad", b.grad), ("dgrad", b.dgrad)):
        got = fn(g["x"])
        ref = g[f"poly_{name}_{p}"]
        assert got.shape == ref.shape
        # numpy's x**k is the platform pow (about 1 % of the fixture's entries are 1 ulp off the correctly rounded
        # value); the device rounds an exact double-double product once
        np.testing.assert_allclose(got, ref, rtol=2.3e-16, atol=0)
        assert np.mean(got == ref) > 0.98
    # the device monomials are the correctly rounded powers, bit for bit
    from fractions import Fraction

    got = b.eval(g["x"])
    exact = np.array([[float(Fraction(float(v)) ** k) for k in range(p)] for v in g["x"]])
    np.testing.assert_array_equal(got, exact)


@pytest.mark.parametrize("p", [3, 5])
def test_legendre_basis_matches_reference(backend, golden, p):
    from bsde_solver.core.calculus.basis import LegendreBasis

    g = golden("basis.npz")
    b = LegendreBasis(p)
    for name, fn in (("eval", b.eval), ("grad", b.grad), ("dgrad", b.dgrad)):
        np.testing.assert_array_equal(fn(g["x"]), g[f"leg_{name}_{p}"])      # same Horner sequence: bit-exact


# ---------------------------------------------------------------------------------------------- QR chain
@pytest.mark.parametrize("case,mode,start,stop", [("right1_", "right", 1, -1), ("right0_", "right", 0, -1),
                                                  ("left0_", "left", 0, -1), ("left13_", "left", 1, 3)])
def test_orthonormalize_matches_reference(backend, golden, case, mode, start, stop):
    g = golden("orthonormalize.npz")
    cores = cores_from(g, "in_")
    ranks = ranks_of(cores)
    # the fixture's train has its own basis size on every core (shape (3, 4, 3, 3, 4)), as TensorTrain(shape, ranks) allows
    shape = [c.shape[1] for c in cores]
    assert len(set(shape)) > 1
    out = backend.orthonormalize(shape, ranks, flat(cores), 1 if mode == "right" else 0, start, stop)
    ref = cores_from(g, case)
    assert rel(out, flat(ref)) < 1e-13
    # and through the reference-shaped class
    from bsde_solver.core.tensor.tensor_core import TensorCore
    from bsde_solver.core.tensor.tensor_train import TensorTrain

    tt = TensorTrain(shape, list(ranks))
    for i, c in enumerate(cores):
        tt.cores[f"core_{i}"] = TensorCore.like(c, tt.cores[f"core_{i}"])
    tt.orthonormalize(mode=mode, start=start, stop=stop)
    assert rel(flat([np.asarray(c) for c in tt.cores.values()]), flat(ref)) < 1e-13


def test_orthonormalize_against_oracle(backend):
    import oracle.tt_oracle as O

    rng = np.random.RandomState(3)
    for p, ranks in ((3, [1, 2, 3, 3, 2, 1]), (4, [1, 4, 4, 4, 4, 4, 1]), (4, [1, 3, 5, 2, 1])):
        cores = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(len(ranks) - 1)]
        for mode, start, stop in (("right", 1, -1), ("right", 0, -1), ("left", 0, -1), ("left", 1, len(cores) - 2)):
            ref = O.orthonormalize([c.copy() for c in cores], mode, start, stop)
            out = backend.orthonormalize(p, np.array(ranks, dtype=np.int32), flat(cores), 1 if mode == "right" else 0, start, stop)
            assert rel(out, flat(ref)) < 1e-13, (p, ranks, mode)
    # orthonormality identities the reference's scripts eyeball (tests/core/test_tensor_network.py:141-187)
    out = unflat(backend.orthonormalize(4, np.array([1, 4, 4, 4, 4, 4, 1], dtype=np.int32),
                                        flat([rng.rand(r0, 4, r1) for r0, r1 in zip([1, 4, 4, 4, 4, 4], [4, 4, 4, 4, 4, 1])]), 1, 0, -1),
                 [1, 4, 4, 4, 4, 4, 1], 4)
    for c in out[1:]:
        m = c.reshape(c.shape[0], -1)
        np.testing.assert_allclose(m @ m.T, np.eye(c.shape[0]), atol=1e-14)


# ---------------------------------------------------------------------------------------------- eval / grad
@pytest.mark.parametrize("name", ["poly", "leg"])
def test_eval_and_grad_match_reference(backend, golden, name):
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis, PolynomialBasis

    g = golden("eval_grad.npz")
    p = int(g["p"])
    cores = cores_from(g, "core_")
    basis = PolynomialBasis(p) if name == "poly" else LegendreBasis(p)
    inp = _inputs(backend, g["x"], BASIS_POLY if name == "poly" else BASIS_LEGENDRE, p, coef=basis._coef())
    y = backend.tt_eval(inp, ranks_of(cores), flat(cores)).cpu().numpy()
    z = backend.tt_grad(inp, ranks_of(cores), flat(cores)).cpu().numpy().T
    assert rel(y, g[f"eval_{name}"]) < 1e-13
    assert rel(z, g[f"grad_{name}"]) < 1e-13
    # the reference's calling convention: explicit basis-value lists
    phis = [basis.eval(g["x"][:, i]) for i in range(g["x"].shape[1])]
    dphis = [basis.grad(g["x"][:, i]) for i in range(g["x"].shape[1])]
    inp2 = _phi_inputs(backend, phis, dphis)
    assert rel(backend.tt_eval(inp2, ranks_of(cores), flat(cores)).cpu().numpy(), g[f"eval_{name}"]) < 1e-13
    assert rel(backend.tt_grad(inp2, ranks_of(cores), flat(cores)).cpu().numpy().T, g[f"grad_{name}"]) < 1e-13


# ---------------------------------------------------------------------------------------------- normal equations
@pytest.mark.parametrize("d,p,r,N", [(5, 3, 2, 1000), (6, 4, 4, 4097), (4, 4, 3, 333), (3, 5, 2, 64), (4, 3, 1, 100)])
def test_normal_equations_match_oracle(backend, d, p, r, N):
    """A = V^T V and b = V^T y of every core against the oracle's design matrix (als.py:44-55)."""
    import oracle.tt_oracle as O
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis

    rng = np.random.RandomState(d * 100 + p)
    ranks = [1] + [r] * (d - 1) + [1]
    cores = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    x = rng.randn(N, d) * 0.7
    y = rng.randn(N)
    yd = backend.to_device(y)
    for kind in ("poly", "leg", "phi"):
        if kind == "leg":
            phis = [O.legendre_eval(x[:, i], p) for i in range(d)]
            inp = _inputs(backend, x, BASIS_LEGENDRE, p, coef=LegendreBasis(p)._coef())
        else:
            phis = [O.poly_eval(x[:, i], p) for i in range(d)]
            inp = _inputs(backend, x, BASIS_POLY, p) if kind == "poly" else _phi_inputs(backend, phis)
        Ls = O.left_interfaces(cores, phis)
        Rs = O.right_interfaces(cores, phis)
        for j in range(d):
            V = O.design_matrix(Ls[j], phis[j], Rs[j])
            A, b = backend.normal_eq(inp, yd, ranks, flat(cores), j)
            assert rel(A, V.T @ V) < 1e-13, (kind, j)
            assert rel(b, V.T @ y) < 1e-12, (kind, j)


# ---------------------------------------------------------------------------------------------- solver
def test_solver_matches_numpy_lstsq(backend):
    rng = np.random.RandomState(7)
    for n in (1, 6, 12, 27, 64, 100):
        M = rng.randn(3 * n + 5, n)
        A = M.T @ M
        b = rng.randn(n)
        for method, sid in (("lstsq", 0), ("lu", 1), ("qr", 2), ("solve", 3)):
            x, info = backend.solve(A, b, sid)
            ref = np.linalg.lstsq(A, b, rcond=None)[0]
            assert rel(x, ref) < 1e-9 * max(1.0, np.linalg.cond(A) / 1e4), (n, method)
    # rank-deficient: lstsq's minimum-norm solution (the step-0 fit, every sample identical)
    for n in (6, 16, 64):
        v = rng.randn(n)
        A = 2000.0 * np.outer(v, v)
        b = 3.7 * v
        x, info = backend.solve(A, b, 0)
        ref = np.linalg.lstsq(A, b, rcond=None)[0]
        assert info[1] == 1 and info[3] == 1          # jacobi path, numerical rank 1
        assert rel(x, ref) < 1e-12
    # rank n-2
    n = 20
    M = rng.randn(n - 2, n)
    A = M.T @ M
    b = A @ rng.randn(n)
    x, info = backend.solve(A, b, 0)
    assert info[3] == n - 2
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-9


def test_solver_beyond_the_shared_memory_of_one_cta(backend):
    """solver(A, b, method) (solve.py:6-19) on systems of up to 512 unknowns: beyond ~140 the matrix lives in the solver's
    global-memory workspace.  Positive definite and rank-deficient Gram matrices against numpy.linalg.lstsq, the three
    elimination methods against numpy.linalg.solve; what only the small-system fallbacks serve says so."""
    from bsde_solver._lib import BackendError
    from bsde_solver.core.optimisation.solve import solver

    rng = np.random.RandomState(17)
    for n in (150, 256, 400, 512):
        M = rng.randn(2 * n + 5, n)
        A = M.T @ M
        b = rng.randn(n)
        ref = np.linalg.solve(A, b)
        for method in ("lstsq", "lu", "qr", "solve"):
            assert rel(solver(A, b, method), ref) < 1e-9 * max(1.0, np.linalg.cond(A) / 1e4), (n, method)
    n = 200                                        # numerical rank n - 5: lstsq's minimum-norm solution
    M = rng.randn(n - 5, n)
    A = M.T @ M
    b = A @ rng.randn(n)
    x, info = backend.solve(A, b, 0)
    assert info[1] == 1 and info[3] == n - 5
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-8
    G = rng.randn(160, 160)                        # not symmetric: general lstsq, served up to 100 unknowns only
    with pytest.raises(BackendError):
        solver(G, rng.randn(160), "lstsq")
    np.testing.assert_allclose(solver(G, G @ np.ones(160), "lu"), np.ones(160), rtol=1e-8)
    with pytest.raises(BackendError):
        backend.solve(np.eye(513), np.ones(513), 0)


def test_lstsq_eigen_path_symmetric_and_general_input(backend):
    """'lstsq' beyond the Cholesky gate: symmetric systems (every micro-step) go through the two-sided Jacobi
    eigen-decomposition — also indefinite ones, singular values are |lambda| — and non-symmetric input through the
    one-sided Jacobi SVD.  Graded, numerically rank-deficient Gram matrices of monomial features are the systems that
    actually reach this path (truncation at eps * n * s_max as numpy.linalg.lstsq, solve.py:11)."""
    rng = np.random.RandomState(9)
    # graded Gram matrix with a nearly dependent column: fitted values are what is well defined
    for n_l, p, n_r in ((4, 4, 4), (2, 3, 2), (3, 3, 3)):
        N = 4096
        L = rng.randn(N, n_l) * np.exp(rng.randn(N, 1))
        R = rng.randn(N, n_r) * np.exp(rng.randn(N, 1))
        xs = 1.0 + 0.3 * rng.randn(N)
        V = np.einsum("sa,sm,sb->samb", L, np.stack([xs**k for k in range(p)], axis=1), R).reshape(N, -1)
        V[:, -1] = V[:, 0] + 1e-9 * V[:, 1]
        A = V.T @ V
        A = 0.5 * (A + A.T)
        b = V.T @ np.log(1.0 + (V[:, :3] ** 2).sum(axis=1))
        ref, _, rank_ref, _ = np.linalg.lstsq(A, b, rcond=None)
        x, info = backend.solve(A, b, 0)
        assert info[1] == 1 and info[3] == rank_ref, (A.shape[0], info.tolist(), rank_ref)
        assert rel(V @ x, V @ ref) < 1e-9
        try:                                            # the one-sided SVD on the same system
            backend.set_option("svd_onesided", 1)
            x1, info1 = backend.solve(A, b, 0)
        finally:
            backend.set_option("svd_onesided", 0)
        assert info1[3] == rank_ref and rel(V @ x1, V @ ref) < 1e-9
    # symmetric indefinite, exactly singular (rank n - 3)
    n = 24
    Q, _ = np.linalg.qr(rng.randn(n, n))
    lam = np.concatenate([rng.uniform(0.5, 3.0, n - 3) * rng.choice([-1.0, 1.0], n - 3), np.zeros(3)])
    A = (Q * lam) @ Q.T
    A = 0.5 * (A + A.T)
    b = A @ rng.randn(n)
    x, info = backend.solve(A, b, 0)
    assert info[3] == n - 3
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-10
    # non-symmetric square system, rank n - 1
    n = 17
    M = rng.randn(n, n)
    M[:, -1] = M[:, 0] - 2.0 * M[:, 3]
    b = M @ rng.randn(n)
    x, info = backend.solve(M, b, 0)
    assert info[3] == n - 1
    assert rel(x, np.linalg.lstsq(M, b, rcond=None)[0]) < 1e-10


def test_solver_rejects_unknown_method(backend):
    from bsde_solver._lib import BackendError
    from bsde_solver.core.optimisation.solve import solver

    with pytest.raises(ValueError):
        solver(np.eye(2), np.ones(2), "cholesky?")
    with pytest.raises(BackendError):
        backend.solve(np.eye(2), np.ones(2), 9)


# ---------------------------------------------------------------------------------------------- paths, model
def test_paths_replay_is_bit_exact(backend, golden):
    g = golden("paths.npz")
    for tag, mid, params, x0 in (("bs", 0, [0.05, 0.4], [1.0, 0.5, 1.0, 0.5]), ("hjb", 1, [np.sqrt(2)], [0.0] * 4)):
        x_ref, xi = g[f"{tag}_x"], g[f"{tag}_xi"]
        N, T1, d = x_ref.shape
        xi_dev = backend.to_device(np.ascontiguousarray(xi.transpose(1, 2, 0)))
        x, _ = backend.paths_simulate(mid, np.array(params), np.array(x0), N, T1 - 1, 1.0, xi=xi_dev)
        np.testing.assert_array_equal(x.permute(2, 0, 1).cpu().numpy(), x_ref)


def test_paths_from_per_sample_starting_points(backend):
    """generate_trajectories takes a general (batch, dim) X0 (path.py:6-34; every reference script tiles one row): the
    replay of host increments from per-sample starting points against the oracle's Euler scheme, bit for bit, through
    the reference-shaped entry point and the device-resident one."""
    import oracle.tt_oracle as O
    from bsde_solver.bsde import HJB, BlackScholes
    from bsde_solver.stochastic.path import generate_trajectories, generate_trajectories_device

    rng = np.random.RandomState(21)
    N, d, steps = 777, 5, 6
    X0 = 0.5 + rng.rand(N, d)
    for model, omodel in ((BlackScholes(X0, 1.0 / steps, 1.0, 0.05, 0.4), O.BlackScholes(1.0, 0.05, 0.4)),
                          (HJB(X0, 1.0 / steps, 1.0, np.sqrt(2)), O.HJB(1.0, np.sqrt(2)))):
        np.random.seed(4)
        x, xi = generate_trajectories(X0, steps, model)
        assert x.shape == (N, steps + 1, d)
        x_ref, _ = O.generate_trajectories(X0, steps, omodel, xi=xi)
        np.testing.assert_array_equal(x, x_ref)
        np.testing.assert_array_equal(x[:, 0], X0)
        xd, _ = generate_trajectories_device(X0, steps, model, seed=9)
        np.testing.assert_array_equal(xd[0].cpu().numpy(), X0.T)
        assert np.isfinite(xd.cpu().numpy()).all() and not np.array_equal(xd[1].cpu().numpy(), X0.T)
    with pytest.raises(ValueError):
        backend.paths_simulate(1, np.array([1.0]), X0, N + 1, steps, 1.0)


def test_model_h_g_match_reference(backend, golden):
    from bsde_solver.bsde import HJB, BlackScholes

    g = golden("paths.npz")
    X0 = np.zeros((64, 4))
    hjb = HJB(X0, 0.1, 1, np.sqrt(2))
    bs = BlackScholes(X0, 0.1, 1, 0.05, 0.4)
    np.testing.assert_allclose(hjb.h(g["hjb_x"][:, 3], 0.3, g["yv"], g["z"]), g["hjb_h"], rtol=1e-14)
    np.testing.assert_allclose(hjb.g(g["hjb_x"][:, -1]), g["hjb_g"], rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(bs.h(g["bs_x"][:, 3], 0.3, g["yv"], g["z"]), g["bs_h"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(bs.g(g["bs_x"][:, -1]), g["bs_g"], rtol=1e-14)


def test_philox_paths_statistics_and_shard_invariance(backend):
    """Counter-based increments: a shard regenerates exactly its slice of the global stream; moments are N(0,1)."""
    N, d, steps = 1 << 14, 6, 9
    x, xi = backend.paths_simulate(1, np.array([np.sqrt(2)]), np.zeros(d), N, steps, 1.0, seed=1234)
    xa, xia = backend.paths_simulate(1, np.array([np.sqrt(2)]), np.zeros(d), N // 2, steps, 1.0, seed=1234, sample_offset=N // 2)
    assert (xia == xi[:, :, N // 2:]).all() and (xa == x[:, :, N // 2:]).all()
    z = xi[1:].cpu().numpy().ravel()
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 0.01 and abs((z**4).mean() - 3) < 0.1
    c = np.corrcoef(xi[1].cpu().numpy())            # assets independent
    assert np.abs(c - np.eye(d)).max() < 0.05
    # terminal law: X_T ~ N(0, 2 T I)
    assert abs(x[-1].cpu().numpy().var() - 2.0) < 0.05


def test_transpose_and_mean(backend):
    import torch

    a = torch.randn(1000, 7, dtype=torch.float64, device="cuda")
    assert (backend.transpose(a, 1000, 7) == a.T).all()
    v = torch.randn(123457, dtype=torch.float64, device="cuda")
    assert abs(backend.mean(v) - v.mean().item()) < 1e-15


def test_philox_stream_matches_the_restated_generator(backend):
    """The device's counter-based increments against the numpy restatement of the same Philox4x32-10 + Box-Muller
    stream: the integer part is exact, log / sincospi differ from numpy's by a few ulp."""
    import oracle.tt_oracle as O

    N, d, steps = 777, 3, 7
    _, xi = backend.paths_simulate(1, np.array([1.0]), np.zeros(d), N, steps, 1.0, seed=(5 << 32) + 12345, sample_offset=1 << 33)
    ref = O.philox_increments((5 << 32) + 12345, N, d, steps, sample_offset=1 << 33)
    np.testing.assert_allclose(xi.cpu().numpy(), ref, rtol=0, atol=5e-15)


# ---------------------------------------------------------------------------------------------- Hessian, PDE residual
@pytest.mark.parametrize("name", ["poly", "leg"])
def test_hessian_matches_reference(backend, golden, name):
    """bsde_tt_hessian against the reference's hessian() (core/calculus/hessian.py:6-26): fused bases from x, the reference's
    calling convention (lists of basis / derivative / second-derivative values, batch=True and a single point)."""
    from bsde_solver.core.calculus.basis import LegendreBasis, PolynomialBasis
    from bsde_solver.core.calculus.hessian import hessian
    from bsde_solver.core.tensor.tensor_core import TensorCore
    from bsde_solver.core.tensor.tensor_train import TensorTrain

    g = golden("hessian.npz")
    x = g["x"]
    N, d = x.shape
    p = int(g[f"p_{name}"])
    cores = cores_from(g, f"cores_{name}_")
    basis = PolynomialBasis(p) if name == "poly" else LegendreBasis(p)
    tt = TensorTrain([p] * d, ranks_of(cores).tolist())
    for i, c in enumerate(cores):
        tt.cores[f"core_{i}"] = TensorCore.like(c, tt.cores[f"core_{i}"])
    Hd = hessian(tt, basis.device_inputs(backend.to_device(np.ascontiguousarray(x.T)))).cpu().numpy()
    assert rel(Hd, g[f"H_{name}"]) < 1e-13
    lists = [[TensorCore(f(x[:, i]), name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i in range(d)] for f in (basis.eval, basis.grad, basis.dgrad)]
    assert rel(hessian(tt, *lists, batch=True), g[f"H_{name}"]) < 1e-13
    one = [[f(x[3:4, i])[0] for i in range(d)] for f in (basis.eval, basis.grad, basis.dgrad)]
    assert rel(hessian(tt, *one, batch=False), g[f"H1_{name}"]) < 1e-13


def test_pde_loss_vanishes_on_the_exact_solution(backend):
    """PDELoss (loss.py:10-33) on the closed-form Black-Scholes-Barenblatt solution v = exp((r + sigma^2)(T - t)) |x|^2
    (bsde.py:54-55): the residual d_t v + 1/2 Tr(sigma sigma^T D^2 v) + h must vanish; with the Hessian of a train that
    represents |x|^2 exactly (rank 2, 3 monomials) it does to rounding."""
    from bsde_solver.bsde import BlackScholes
    from bsde_solver.core.calculus.basis import PolynomialBasis
    from bsde_solver.core.calculus.derivative import multi_derivative
    from bsde_solver.core.calculus.hessian import hessian
    from bsde_solver.core.tensor.tensor_train import TensorTrain
    from bsde_solver.loss import PDELoss
    from bsde_solver.utils import fast_contract

    rng = np.random.RandomState(1)
    N, d, p = 300, 4, 3
    x = 0.5 + rng.rand(N, d)
    r, sig, T, t = 0.05, 0.4, 1.0, 0.3
    model = BlackScholes(x[:1], 0.02, T, r, sig)
    # |x|^2 as a rank-2 train: core_k carries (running sum, 1)
    cores = []
    for k in range(d):
        rl, rr = (1 if k == 0 else 2), (1 if k == d - 1 else 2)
        c = np.zeros((rl, p, rr))
        if k == 0:
            c[0, 2, 0], c[0, 0, 1] = 1.0, 1.0            # (x^2, 1)
        elif k == d - 1:
            c[0, 0, 0], c[1, 2, 0] = 1.0, 1.0            # sum + x^2
        else:
            c[0, 0, 0], c[1, 2, 0], c[1, 0, 1] = 1.0, 1.0, 1.0
        cores.append(c)
    tt = TensorTrain.from_tensor(np.zeros([p] * d), [1] + [2] * (d - 1) + [1]) if False else TensorTrain([p] * d, [1] + [2] * (d - 1) + [1])
    from bsde_solver.core.tensor.tensor_core import TensorCore
    for i, c in enumerate(cores):
        tt.cores[f"core_{i}"] = TensorCore.like(c, tt.cores[f"core_{i}"])
    inp = PolynomialBasis(p).device_inputs(backend.to_device(np.ascontiguousarray(x.T)))
    scale = np.exp((r + sig**2) * (T - t))
    v = scale * fast_contract(tt, inp).cpu().numpy()
    np.testing.assert_allclose(v, scale * (x**2).sum(axis=1), rtol=1e-13)
    vx = scale * multi_derivative(tt, inp).cpu().numpy().T
    vxx = scale * hessian(tt, inp).cpu().numpy().transpose(2, 0, 1)
    vt = -(r + sig**2) * v
    loss = PDELoss(model)(t, x, v, vt, vx, vxx)
    assert np.abs(loss).max() < 1e-12 * np.abs(v).max()
