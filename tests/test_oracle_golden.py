"""CPU: the numpy oracle (oracle/tt_oracle.py) against the fixtures recorded from the UNMODIFIED reference
(oracle/make_golden.py, run where /root/reference exists).  This is what pins the checker the GPU tests use."""
import numpy as np
import pytest

import oracle.tt_oracle as O
from util import cores_from, rel


def test_basis_bit_exact(golden):
    g = golden("basis.npz")
    x = g["x"]
    for p in (3, 4, 6):
        np.testing.assert_array_equal(O.poly_eval(x, p), g[f"poly_eval_{p}"])
        np.testing.assert_array_equal(O.poly_grad(x, p), g[f"poly_grad_{p}"])
        np.testing.assert_array_equal(O.poly_dgrad(x, p), g[f"poly_dgrad_{p}"])
    for p in (3, 5):
        np.testing.assert_array_equal(O.legendre_eval(x, p), g[f"leg_eval_{p}"])
        np.testing.assert_array_equal(O.legendre_grad(x, p), g[f"leg_grad_{p}"])
        np.testing.assert_array_equal(O.legendre_dgrad(x, p), g[f"leg_dgrad_{p}"])


def test_legendre_coefficients_reproduce_polyval(golden):
    g = golden("basis.npz")
    x = g["x"]
    c = O.legendre_coefficients(5)                 # ascending powers
    for k, name in enumerate(("eval", "grad", "dgrad")):
        val = sum(c[k][:, q][None, :] * x[:, None] ** q for q in range(5))
        np.testing.assert_allclose(val, g[f"leg_{name}_5"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("case,mode,start,stop", [("right1_", "right", 1, -1), ("right0_", "right", 0, -1),
                                                  ("left0_", "left", 0, -1), ("left13_", "left", 1, 3)])
def test_orthonormalize(golden, case, mode, start, stop):
    g = golden("orthonormalize.npz")
    out = O.orthonormalize(cores_from(g, "in_"), mode, start, stop)
    for a, b in zip(out, cores_from(g, case)):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-14)


@pytest.mark.parametrize("name,ev,gr", [("poly", O.poly_eval, O.poly_grad), ("leg", O.legendre_eval, O.legendre_grad)])
def test_eval_and_grad(golden, name, ev, gr):
    g = golden("eval_grad.npz")
    p, x, cores = int(g["p"]), g["x"], cores_from(g, "core_")
    phis = [ev(x[:, i], p) for i in range(x.shape[1])]
    dphis = [gr(x[:, i], p) for i in range(x.shape[1])]
    assert rel(O.tt_eval(cores, phis), g[f"eval_{name}"]) < 1e-14
    assert rel(O.tt_grad(cores, phis, dphis), g[f"grad_{name}"]) < 1e-14


@pytest.mark.parametrize("name,ev", [("als_kat.npz", O.poly_eval), ("als_c2small.npz", O.poly_eval),
                                     ("als_c3small.npz", O.poly_eval), ("als_legendre.npz", O.legendre_eval)])
def test_als_against_reference(golden, name, ev):
    """Normal equations and solutions of the first micro-steps at kernel-level tolerance; fitted values after the
    fixture's sweeps within the reference's own noise floor for that sweep count."""
    g = golden(name)
    x, y, p = g["x"], g["y"], int(g["p"])
    phis = [ev(x[:, i], p) for i in range(x.shape[1])]
    d = x.shape[1]
    # teacher-forced first left half-sweep: each A_k, b_k from the reference's own state (its solutions x_0..x_{k-1})
    cores = cores_from(g, "init_")
    O.orthonormalize(cores, "right", start=1)
    Rs = O.right_interfaces(cores, phis)
    L = np.ones((x.shape[0], 1))
    for k in range(min(int(g["n_solves_kept"]), d - 1)):
        V = O.design_matrix(L, phis[k], Rs[k])
        assert rel(V.T @ V, g[f"A_{k}"]) < 1e-12, k
        assert rel(V.T @ y, g[f"b_{k}"]) < 1e-12, k
        sol = O.solver(V.T @ V, V.T @ y, "lstsq")
        assert rel(V @ sol, V @ g[f"x_{k}"]) < 1e-7, k          # same fitted values from the solve (gauge-free)
        Q, S = np.linalg.qr(O.left_unfold(g[f"x_{k}"].reshape(cores[k].shape)))
        cores[k + 1] = (S @ O.right_unfold(cores[k + 1])).reshape(cores[k + 1].shape)
        cores[k] = Q.reshape(cores[k].shape)
        L = np.einsum("sa,sab->sb", L, O.core_times_phi(cores[k], phis[k]))
    cores = O.als(phis, y, n_iter=int(g["n_iter"]), cores=cores_from(g, "init_"), randomize=False)
    fitted = O.tt_eval(cores, phis)
    tol = {"als_kat.npz": 1e-12, "als_c2small.npz": 1e-8, "als_c3small.npz": 1e-9, "als_legendre.npz": 1e-13}[name]
    assert rel(fitted, g["fitted"]) < tol
    assert [c.shape for c in cores] == [c.shape for c in cores_from(g, "final_")]


def test_als_known_answer(golden):
    """||x||^2 is exactly representable with p = 3, r = 2 (reference: tests/optimisation/test.py:35-64)."""
    g = golden("als_kat.npz")
    x, y = g["x"], g["y"]
    phis = [O.poly_eval(x[:, i], 3) for i in range(6)]
    cores = O.als(phis, y, n_iter=10, cores=cores_from(g, "init_"), randomize=False)
    assert rel(O.tt_eval(cores, phis), y) < 1e-6


@pytest.mark.parametrize("tag", ["a", "b"])
def test_salsa_against_reference(golden, tag):
    """SALSA: teacher-forced normal equations of the first sweeps (where the reference is reproducible across
    contraction orders) and the rank bookkeeping of the whole run; final fitted values at residual level."""
    g = golden(f"salsa_{tag}.npz")
    x, y, p, r = g["x"], g["y"], int(g["p"]), int(g["r"])
    d = x.shape[1]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    trace = []
    ranks = [1] + [r] * (d - 1) + [1]
    cores = O.salsa(phis, y, n_iter=int(g["n_iter"]), ranks=ranks, cores=cores_from(g, "init_"), max_rank=int(g["max_rank"]),
                    randomize=False, init_ranks=ranks, trace=trace)
    solves = [t for t in trace if "A" in t]
    assert len(solves) == int(g["n_solves_kept"])
    for k in range(2 * d):                                        # sweeps 0 and 1
        assert solves[k]["A"].shape == g[f"A_{k}"].shape
        assert rel(solves[k]["A"], g[f"A_{k}"]) < 1e-8, k
        assert rel(solves[k]["b"], g[f"b_{k}"]) < 1e-8, k
    for k, t in enumerate(solves):                                # rank bookkeeping of every micro-step: exact
        assert t["A"].shape == g[f"A_{k}"].shape, k
    assert [c.shape for c in cores] == [c.shape for c in cores_from(g, "final_")]
    fitted = O.tt_eval(cores, phis)
    res_ref = rel(g["fitted"], y)
    assert abs(rel(fitted, y) - res_ref) < 0.5 * res_ref + 1e-9


def test_paths_and_models(golden):
    g = golden("paths.npz")
    bs, hjb = O.BlackScholes(1, 0.05, 0.4), O.HJB(1, np.sqrt(2))
    x, _ = O.generate_trajectories(g["bs_x"][:, 0], 10, bs, xi=g["bs_xi"])
    np.testing.assert_array_equal(x, g["bs_x"])
    x, _ = O.generate_trajectories(g["hjb_x"][:, 0], 10, hjb, xi=g["hjb_xi"])
    np.testing.assert_array_equal(x, g["hjb_x"])
    np.testing.assert_allclose(hjb.h(g["hjb_x"][:, 3], 0.3, g["yv"], g["z"]), g["hjb_h"], rtol=1e-15)
    np.testing.assert_allclose(hjb.g(g["hjb_x"][:, -1]), g["hjb_g"], rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(bs.h(g["bs_x"][:, 3], 0.3, g["yv"], g["z"]), g["bs_h"], rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(bs.g(g["bs_x"][:, -1]), g["bs_g"], rtol=1e-15)


def test_explicit_scheme_small(golden):
    """Explicit ALS scheme on HJB, 5 steps (loop of src/scripts/pipeline.py:84-141), replaying the reference's
    increments and initial cores."""
    g = golden("explicit_hjb_small.npz")
    X, p, r, n_iter = g["X"], int(g["p"]), int(g["r"]), int(g["n_iter"])
    N, T1, d = X.shape
    steps = T1 - 1
    dt = 1 / steps
    model = O.HJB(1, np.sqrt(2))
    Y = np.zeros((N, steps + 1))
    Y[:, -1] = model.g(X[:, -1])
    fit = 0

    def phis_at(n, f):
        return [f(X[:, n, i], p) for i in range(d)]

    cores = O.als(phis_at(steps, O.poly_eval), Y[:, -1], n_iter=n_iter, cores=cores_from(g, f"init{fit}_"), randomize=False)
    for n in range(steps - 1, -1, -1):
        fit += 1
        Z = O.tt_grad(cores, phis_at(n + 1, O.poly_eval), phis_at(n + 1, O.poly_grad))
        target = model.h(X[:, n + 1], (n + 1) * dt, Y[:, n + 1], Z) * dt + Y[:, n + 1]
        cores = O.als(phis_at(n, O.poly_eval), target, n_iter=n_iter, cores=cores_from(g, f"init{fit}_"), randomize=False)
        Y[:, n] = O.tt_eval(cores, phis_at(n, O.poly_eval))
    assert rel(Z, g["Z_last"]) < 1e-6
    assert rel(Y, g["Y"]) < 1e-7
    assert abs(Y[:, 0].mean() - g["Y"][:, 0].mean()) < 1e-9 * abs(g["Y"][:, 0].mean())


@pytest.mark.parametrize("seed", [0, 1])
def test_pipeline_c1_Y0(golden, seed):
    """BASELINE config 1 (explicit ALS pipeline defaults): the oracle consumes the numpy stream in the reference's
    order (paths, then d cores per fit) and reproduces its Y0 within the reference's own run-to-run spread."""
    g = golden("pipeline_c1.npz")
    np.random.seed(seed)
    N, d, steps, p, r = 2000, 4, 50, 3, 2
    X0 = np.broadcast_to(np.array([1.0, 0.5, 1.0, 0.5]), (N, d))
    model = O.BlackScholes(1, 0.05, 0.4)
    X, _ = O.generate_trajectories(X0, steps, model)
    np.testing.assert_array_equal(X[:8, -1], g[f"X_last_seed{seed}"])
    Y, _ = O.explicit_scheme(X, model, p=p, ranks=(1, r, r, r, 1), n_iter=10)
    ref = float(g[f"Y0_seed{seed}"])
    assert abs(Y[:, 0].mean() - ref) < 5e-9 * abs(ref)
    assert rel(Y.mean(axis=0), g[f"Ymean_seed{seed}"]) < 1e-7


def test_philox_stream_properties():
    xi = O.philox_increments(99, 4096, 3, 6)
    assert not xi[0].any()
    z = xi[1:].ravel()
    assert abs(z.mean()) < 0.03 and abs(z.var() - 1) < 0.03
    # counter-based: a shard regenerates its slice of the global stream
    np.testing.assert_array_equal(O.philox_increments(99, 1024, 3, 6, sample_offset=3072), xi[:, :, 3072:])
    # known answer for Philox4x32-10 (Random123 kat_vectors: counter = key = 0)
    out = O._philox4x32_10(np.zeros((1, 4), dtype=np.uint64), 0, 0)[0]
    assert [int(v) for v in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_hessian_against_reference(golden):
    """hessian() of the reference (core/calculus/hessian.py:6-26), batch and single point, monomials and Legendre."""
    g = golden("hessian.npz")
    x = g["x"]
    d = x.shape[1]
    for name, ev, gr, dg in (("poly", O.poly_eval, O.poly_grad, O.poly_dgrad), ("leg", O.legendre_eval, O.legendre_grad, O.legendre_dgrad)):
        p = int(g[f"p_{name}"])
        cores = cores_from(g, f"cores_{name}_")
        H = O.tt_hessian(cores, [ev(x[:, i], p) for i in range(d)], [gr(x[:, i], p) for i in range(d)], [dg(x[:, i], p) for i in range(d)])
        assert rel(H, g[f"H_{name}"]) < 1e-13
        assert rel(H[:, :, 3], g[f"H1_{name}"]) < 1e-13
