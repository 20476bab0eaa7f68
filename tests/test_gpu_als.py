"""GPU parity of the ALS fit (bsde_fit_als through the reference-shaped Python entry point) against the golden
fixtures recorded from the unmodified reference and against the numpy oracle on fresh seeded inputs.

What is asserted follows SURVEY.md §8(c): fitted values (gauge invariant) at <= 1e-9 relative after the fixture's
sweeps; the first normal equations and solves of the run at kernel-level tolerance; cores after the sweeps at a
condition-scaled tolerance (the reference itself moves by 7.5e-8 between two valid contraction orders)."""
import numpy as np
import pytest

from util import cores_from, flat, ranks_of, rel, unflat

pytestmark = pytest.mark.gpu


def _fit_from_golden(g, basis_cls, n_iter=None):
    from bsde_solver.core.optimisation.als import ALS, last_fit_info
    from bsde_solver.core.tensor.tensor_core import TensorCore
    from bsde_solver.core.tensor.tensor_train import TensorTrain
    from bsde_solver.utils import fast_contract

    x, y, p = g["x"], g["y"], int(g["p"])
    init = cores_from(g, "init_")
    d = x.shape[1]
    basis = basis_cls(p)
    phis = [TensorCore(basis.eval(x[:, i]), name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i in range(d)]
    tt = TensorTrain([p] * d, list(ranks_of(init)))

    # feed the reference's own initial cores: ALS re-randomises init_tt (als.py:37), so replay them through randomize
    def replay(self):
        for k, name in enumerate(self.cores):
            self.cores[name][:] = init[k]
        return self

    orig = TensorTrain.randomize
    TensorTrain.randomize = replay
    try:
        out = ALS(phis, y, n_iter=int(g["n_iter"]) if n_iter is None else n_iter, ranks=tuple(ranks_of(init)), init_tt=tt)
    finally:
        TensorTrain.randomize = orig
    assert out is tt                                   # mutated in place and returned (als.py:36-37,96)
    fitted = np.asarray(fast_contract(out, phis)).squeeze()
    return out, fitted, dict(last_fit_info)


@pytest.mark.parametrize("name,basis", [("als_kat.npz", "poly"), ("als_c2small.npz", "poly"), ("als_c3small.npz", "poly"),
                                        ("als_legendre.npz", "leg")])
def test_als_fitted_values_match_reference(backend, golden, name, basis):
    from bsde_solver.core.calculus.basis import LegendreBasis, PolynomialBasis

    g = golden(name)
    tt, fitted, info = _fit_from_golden(g, PolynomialBasis if basis == "poly" else LegendreBasis)
    ref = g["fitted"]
    err = rel(fitted, ref)
    res_ref = rel(ref, g["y"])
    # floor of the comparison: the numpy oracle on the same inputs against the same reference output (the fixtures
    # stop after 3-4 sweeps, where two valid summation orders still differ by ~1e-9: SURVEY.md App. B)
    import oracle.tt_oracle as O

    ev = O.poly_eval if basis == "poly" else O.legendre_eval
    phis = [ev(g["x"][:, i], int(g["p"])) for i in range(g["x"].shape[1])]
    oc = O.als(phis, g["y"], n_iter=int(g["n_iter"]), cores=cores_from(g, "init_"), randomize=False)
    floor = rel(O.tt_eval(oc, phis), ref)
    print(name, "fitted rel diff", err, "oracle-vs-reference floor", floor, "reference residual", res_ref, info)
    assert err < max(1e-9, 4 * floor)
    # residual of the fit itself is the reference's
    assert abs(rel(fitted, g["y"]) - res_ref) < max(1e-9, 4 * floor)
    final = cores_from(g, "final_")
    assert [c.shape for c in tt.cores.values()] == [c.shape for c in final]      # rank bookkeeping: exact


def test_als_kat_known_answer(backend, golden):
    """tests/optimisation/test.py:35-64 of the reference: ||x||^2 (d=6, p=3, r=2) is exactly representable."""
    from bsde_solver.core.calculus.basis import PolynomialBasis

    g = golden("als_kat.npz")
    _, fitted, _ = _fit_from_golden(g, PolynomialBasis)
    # the reference's own run ends at a relative residual of 6.5e-8 (A^T A of an exact fit is singular to rounding)
    assert rel(fitted, g["y"]) < 1e-6
    assert rel(g["fitted"], g["y"]) < 1e-6


@pytest.mark.parametrize("name", ["als_kat.npz", "als_c2small.npz", "als_c3small.npz"])
def test_first_micro_steps_match_reference(backend, golden, name):
    """Teacher-forced: the reference's A, b of micro-step k are reproduced from the reference's own state.
    The first left half-sweep only depends on the initial cores and the solutions x_0..x_{k-1}, which we replay
    through the oracle's QR so that each A_k is compared on identical inputs."""
    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    g = golden(name)
    x, y, p = g["x"], g["y"], int(g["p"])
    d = x.shape[1]
    cores = [c.copy() for c in cores_from(g, "init_")]
    O.orthonormalize(cores, "right", start=1)
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    yd = backend.to_device(y)
    kmax = min(int(g["n_solves_kept"]), d - 1)
    for k in range(kmax):
        A, b = backend.normal_eq(inp, yd, ranks_of(cores), flat(cores), k)
        assert rel(A, g[f"A_{k}"]) < 1e-12, k
        assert rel(b, g[f"b_{k}"]) < 1e-12, k
        # advance with the reference's own solution (als.py:65-77)
        X = g[f"x_{k}"].reshape(cores[k].shape)
        Q, S = np.linalg.qr(O.left_unfold(X))
        cores[k + 1] = (S @ O.right_unfold(cores[k + 1])).reshape(cores[k + 1].shape)
        cores[k] = Q.reshape(cores[k].shape)


def test_als_matches_oracle_on_fresh_inputs(backend):
    """Seeded inputs the fixtures do not cover: ragged ranks, p=5, Legendre, N not a multiple of 32."""
    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis

    rng = np.random.RandomState(5)
    for d, p, ranks, N, kind in ((5, 3, [1, 2, 3, 3, 2, 1], 1531, "poly"), (4, 5, [1, 3, 3, 3, 1], 2048, "poly"),
                                 (6, 4, [1, 4, 4, 4, 4, 4, 1], 5000, "leg"), (2, 3, [1, 2, 1], 257, "poly"),
                                 (8, 4, [1, 4, 4, 4, 4, 4, 4, 4, 1], 8192, "poly")):
        x = rng.rand(N, d) * 2 - 1
        y = np.cos(x.sum(axis=1)) + 0.3 * x[:, 0] * x[:, -1]
        init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
        ev = O.poly_eval if kind == "poly" else O.legendre_eval
        phis = [ev(x[:, i], p) for i in range(d)]
        ref = O.als(phis, y, n_iter=6, cores=[c.copy() for c in init], randomize=False)
        ref_fit = O.tt_eval(ref, phis)
        inp = DeviceInputs(BASIS_POLY if kind == "poly" else BASIS_LEGENDRE, p, backend.to_device(np.ascontiguousarray(x.T)),
                           coef=None if kind == "poly" else LegendreBasis(p)._coef())
        buf = flat(init)
        for use_graph in (1, 0):
            backend.set_option("use_graph", use_graph)
            got = buf.copy()
            info = backend.fit_als(inp, backend.to_device(y), 6, np.array(ranks, dtype=np.int32), 0, got)
            fit = backend.tt_eval(inp, ranks, got).cpu().numpy()
            assert rel(fit, ref_fit) < 1e-9, (d, p, kind, use_graph, info)
        backend.set_option("use_graph", 1)


def test_fused_interface_update_is_bit_identical(backend):
    """ALS forms the interface a micro-step leaves behind inside the NEXT assembly kernel (FusedInterface, assembly.cu).
    It must give exactly the bits of the separate interface kernels: same cores after the fit, eager and graph."""
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis

    rng = np.random.RandomState(11)
    for d, p, ranks, N, kind in ((2, 3, [1, 2, 1], 257, "poly"), (3, 4, [1, 4, 3, 1], 1000, "poly"),
                                 (6, 4, [1, 4, 4, 4, 4, 4, 1], 40000, "poly"), (5, 3, [1, 2, 3, 3, 2, 1], 1531, "leg"),
                                 (4, 5, [1, 3, 5, 3, 1], 2048, "poly"), (12, 4, [1] + [4] * 11 + [1], 70001, "poly")):
        x = rng.randn(N, d) * 0.7
        y = np.log(0.5 + 0.5 * (x**2).sum(axis=1)) + 0.1 * x[:, 0]
        init = flat([rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)])
        inp = DeviceInputs(BASIS_POLY if kind == "poly" else BASIS_LEGENDRE, p, backend.to_device(np.ascontiguousarray(x.T)),
                           coef=None if kind == "poly" else LegendreBasis(p)._coef())
        yd = backend.to_device(y)
        out = {}
        try:
            for fuse in (0, 1):
                for use_graph in (1, 0):
                    backend.set_option("fuse_interface", fuse)
                    backend.set_option("use_graph", use_graph)
                    got = init.copy()
                    n0 = backend.launch_count()
                    backend.fit_als(inp, yd, 4, np.array(ranks, dtype=np.int32), 0, got)
                    out[(fuse, use_graph)] = (got, backend.launch_count() - n0)
        finally:
            backend.set_option("fuse_interface", 1)
            backend.set_option("use_graph", 1)
        base = out[(0, 0)][0]
        assert np.all(np.isfinite(base))
        for key, (got, launches) in out.items():
            assert np.array_equal(got, base), (d, p, kind, key)
        # one launch fewer per micro-step (first sweep: the first micro-step has nothing to fuse)
        assert out[(1, 1)][1] < out[(0, 1)][1]


def test_programmatic_dependent_launch_is_bit_identical(backend):
    """Option "pdl" (DESIGN.md section 4a): inside the captured sweeps every kernel -> kernel edge becomes programmatic and the
    solve kernel reads its expansion map ahead of griddepcontrol.wait.  Nothing a kernel consumes may change: same cores to
    the last bit as with plain launches and as with eager launches, for fast-path shapes, general shapes (k_micro_solve,
    generic assembly), Legendre inputs and a fit whose solves truncate."""
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis

    rng = np.random.RandomState(77)
    for d, p, ranks, N, kind, target in ((6, 4, [1, 4, 4, 4, 4, 4, 1], 40000, "poly", "log"), (4, 5, [1, 3, 5, 3, 1], 2048, "poly", "log"),
                                         (5, 3, [1, 2, 3, 3, 2, 1], 1531, "leg", "log"), (12, 4, [1] + [4] * 11 + [1], 70001, "poly", "log"),
                                         (6, 4, [1, 4, 4, 4, 4, 4, 1], 1 << 13, "poly", "square"), (3, 6, [1, 5, 5, 1], 3000, "poly", "log")):
        if target == "square":      # the converging fit of test_truncated_solves_inside_a_converging_fit: its solves truncate
            x = np.tile([1.0, 0.5], d // 2) * np.exp(-0.03 + 0.4 * rng.randn(N, d))
            y = (x**2).sum(axis=1)
        else:
            x = rng.randn(N, d) * 0.7
            y = np.log(0.5 + 0.5 * (x**2).sum(axis=1)) + 0.1 * x[:, 0]
        init = flat([rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)])
        inp = DeviceInputs(BASIS_POLY if kind == "poly" else BASIS_LEGENDRE, p, backend.to_device(np.ascontiguousarray(x.T)),
                           coef=None if kind == "poly" else LegendreBasis(p)._coef())
        yd = backend.to_device(y)
        out, edges = {}, {}
        try:
            for key, (pdl, use_graph) in {"pdl": (1, 1), "plain": (0, 1), "eager": (0, 0)}.items():
                backend.set_option("pdl", pdl)
                backend.set_option("use_graph", use_graph)
                got = init.copy()
                info = backend.fit_als(inp, yd, 6, np.array(ranks, dtype=np.int32), 0, got)
                out[key] = (got, [int(v) for v in info[:3]])
                edges[key] = (int(backend.get_stat("graph_edges")), int(backend.get_stat("graph_pdl_edges")))
        finally:
            backend.set_option("pdl", 1)
            backend.set_option("use_graph", 1)
        assert np.all(np.isfinite(out["eager"][0]))
        for key in ("pdl", "plain"):
            assert np.array_equal(out[key][0], out["eager"][0]), (d, p, kind, target, key)
            assert out[key][1] == out["eager"][1], (d, p, kind, target, key)
        if target == "square":
            assert out["pdl"][1][1] > 10, out["pdl"][1]      # this fit converges: many of its solves truncate
        # with the option the kernel -> kernel edges of the captured sweep are programmatic, without it none is
        assert edges["pdl"][1] > 0 and edges["plain"] == (edges["pdl"][0], 0), edges
        if p == 4 and max(ranks) == 4:                       # a chain of 3 kernels per micro-step, every edge programmatic
            assert edges["pdl"] == (3 * 2 * (d - 1) - 1, 3 * 2 * (d - 1) - 1), edges


def test_fast_solve_kernel_matches_the_general_one(backend):
    """k_micro_solve_fast (register-resident Cholesky, n <= 64) against k_micro_solve alone on the same fits: even and odd
    numbers of unknowns, ragged ranks, and agreement with the numpy oracle."""
    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(21)
    for d, p, ranks, N in ((4, 3, [1, 3, 3, 3, 1], 3000), (6, 4, [1, 4, 4, 4, 4, 4, 1], 20000), (5, 3, [1, 2, 3, 3, 2, 1], 1531),
                           (3, 5, [1, 3, 3, 1], 2500), (4, 4, [1, 4, 4, 2, 1], 4097)):
        x = rng.rand(N, d) * 2 - 1
        y = np.cos(x.sum(axis=1)) + 0.3 * x[:, 0] * x[:, -1]
        init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
        phis = [O.poly_eval(x[:, i], p) for i in range(d)]
        ref = O.tt_eval(O.als(phis, y, n_iter=5, cores=[c.copy() for c in init], randomize=False), phis)
        inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
        yd = backend.to_device(y)
        fits, infos = {}, {}
        try:
            for fast in (0, 1):
                backend.set_option("fast_solve", fast)
                got = flat(init)
                infos[fast] = backend.fit_als(inp, yd, 5, np.array(ranks, dtype=np.int32), 0, got)
                fits[fast] = backend.tt_eval(inp, ranks, got).cpu().numpy()
        finally:
            backend.set_option("fast_solve", 1)
        assert list(infos[0]) == list(infos[1]), (ranks, infos)          # same solver path taken by every micro-step
        assert infos[1][0] > 0
        assert rel(fits[1], fits[0]) < 1e-10, (ranks, rel(fits[1], fits[0]))
        assert rel(fits[1], ref) < 1e-9, (ranks, rel(fits[1], ref))


def test_truncated_solves_inside_a_converging_fit(backend):
    """A train with more rank than the target needs (y = |x|^2 has TT rank 2, the train rank 4): once the fit converges most
    normal equations are numerically rank-deficient and numpy.linalg.lstsq truncates.  k_micro_solve_fast then runs
    vh_lstsq (right-looking pivoted Cholesky + odd-even Jacobi) and, with the per-core hints, skips the factorisation the
    next time the core comes up.  Asserted: the fit reaches the target as the numpy oracle does; hints on / off and the
    general kernel (left-looking factorisation, same Jacobi) give the same fitted values to rounding amplified by the
    truncation; truncated solves were actually taken; repeated runs give the same bits."""
    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(51)
    d, p, N, sweeps = 6, 4, 1 << 13, 6
    ranks = np.array([1] + [4] * (d - 1) + [1], dtype=np.int32)
    x = np.tile([1.0, 0.5], d // 2) * np.exp(-0.03 + 0.4 * rng.randn(N, d))
    y = (x**2).sum(axis=1)
    init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ref = O.tt_eval(O.als(phis, y, n_iter=sweeps, cores=[c.copy() for c in init], randomize=False), phis)
    ref_res = rel(ref, y)
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    yd = backend.to_device(y)
    fits, infos, cores = {}, {}, {}
    try:
        for key, (fast, hints) in {"hints": (1, 1), "no_hints": (1, 0), "general": (0, 1), "again": (1, 1)}.items():
            backend.set_option("fast_solve", fast)
            backend.set_option("solve_hints", hints)
            got = flat(init)
            infos[key] = backend.fit_als(inp, yd, sweeps, ranks, 0, got)
            cores[key] = got
            fits[key] = backend.tt_eval(inp, ranks, got).cpu().numpy()
    finally:
        backend.set_option("fast_solve", 1)
        backend.set_option("solve_hints", 1)
    assert infos["hints"][1] > 10 and infos["no_hints"][1] > 10, infos         # truncated solves taken
    assert infos["hints"][1] >= infos["no_hints"][1], infos                    # hinted cores skip the factorisation
    assert np.array_equal(cores["hints"], cores["again"])
    res = rel(fits["hints"], y)
    assert res < max(10 * ref_res, 1e-6), (res, ref_res)
    for key in ("no_hints", "general"):
        assert rel(fits[key], fits["hints"]) < max(1e-7, 10 * res), (key, rel(fits[key], fits["hints"]), res)
    assert rel(fits["hints"], ref) < max(1e-7, 10 * max(res, ref_res)), (rel(fits["hints"], ref), res, ref_res)


def test_step0_fit_returns_mean(backend):
    """at time step 0 all samples coincide, A has rank 1, lstsq's minimum-norm solution fits mean(y)."""
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(9)
    d, p, r, N = 4, 3, 2, 2000
    x = np.tile(np.array([1.0, 0.5, 1.0, 0.5]), (N, 1))
    y = rng.randn(N) + 3.0
    ranks = np.array([1, r, r, r, 1], dtype=np.int32)
    init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    got = flat(init)
    info = backend.fit_als(inp, backend.to_device(y), 3, ranks, 0, got)
    fit = backend.tt_eval(inp, ranks, got).cpu().numpy()
    assert info[1] > 0 and info[3] == 1                 # every solve is rank 1 and goes through the SVD path
    np.testing.assert_allclose(fit, y.mean(), rtol=1e-11)


def test_als_errors(backend):
    from bsde_solver._lib import BackendError
    from bsde_solver.core.optimisation.als import ALS

    x = np.random.RandomState(0).rand(50, 3)
    phis = [np.array([x[:, i] ** k for k in range(3)]).T for i in range(3)]
    with pytest.raises(ValueError):
        ALS(phis, np.ones(50), ranks=(1, 2, 2, 1), optimizer="nope")
    with pytest.raises(ValueError):
        ALS(phis, np.ones(49), ranks=(1, 2, 2, 1))
    with pytest.raises(BackendError):
        ALS(phis, np.ones(50), ranks=(2, 2, 2, 1))       # boundary rank must be 1


def test_specialised_rank4_assembly_is_bit_identical_to_the_general_form(backend):
    """k_assemble_r4<FULL> (interior launches: rank-4 neighbour, N a multiple of 32; component counts, tail masks and
    address arithmetic compiled out) must produce the bits of the general instantiation: same cores after the fit."""
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(31)
    for d, N in ((7, 1 << 15), (12, 96000)):
        p, ranks = 4, np.array([1] + [4] * (d - 1) + [1], dtype=np.int32)
        x = rng.randn(N, d) * 0.6
        y = np.log(0.5 + 0.5 * (x**2).sum(axis=1)) - 0.2 * x[:, 1] * x[:, 2]
        init = flat([rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)])
        inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
        yd = backend.to_device(y)
        out = {}
        try:
            for full in (1, 0):
                backend.set_option("r4_full", full)
                got = init.copy()
                backend.fit_als(inp, yd, 3, ranks, 0, got)
                out[full] = got
        finally:
            backend.set_option("r4_full", 1)
        assert np.all(np.isfinite(out[1]))
        assert np.array_equal(out[0], out[1]), (d, N)


def test_moment_reduction_row_groups_agree(backend):
    """The partial rows of the assembly kernel are summed in 1..6 row groups (k_moment_reduce) that the solve kernel adds
    while staging; a different grouping is a different (fixed) summation tree: same fit to rounding."""
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(32)
    d, p, N = 6, 4, 1 << 17
    ranks = np.array([1] + [4] * (d - 1) + [1], dtype=np.int32)
    x = rng.randn(N, d) * 0.5
    y = np.sin(x[:, 0]) + 0.3 * x[:, 1] * x[:, 2] + 0.1 * (x**2).sum(axis=1)
    init = flat([rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)])
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    yd = backend.to_device(y)
    fits = {}
    try:
        for groups in (1, 3, 6):
            backend.set_option("red_groups", groups)
            got = init.copy()
            backend.fit_als(inp, yd, 3, ranks, 0, got)
            fits[groups] = backend.tt_eval(inp, ranks, got).cpu().numpy()
    finally:
        backend.set_option("red_groups", 3)
    for groups in (3, 6):
        assert rel(fits[groups], fits[1]) < 1e-10, groups


@pytest.mark.parametrize("d,p,ranks,N,kind", [(4, 6, [1, 5, 5, 5, 1], 4096, "poly"),          # n = 150 (r = 5, p = 6)
                                              (5, 4, [1, 4, 8, 8, 4, 1], 6000, "poly"),       # n = 256
                                              (4, 8, [1, 8, 8, 8, 1], 8192, "leg")])          # n = 512, 47k moment slots
def test_cores_beyond_the_shared_memory_solver(backend, d, p, ranks, N, kind):
    """Cores whose system does not fit the one-CTA solver's shared memory (n = r p r' > ~140, up to 512 = every shape the
    assembly accepts) are solved in a global-memory workspace (dense.cu, MicroSolveArgs::big_ws): normal equations and
    fitted values against the oracle, eager and graph-captured sweeps."""
    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY
    from bsde_solver.core.calculus.basis import LegendreBasis

    rng = np.random.RandomState(11 * d + p)
    x = rng.rand(N, d) * 2 - 1
    y = np.cos(x.sum(axis=1)) + 0.3 * x[:, 0] * x[:, -1] + 0.1 * np.sin(3 * x[:, 1])
    init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    ev = O.poly_eval if kind == "poly" else O.legendre_eval
    phis = [ev(x[:, i], p) for i in range(d)]
    inp = DeviceInputs(BASIS_POLY if kind == "poly" else BASIS_LEGENDRE, p, backend.to_device(np.ascontiguousarray(x.T)),
                       coef=None if kind == "poly" else LegendreBasis(p)._coef())
    yd = backend.to_device(y)
    # the largest core's normal equations
    j = int(np.argmax([ranks[i] * ranks[i + 1] for i in range(d)]))
    Ls, Rs = O.left_interfaces(init, phis), O.right_interfaces(init, phis)
    V = O.design_matrix(Ls[j], phis[j], Rs[j])
    A, b = backend.normal_eq(inp, yd, ranks, flat(init), j)
    assert A.shape[0] == ranks[j] * p * ranks[j + 1] > 144
    assert rel(A, V.T @ V) < 1e-13 and rel(b, V.T @ y) < 1e-12
    sweeps = 3
    ref = O.als(phis, y, n_iter=sweeps, cores=[c.copy() for c in init], randomize=False)
    ref_fit = O.tt_eval(ref, phis)
    floor = np.linalg.norm(ref_fit - y) / np.linalg.norm(y)
    for use_graph in (1, 0):
        backend.set_option("use_graph", use_graph)
        got = flat(init).copy()
        info = backend.fit_als(inp, yd, sweeps, np.array(ranks, dtype=np.int32), 0, got)
        fit = backend.tt_eval(inp, ranks, got).cpu().numpy()
        print(f"n_max={A.shape[0]} graph={use_graph} solves={info.tolist()} fitted values vs oracle {rel(fit, ref_fit):.2e} residual {floor:.2e}")
        assert rel(fit, ref_fit) < 1e-8, (d, p, kind, use_graph, info)
    backend.set_option("use_graph", 1)


@pytest.mark.parametrize("d,p,ranks,N", [(4, 6, [1, 5, 5, 5, 1], 3000), (5, 4, [1, 4, 8, 8, 4, 1], 5000)])
def test_truncated_solves_of_large_cores(backend, d, p, ranks, N):
    """y = |x|^2 needs rank 2: the converged systems of the wider train are rank-deficient and take the truncated path
    (pivoted Cholesky + one-sided Jacobi, vh_lstsq.cuh) in the global workspace, 150 and 256 unknowns per core."""
    import time

    import oracle.tt_oracle as O
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    rng = np.random.RandomState(3 * d + p)
    x = rng.rand(N, d) * 1.5 - 0.5
    y = (x * x).sum(axis=1)
    init = [rng.rand(ranks[i], p, ranks[i + 1]) * 2 - 1 for i in range(d)]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    yd = backend.to_device(y)
    sweeps = 4
    ref_fit = O.tt_eval(O.als(phis, y, n_iter=sweeps, cores=[c.copy() for c in init], randomize=False), phis)
    floor = np.linalg.norm(ref_fit - y) / np.linalg.norm(y)
    got = flat(init).copy()
    t0 = time.perf_counter()
    info = backend.fit_als(inp, yd, sweeps, np.array(ranks, dtype=np.int32), 0, got)
    dt = time.perf_counter() - t0
    fit = backend.tt_eval(inp, ranks, got).cpu().numpy()
    res = np.linalg.norm(fit - y) / np.linalg.norm(y)
    print(f"ranks={ranks} p={p} solves={info.tolist()} fit {dt * 1e3:.1f} ms, fitted values vs oracle {rel(fit, ref_fit):.2e}, "
          f"residual {res:.2e} (oracle {floor:.2e})")
    assert info[1] > 0, info                        # truncated solves happened
    # (truncated fits are only reproducible to the level of their own residual: SURVEY.md §8c, the reference's chaos)
    assert res < 3 * floor + 1e-9 and rel(fit, ref_fit) < max(1e-9, floor), (info, res, floor, rel(fit, ref_fit))
