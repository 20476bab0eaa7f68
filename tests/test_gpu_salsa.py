"""GPU parity of SALSA (bsde_fit_salsa) against the numpy oracle (a literal restatement of als.py:99-323 that the CPU
suite pins to the reference fixtures) and against the fixtures themselves.

SALSA is numerically chaotic over many sweeps (the reference itself ends with different ranks under two valid
contraction orders), so parity is asserted where the reference is reproducible: sweep by sweep over the
first sweeps from identical states (fitted values, eta / omega / min_sv, bond ranks), the rank bookkeeping of the
fixture runs, and residual-level agreement at the end."""
import numpy as np
import pytest

import oracle.tt_oracle as O
from util import cores_from, flat, ranks_of, rel, unflat

pytestmark = pytest.mark.gpu


def _device_salsa(backend, x, y, p, init, n_iter, max_rank, init_ranks, do_reg=True, omega=1.0, freq_omega=1.05, min_sv=0.2):
    from bsde_solver._backend import DeviceInputs
    from bsde_solver._lib import BASIS_POLY

    d = x.shape[1]
    inp = DeviceInputs(BASIS_POLY, p, backend.to_device(np.ascontiguousarray(x.T)))
    ranks = np.array(ranks_of(init), dtype=np.int32)
    cap = max(max_rank, int(ranks.max()))
    buf = np.zeros(d * cap * p * cap)
    f = flat(init)
    buf[: f.size] = f
    state = backend.fit_salsa(inp, backend.to_device(y), n_iter, ranks, init_ranks, omega, freq_omega, min_sv, max_rank, 0, do_reg, buf)
    cores = unflat(buf, ranks, p)
    fit = backend.tt_eval(inp, ranks, flat(cores)).cpu().numpy()
    return cores, ranks, fit, state


@pytest.mark.parametrize("n_iter", [1, 2, 3, 4])
def test_salsa_sweep_by_sweep_matches_oracle(backend, golden, n_iter):
    g = golden("salsa_a.npz")
    x, y, p, r = g["x"], g["y"], int(g["p"]), int(g["r"])
    d = x.shape[1]
    init = cores_from(g, "init_")
    ranks0 = [1] + [r] * (d - 1) + [1]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    trace = []
    ref = O.salsa(phis, y, n_iter=n_iter, ranks=ranks0, cores=[c.copy() for c in init], max_rank=5, randomize=False,
                  init_ranks=ranks0, trace=trace)
    ref_fit = O.tt_eval(ref, phis)
    cores, ranks, fit, state = _device_salsa(backend, x, y, p, init, n_iter, 5, ranks0)
    end = [t for t in trace if t.get("sweep_end")][-1]
    print("n_iter", n_iter, "fit diff", rel(fit, ref_fit), "ranks", ranks.tolist(), "state", state[:3], "oracle", end["eta"], end["omega"], end["min_sv"])
    assert ranks.tolist() == [c.shape[0] for c in ref] + [1]            # rank bookkeeping: exact
    if n_iter <= 2:
        # before the first rank adaptation (warm-up of 2 sweeps, als.py:213-225) the algorithm is well conditioned
        assert rel(fit, ref_fit) < 1e-9
        np.testing.assert_allclose(state[:3], [end["eta"], end["omega"], end["min_sv"]], rtol=1e-8, atol=1e-14)
    else:
        # From the first adaptation on, the new rank-one directions enter with singular value 0.01*min_sv and eta ~ 1e-4:
        # the reference amplifies rounding noise to ~1e-4..1e-3 within one sweep (flipping the signs LAPACK returns for
        # the singular vectors moves the oracle's own fitted values by 3.6e-4 here).  Residual level.
        res, res_ref = rel(fit, y), rel(ref_fit, y)
        assert rel(fit, ref_fit) < 5e-3
        assert abs(res - res_ref) < 0.25 * res_ref
        np.testing.assert_allclose(state[:3], [end["eta"], end["omega"], end["min_sv"]], rtol=0.5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_salsa_fixture_runs(backend, golden, tag):
    """The reference's own runs (4 and 7 sweeps, max_rank 5): same final bond ranks, same fit quality."""
    g = golden(f"salsa_{tag}.npz")
    x, y, p, r = g["x"], g["y"], int(g["p"]), int(g["r"])
    d = x.shape[1]
    ranks0 = [1] + [r] * (d - 1) + [1]
    cores, ranks, fit, state = _device_salsa(backend, x, y, p, cores_from(g, "init_"), int(g["n_iter"]), int(g["max_rank"]), ranks0)
    final = cores_from(g, "final_")
    res_ref, res = rel(g["fitted"], y), rel(fit, y)
    print(tag, "ranks", ranks.tolist(), "reference", ranks_of(final).tolist(), "residual", res, "reference", res_ref)
    assert ranks.tolist() == ranks_of(final).tolist()
    assert abs(res - res_ref) < 0.5 * res_ref + 1e-9
    assert state[3] == 0


def test_salsa_without_regulariser_and_python_entry_point(backend, golden):
    """do_reg=False (the step-0 call of the pipelines, pipeline.py:135-136) through the reference-shaped SALSA()."""
    from bsde_solver.core.optimisation.als import SALSA
    from bsde_solver.core.tensor.tensor_core import TensorCore
    from bsde_solver.core.tensor.tensor_train import TensorTrain
    from bsde_solver.utils import fast_contract

    g = golden("salsa_a.npz")
    x, y, p, r = g["x"], g["y"], int(g["p"]), int(g["r"])
    d = x.shape[1]
    init = cores_from(g, "init_")
    ranks0 = (1,) + (r,) * (d - 1) + (1,)
    phis_np = [O.poly_eval(x[:, i], p) for i in range(d)]
    ref = O.salsa(phis_np, y, n_iter=3, ranks=list(ranks0), cores=[c.copy() for c in init], max_rank=5, randomize=False,
                  init_ranks=list(ranks0), do_reg=False)
    ref_fit = O.tt_eval(ref, phis_np)
    phis = [TensorCore(ph, name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i, ph in enumerate(phis_np)]
    tt = TensorTrain([p] * d, ranks0)

    def replay(self):
        for k, name in enumerate(self.cores):
            self.cores[name][:] = init[k]
        return self

    orig = TensorTrain.randomize
    TensorTrain.randomize = replay
    try:
        out = SALSA(phis, y, n_iter=3, ranks=ranks0, init_tt=tt, max_rank=5, do_reg=False)
    finally:
        TensorTrain.randomize = orig
    assert out is tt
    fit = np.asarray(fast_contract(out, phis)).squeeze()
    assert [c.shape for c in out.cores.values()] == [c.shape for c in ref]
    assert rel(fit, ref_fit) < 5e-3                    # third sweep = first rank adaptation: residual level (see above)
    assert abs(rel(fit, y) - rel(ref_fit, y)) < 0.25 * rel(ref_fit, y)
    # two sweeps (no adaptation yet): exact
    ref2 = O.salsa(phis_np, y, n_iter=2, ranks=list(ranks0), cores=[c.copy() for c in init], max_rank=5, randomize=False,
                   init_ranks=list(ranks0), do_reg=False)
    tt2 = TensorTrain([p] * d, ranks0)
    TensorTrain.randomize = replay
    try:
        out2 = SALSA(phis, y, n_iter=2, ranks=ranks0, init_tt=tt2, max_rank=5, do_reg=False)
    finally:
        TensorTrain.randomize = orig
    assert rel(np.asarray(fast_contract(out2, phis)).squeeze(), O.tt_eval(ref2, phis_np)) < 1e-9
    assert tuple(out.ranks) == ranks0           # the constructor ranks stay stale, as in the reference (als.py:122-123)


def test_salsa_rank_growth_on_fresh_problem(backend):
    """A rank-3 target fitted from rank 1: bonds must grow, one per warm-up period, never beyond the caps."""
    rng = np.random.RandomState(4)
    N, d, p = 4000, 5, 3
    x = rng.rand(N, d) * 2 - 1
    y = np.sin(x.sum(axis=1)) + x[:, 0] * x[:, 1] * x[:, 2]
    ranks0 = [1] * (d + 1)
    init = [rng.rand(1, p, 1) * 2 - 1 for _ in range(d)]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ref = O.salsa(phis, y, n_iter=6, ranks=ranks0, cores=[c.copy() for c in init], max_rank=3, randomize=False, init_ranks=ranks0)
    cores, ranks, fit, state = _device_salsa(backend, x, y, p, init, 6, 3, ranks0)
    print("grown ranks", ranks.tolist(), "oracle", [c.shape[0] for c in ref] + [1])
    assert ranks.tolist() == [c.shape[0] for c in ref] + [1]
    assert max(ranks) > 1 and max(ranks) <= 3
    assert abs(rel(fit, y) - rel(O.tt_eval(ref, phis), y)) < 0.2 * rel(fit, y) + 1e-9


def test_salsa_step0_identical_samples(backend):
    """Time step 0 of the pipelines: every sample sits at X_0, the design matrix has rank one, the cores handed to the
    SVD moves are rank deficient (singular values exactly 0).  The fit must stay finite and reproduce the oracle's
    (ridge-shrunk) constant; SALSA is called with do_reg=False there (pipeline.py:135-136)."""
    rng = np.random.RandomState(12)
    N, d, p, r = 2000, 4, 3, 3
    x = np.tile(np.array([0.0, 0.0, 0.0, 0.0]), (N, 1))
    y = rng.randn(N) * 0.3 - 0.55
    ranks0 = [1] + [r] * (d - 1) + [1]
    init = [rng.rand(ranks0[i], p, ranks0[i + 1]) * 2 - 1 for i in range(d)]
    phis = [O.poly_eval(x[:, i], p) for i in range(d)]
    ref = O.salsa(phis, y, n_iter=4, ranks=ranks0, cores=[c.copy() for c in init], max_rank=3, randomize=False,
                  init_ranks=ranks0, do_reg=False)
    ref_fit = O.tt_eval(ref, phis)
    cores, ranks, fit, state = _device_salsa(backend, x, y, p, init, 4, 3, ranks0, do_reg=False)
    print("step-0 fit", fit[:2], "oracle", ref_fit[:2], "mean(y)", y.mean(), "state", state[:3])
    assert np.isfinite(fit).all() and np.ptp(fit) < 1e-12
    assert abs(fit[0] - ref_fit[0]) < 1e-6 * abs(ref_fit[0])
