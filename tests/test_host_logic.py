"""CPU: host-side mirror of the reference interface — named-axis cores, the train container, the wire format of the
C ABI, argument checking.  (Everything numerical lives behind the library and is tested under -m gpu.)"""
import numpy as np
import pytest

import oracle.tt_oracle as O
from bsde_solver import reload_backend
from bsde_solver.core.tensor.tensor_core import TensorCore
from bsde_solver.core.tensor.tensor_network import TensorNetwork
from bsde_solver.core.tensor.tensor_train import TensorTrain, left_unfold, pack_cores, right_unfold, unpack_cores


def test_reload_backend_contract():
    reload_backend("numpy")
    reload_backend("cupy")          # accepted name; there is one CUDA backend behind both
    with pytest.raises(ValueError):
        reload_backend("jax")


def test_tensor_core_unfold_rename_like():
    a = np.arange(24.0).reshape(2, 3, 4)
    c = TensorCore(a, indices=("r_0", "m_1", "r_1"), name="core_0")
    lu, ru = left_unfold(c), right_unfold(c)
    assert lu.shape == (6, 4) and lu.indices == ("r_0+m_1", "r_1")
    assert ru.shape == (2, 12) and ru.indices == ("r_0", "m_1+r_1")
    np.testing.assert_array_equal(np.asarray(lu), a.reshape(6, 4))
    t = c.unfold("r_1", ("m_1", "r_0"))
    np.testing.assert_array_equal(np.asarray(t), a.transpose(2, 1, 0).reshape(4, 6))
    rest = c.unfold("m_1", -1)
    assert rest.shape == (3, 8) and rest.indices == ("m_1", "r_0+r_1")
    r = c.copy().rename("r_*", "s_*")
    assert r.indices == ("s_0", "m_1", "s_1") and c.indices == ("r_0", "m_1", "r_1")
    k = TensorCore.like(np.arange(24.0), c)
    assert k.shape == (2, 3, 4) and k.indices == c.indices and k.name == "core_0"
    assert c.size_at("m_1") == 3
    assert "r_0 {2}" in repr(c)


def test_tensor_train_container_and_wire_format():
    np.random.seed(3)
    tt = TensorTrain([3, 3, 3, 3], [1, 2, 3, 2, 1])
    assert tt.order == 4 and list(tt.cores) == ["core_0", "core_1", "core_2", "core_3"]
    assert tt.cores["core_1"].indices == ("r_1", "m_2", "r_2") and tt.cores["core_1"].shape == (2, 3, 3)
    np.random.seed(3)
    ref = O.randomize_cores([3, 3, 3, 3], [1, 2, 3, 2, 1])
    np.random.seed(3)
    assert tt.randomize() is tt
    for a, b in zip(tt.cores.values(), ref):          # same draws from the global numpy stream as the reference
        np.testing.assert_array_equal(np.asarray(a), b)
    ranks, flat = pack_cores(tt)
    assert ranks.tolist() == [1, 2, 3, 2, 1] and flat.size == 6 + 18 + 18 + 6
    np.testing.assert_array_equal(flat[6:24], ref[1].ravel())
    tt2 = TensorTrain([3, 3, 3, 3], [1, 2, 3, 2, 1])
    unpack_cores(tt2, ranks, flat)
    for a, b in zip(tt2.cores.values(), ref):
        np.testing.assert_array_equal(np.asarray(a), b)
    assert tt2.cores["core_2"].indices == ("r_2", "m_3", "r_3") and tt2.cores["core_2"].name == "core_2"
    # rank growth (SALSA): unpack with new ranks keeps names and labels
    grown = np.zeros(1 * 3 * 3 + 3 * 3 * 3 + 3 * 3 * 2 + 2 * 3 * 1)
    unpack_cores(tt2, [1, 3, 3, 2, 1], grown)
    assert tt2.cores["core_0"].shape == (1, 3, 3) and tt2.cores["core_1"].shape == (3, 3, 3)
    cp = tt.copy()
    cp.cores["core_0"][:] = 0
    assert np.asarray(tt.cores["core_0"]).any()


def test_from_tensor_reconstructs():
    rng = np.random.RandomState(0)
    full = np.einsum("ia,ajb,bk->ijk", rng.randn(3, 2), rng.randn(2, 4, 2), rng.randn(2, 3))
    tt = TensorTrain.from_tensor(full, [1, 2, 2, 1])
    cores = [np.asarray(c) for c in tt.cores.values()]
    back = np.einsum("xia,ajb,bky->ijk", *cores)
    np.testing.assert_allclose(back, full, atol=1e-12)


def test_tensor_network_contract_and_sample():
    rng = np.random.RandomState(1)
    core = TensorCore(rng.randn(2, 3, 2), indices=("r_0", "m_1", "r_1"))
    phi = TensorCore(rng.randn(5, 3), indices=("batch", "m_1"))
    out = TensorNetwork(cores=[core, phi], names=["core_0", "phi_0"]).contract(indices=("batch", "r_0", "r_1"))
    np.testing.assert_allclose(np.asarray(out), np.einsum("amb,sm->sab", np.asarray(core), np.asarray(phi)), atol=1e-14)
    assert out.indices == ("batch", "r_0", "r_1")
    tn = TensorNetwork(cores=[phi.copy()], names=["phi_0"])
    assert np.asarray(tn.sample(2)["phi_0"]).shape == (3,)
    with pytest.raises(ValueError):
        tn.add_core(phi.copy(), "phi_0")


def test_stack_phis_layout():
    from bsde_solver._inputs import stack_phis

    x = np.random.RandomState(2).randn(7, 3)
    phis = [O.poly_eval(x[:, i], 4) for i in range(3)]
    out = stack_phis(phis)
    assert out.shape == (3, 4, 7) and out.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(out[1, 2], x[:, 1] ** 2)
    with pytest.raises(ValueError):
        stack_phis([np.ones((7, 3)), np.ones((7, 4))])
    with pytest.raises(ValueError):
        stack_phis([np.ones((7, 3)), np.ones((6, 3))])


def test_legendre_table_is_scipy_polyval_order():
    from bsde_solver.core.calculus.basis import LegendreBasis

    b = LegendreBasis(5)
    asc = O.legendre_coefficients(5)
    np.testing.assert_array_equal(b._coef()[:, :, ::-1], asc)      # descending (device Horner) == ascending reversed


def test_solver_name_check_is_a_value_error():
    from bsde_solver.core.optimisation.solve import solver_id

    assert [solver_id(m) for m in ("lstsq", "lu", "qr", "solve")] == [0, 1, 2, 3]
    with pytest.raises(ValueError, match="Unknown method"):
        solver_id("cholesky")


def test_hot_path_does_not_touch_tensor_network_contract(monkeypatch):
    """ALS / fast_contract / multi_derivative go to the CUDA library, never through the host einsum utility."""
    import inspect

    from bsde_solver import utils
    from bsde_solver.core.calculus import derivative
    from bsde_solver.core.optimisation import als

    for mod in (als, utils, derivative):
        src = inspect.getsource(mod)
        assert ".contract(" not in src, mod.__name__


def test_sweep_driver_counts_the_algorithmic_flops_of_the_survey():
    """bench/sweep.py reports fractions of the FP64 peak with the reference algorithm's flop count; the three
    configurations tabulated in SURVEY.md §8(d) pin the formula (n(n+1)+2n per micro-step plus the interface update)."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_sweep", os.path.join(root, "bench", "sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.flops_per_sample_sweep(100, 4, 4) == 866208       # C3
    assert mod.flops_per_sample_sweep(10, 2, 3) == 3396          # C2
    assert mod.flops_per_sample_sweep(500, 4, 4) == 4399008      # C5, d = 500


def test_starting_points_of_the_path_simulation():
    """generate_trajectories(X0, ...) (path.py:6-34): a tiled X0 — what every reference script passes — ships one row, a
    general (batch, dim) X0 ships every sample's starting point."""
    from bsde_solver.stochastic.path import _x0_row

    row = np.array([1.0, 0.5, 1.0])
    x0, n = _x0_row(row)
    assert x0.shape == (3,) and n is None
    x0, n = _x0_row(np.tile(row, (7, 1)))
    assert x0.shape == (3,) and n == 7 and x0.flags["C_CONTIGUOUS"]
    X0 = np.arange(12.0).reshape(4, 3)
    x0, n = _x0_row(X0[:, ::1])
    assert x0.shape == (4, 3) and n == 4
    np.testing.assert_array_equal(x0, X0)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line with the base contract's keys, the true sample
    count of the bounded CPU run in its config, no GPU work — and under a launcher every rank but 0 exits without output."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
    other = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=300)
    assert other.returncode == 0 and other.stdout.strip() == "", other.stdout[-500:] + other.stderr[-500:]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("samples*sweeps/s") and j["unit"] == "samples*sweeps/s"
    assert j["higher_is_better"] is True and j["gpu_launches"] == 0 and j["steps"] == 1 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    n = j["config"]["samples_per_gpu"]
    assert n < (1 << 20) and f"N=2^{n.bit_length() - 1} samples" in j["config"]["workload"]      # the label says what was run
