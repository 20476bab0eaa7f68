"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, read-only) on top of oracle/shim/ (opt_einsum + matplotlib stand-ins).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py            # re-execs itself with PYTHONHASHSEED=0

The reference has no asserting tests or golden vectors, so these fixtures — outputs of the
reference's own functions under fixed numpy seeds — are what pins the oracle (oracle/tt_oracle.py) and,
through it, the CUDA path.  PYTHONHASHSEED is pinned because tensor_network.py:78 orders batched
contraction outputs through `list(set(...))`.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

sys.path[:0] = [os.path.join(HERE, "shim"), REF, os.path.join(REF, "src")]

import numpy as np  # noqa: E402

from bsde_solver.core.tensor.tensor_core import TensorCore  # noqa: E402
from bsde_solver.core.tensor.tensor_train import TensorTrain  # noqa: E402
from bsde_solver.core.optimisation import als as ref_als  # noqa: E402
from bsde_solver.core.calculus.basis import PolynomialBasis, LegendreBasis  # noqa: E402
from bsde_solver.core.calculus.derivative import multi_derivative  # noqa: E402
from bsde_solver.utils import fast_contract  # noqa: E402
from bsde_solver.stochastic.path import generate_trajectories  # noqa: E402
from bsde_solver.bsde import BlackScholes, HJB  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def cores_of(tt):
    return [np.array(tt.cores[f"core_{i}"].view(np.ndarray)) for i in range(tt.order)]


def pack(prefix, arrays):
    return {f"{prefix}{i}": a for i, a in enumerate(arrays)}


def make_phis(x, basis, kind="eval"):
    f = getattr(basis, kind)
    return [TensorCore(f(x[:, i]), name=f"phi_{i+1}", indices=("batch", f"m_{i+1}")) for i in range(x.shape[1])]


class Capture:
    """Records what the reference feeds to / gets from its solver, SVD and randomize calls."""

    def __init__(self):
        self.solves, self.inits, self.svds = [], [], []
        self._solver = ref_als.solver
        self._rand = TensorTrain.randomize
        self._svd = np.linalg.svd

    def __enter__(self):
        cap = self

        def solver(A, Y, method):
            x = cap._solver(A, Y, method)
            cap.solves.append((np.array(A.view(np.ndarray)), np.array(Y.view(np.ndarray)), np.array(x)))
            return x

        def randomize(tt):
            out = cap._rand(tt)
            cap.inits.append(cores_of(tt))
            return out

        def svd(a, *args, **kw):
            out = cap._svd(a, *args, **kw)
            cap.svds.append((np.array(a), [np.array(o) for o in out]))
            return out

        ref_als.solver = solver
        TensorTrain.randomize = randomize
        np.linalg.svd = svd
        return self

    def __exit__(self, *a):
        ref_als.solver = self._solver
        TensorTrain.randomize = self._rand
        np.linalg.svd = self._svd


def golden_basis():
    np.random.seed(11)
    x = np.concatenate([np.random.randn(150) * 1.5, np.array([0.0, 1.0, -1.0, 1e-3, 7.25, -3.5])])
    out = {"x": x}
    for p in (3, 4, 6):
        b = PolynomialBasis(p)
        out[f"poly_eval_{p}"], out[f"poly_grad_{p}"], out[f"poly_dgrad_{p}"] = b.eval(x), b.grad(x), b.dgrad(x)
    for p in (3, 5):
        b = LegendreBasis(p)
        out[f"leg_eval_{p}"], out[f"leg_grad_{p}"], out[f"leg_dgrad_{p}"] = b.eval(x), b.grad(x), b.dgrad(x)
    np.savez_compressed(os.path.join(OUT, "basis.npz"), **out)


def golden_orthonormalize():
    np.random.seed(12)
    shape, ranks = [3, 4, 3, 3, 4], [1, 2, 3, 3, 2, 1]
    out = {"shape": np.array(shape), "ranks": np.array(ranks)}
    tt = TensorTrain(shape, ranks)
    tt.randomize()
    out.update(pack("in_", cores_of(tt)))
    a = tt.copy(); a.orthonormalize(mode="right", start=1); out.update(pack("right1_", cores_of(a)))
    a = tt.copy(); a.orthonormalize(mode="right"); out.update(pack("right0_", cores_of(a)))
    a = tt.copy(); a.orthonormalize(mode="left"); out.update(pack("left0_", cores_of(a)))
    a = tt.copy(); a.orthonormalize(mode="left", start=1, stop=3); out.update(pack("left13_", cores_of(a)))
    np.savez_compressed(os.path.join(OUT, "orthonormalize.npz"), **out)


def golden_eval_grad():
    np.random.seed(13)
    d, p, r, N = 6, 4, 3, 300
    ranks = [1] + [r] * (d - 1) + [1]
    tt = TensorTrain([p] * d, ranks)
    tt.randomize()
    x = np.random.randn(N, d) * 0.8
    out = {"x": x, "p": p}
    out.update(pack("core_", cores_of(tt)))
    for name, basis in (("poly", PolynomialBasis(p)), ("leg", LegendreBasis(p))):
        phis, dphis = make_phis(x, basis), make_phis(x, basis, "grad")
        out[f"eval_{name}"] = np.array(fast_contract(tt, phis).view(np.ndarray)).squeeze()
        out[f"grad_{name}"] = np.array(multi_derivative(tt, phis, dphis))
    np.savez_compressed(os.path.join(OUT, "eval_grad.npz"), **out)


def golden_als(name, x, y, p, r, n_iter, basis_cls=PolynomialBasis, keep_solves=None):
    d = x.shape[1]
    ranks = (1,) + (r,) * (d - 1) + (1,)
    phis = make_phis(x, basis_cls(p))
    with Capture() as cap:
        tt = ref_als.ALS(phis, y, n_iter=n_iter, ranks=ranks, optimizer="lstsq")
    fitted = np.array(fast_contract(tt, make_phis(x, basis_cls(p))).view(np.ndarray)).squeeze()
    out = {"x": x, "y": y, "p": p, "r": r, "n_iter": n_iter, "fitted": fitted}
    out.update(pack("init_", cap.inits[0]))
    out.update(pack("final_", cores_of(tt)))
    ks = len(cap.solves) if keep_solves is None else keep_solves
    for k in range(ks):
        A, b, sol = cap.solves[k]
        out[f"A_{k}"], out[f"b_{k}"], out[f"x_{k}"] = A, b, sol
    out["n_solves_kept"] = ks
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "residual", np.linalg.norm(fitted - y) / np.linalg.norm(y))


def golden_als_kat():
    # tests/optimisation/test.py:35-64 — seed 216540, ||x||^2, d=6, p=3, r=2, N=1000, 10 sweeps
    np.random.seed(216540)
    x = np.array([np.random.rand(6) - 1 / 2 for _ in range(1000)])
    y = np.linalg.norm(x, axis=1) ** 2
    golden_als("als_kat.npz", x, y, p=3, r=2, n_iter=10, keep_solves=20)


def golden_als_c2():
    np.random.seed(21)
    d, N = 10, 2048
    X0 = np.tile(np.array([1.0, 0.5] * (d // 2)), (N, 1))
    model = BlackScholes(X0, 1 / 50, 1, 0.05, 0.4)
    X, _ = generate_trajectories(X0, 50, model)
    x = X[:, 25]
    y = model.g(X[:, 25]) * 1.01 + 0.1 * np.sin(x[:, 0])
    golden_als("als_c2small.npz", x, y, p=3, r=2, n_iter=4, keep_solves=18)


def golden_als_c3():
    np.random.seed(22)
    d, N = 12, 4096
    x = np.random.randn(N, d) * np.sqrt(2 * 0.5)
    y = np.log(0.5 + 0.5 * np.sum(x**2, axis=1))
    golden_als("als_c3small.npz", x, y, p=4, r=4, n_iter=4, keep_solves=22)


def golden_als_legendre():
    np.random.seed(23)
    d, N = 5, 512
    x = np.random.rand(N, d) * 2 - 1
    y = np.cos(np.sum(x, axis=1))
    golden_als("als_legendre.npz", x, y, p=4, r=3, n_iter=3, basis_cls=LegendreBasis, keep_solves=8)


def golden_salsa():
    # tests/optimisation/test.py:91-114 shapes: d=6, p=3, r=2, N=1000, max_rank=5
    np.random.seed(216540)
    x = np.array([np.random.rand(6) - 1 / 2 for _ in range(1000)])
    y = np.linalg.norm(x, axis=1) ** 2 + 0.05 * np.sin(3 * x[:, 0] * x[:, 1])
    d, p, r = 6, 3, 2
    ranks = (1,) + (r,) * (d - 1) + (1,)
    for n_iter, tag in ((4, "a"), (7, "b")):
        np.random.seed(5)
        phis = make_phis(x, PolynomialBasis(p))
        with Capture() as cap:
            tt = ref_als.SALSA(phis, y, n_iter=n_iter, ranks=ranks, max_rank=5)
        fitted = np.array(fast_contract(tt, make_phis(x, PolynomialBasis(p))).view(np.ndarray)).squeeze()
        out = {"x": x, "y": y, "p": p, "r": r, "n_iter": n_iter, "max_rank": 5, "fitted": fitted}
        out.update(pack("init_", cap.inits[0]))
        out.update(pack("final_", cores_of(tt)))
        for k, (A, b, sol) in enumerate(cap.solves):
            out[f"A_{k}"], out[f"b_{k}"], out[f"x_{k}"] = A, b, sol
        out["n_solves_kept"] = len(cap.solves)
        for k, (a, (U, S, Vt)) in enumerate(cap.svds):
            out[f"svd_in_{k}"], out[f"svd_S_{k}"] = a, S
        out["n_svds"] = len(cap.svds)
        np.savez_compressed(os.path.join(OUT, f"salsa_{tag}.npz"), **out)
        print("salsa", tag, "final ranks", [c.shape[0] for c in cores_of(tt)], "residual",
              np.linalg.norm(fitted - y) / np.linalg.norm(y))


def golden_paths():
    out = {}
    np.random.seed(31)
    N, d, steps = 64, 4, 10
    X0 = np.tile(np.array([1.0, 0.5, 1.0, 0.5]), (N, 1))
    x, xi = generate_trajectories(X0, steps, BlackScholes(X0, 1 / steps, 1, 0.05, 0.4))
    out["bs_x"], out["bs_xi"] = x, xi
    X0 = np.zeros((N, d))
    model = HJB(X0, 1 / steps, 1, np.sqrt(2))
    x, xi = generate_trajectories(X0, steps, model)
    out["hjb_x"], out["hjb_xi"] = x, xi
    z = np.random.randn(N, d)
    yv = np.random.randn(N)
    out["z"], out["yv"] = z, yv
    out["hjb_h"], out["hjb_g"] = model.h(x[:, 3], 0.3, yv, z), model.g(x[:, -1])
    bs = BlackScholes(X0, 1 / steps, 1, 0.05, 0.4)
    out["bs_h"], out["bs_g"] = bs.h(out["bs_x"][:, 3], 0.3, yv, z), bs.g(out["bs_x"][:, -1])
    np.savez_compressed(os.path.join(OUT, "paths.npz"), **out)


def golden_pipeline_c1():
    """BASELINE config 1: src/scripts/pipeline.py with line 26 flipped to mode = "ALS" (nothing else changed).
    The module body is executed from its source text so /root/reference stays untouched."""
    path = os.path.join(REF, "src", "scripts", "pipeline.py")
    src = open(path).read()
    assert 'mode = "SALSA"\n# mode = "ALS"' in src
    src = src.replace('mode = "SALSA"\n# mode = "ALS"', 'mode = "ALS"')
    out = {}
    for seed in (0, 1):
        np.random.seed(seed)
        g = {"__name__": "golden_pipeline", "__file__": path}
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(src, path, "exec"), g)
        Y = g["Y"]
        out[f"Y0_seed{seed}"] = np.mean(Y[:, 0])
        out[f"Ymean_seed{seed}"] = Y.mean(axis=0)
        out[f"Ysample_seed{seed}"] = Y[:8]
        out[f"X_last_seed{seed}"] = g["X"][:8, -1]
        print("pipeline C1 seed", seed, "Y0 =", repr(np.mean(Y[:, 0])), "analytic", g["price"])
    np.savez_compressed(os.path.join(OUT, "pipeline_c1.npz"), **out)


def golden_explicit_small():
    """Explicit ALS scheme on HJB (the C3 model) at d=6, N=2048, 5 steps, p=3, r=2 — loop of src/scripts/pipeline.py:84-141."""
    np.random.seed(41)
    N, d, steps, p, r, n_iter = 2048, 6, 5, 3, 2, 6
    dt = 1 / steps
    ranks = (1,) + (r,) * (d - 1) + (1,)
    X0 = np.zeros((N, d))
    model = HJB(X0, dt, 1, np.sqrt(2))
    X, xi = generate_trajectories(X0, steps, model)
    basis = PolynomialBasis(p)
    Y = np.zeros((N, steps + 1))
    Y[:, -1] = model.g(X[:, -1])
    inits = []
    with Capture() as cap:
        V = ref_als.ALS(make_phis(X[:, -1], basis), Y[:, -1], n_iter=n_iter, ranks=ranks)
        for n in range(steps - 1, -1, -1):
            Z = multi_derivative(V, make_phis(X[:, n + 1], basis), make_phis(X[:, n + 1], basis, "grad"))
            h = model.h(X[:, n + 1], (n + 1) * dt, Y[:, n + 1], Z)
            target = h * dt + Y[:, n + 1]
            V = ref_als.ALS(make_phis(X[:, n], basis), target, n_iter=n_iter, ranks=ranks, init_tt=V)
            Y[:, n] = np.array(fast_contract(V, make_phis(X[:, n], basis)).view(np.ndarray)).squeeze()
        inits = cap.inits
    out = {"xi": xi, "X": X, "Y": Y, "p": p, "r": r, "n_iter": n_iter, "Z_last": Z}
    for k, cs in enumerate(inits):
        out.update(pack(f"init{k}_", cs))
    out["n_fits"] = len(inits)
    np.savez_compressed(os.path.join(OUT, "explicit_hjb_small.npz"), **out)
    print("explicit small Y0", Y[:, 0].mean())


def golden_hessian():
    """hessian() of the reference (core/calculus/hessian.py:6-26) at a batch of points and at a single point, for a random
    train in monomials and in Legendre polynomials; used by the parity tests of bsde_tt_hessian."""
    from bsde_solver.core.calculus.hessian import hessian

    np.random.seed(77)
    N, d, r = 40, 5, 3
    out = {"x": np.random.randn(N, d) * 0.8}
    for name, basis in (("poly", PolynomialBasis(4)), ("leg", LegendreBasis(3))):
        p = basis.degree
        tt = TensorTrain([p] * d, (1,) + (r,) * (d - 1) + (1,))
        tt.randomize()
        out.update(pack(f"cores_{name}_", cores_of(tt)))
        x = out["x"]
        H = hessian(tt, make_phis(x, basis), make_phis(x, basis, "grad"), make_phis(x, basis, "dgrad"), batch=True)
        out[f"H_{name}"] = np.array(H)
        def single(kind, k=3):
            f = getattr(basis, kind)
            return [TensorCore(f(x[k:k + 1, i])[0], name=f"phi_{i+1}", indices=(f"m_{i+1}",)) for i in range(d)]
        out[f"H1_{name}"] = np.array(hessian(tt, single("eval"), single("grad"), single("dgrad"), batch=False))
        out[f"p_{name}"] = p
    np.savez_compressed(os.path.join(OUT, "hessian.npz"), **out)
    print("hessian fixture: batch", out["H_poly"].shape, "single", out["H1_poly"].shape,
          "consistent", np.abs(out["H_poly"][:, :, 3] - out["H1_poly"]).max())


def golden_pipeline_c2():
    """BASELINE config 2: Black-Scholes basket, d = 10, N = 2^16 paths, rank 2, 3 monomials, 50 steps, explicit ALS with
    10 sweeps per fit — the loop of src/scripts/pipeline.py:84-141 driven by the reference's own functions (the script's
    literals are for d = 4).  ~10 minutes on 8 cores.  Stores Y0, the per-step means of Y and a few paths; the Brownian
    increments are reproducible from the numpy seed (randn(N, 51, 10) is the first draw)."""
    import time
    seed = 0
    np.random.seed(seed)
    N, d, steps, p, r, n_iter = 1 << 16, 10, 50, 3, 2, 10
    dt = 1 / steps
    ranks = (1,) + (r,) * (d - 1) + (1,)
    X0 = np.tile(np.array([1.0, 0.5] * (d // 2)), (N, 1))
    model = BlackScholes(X0, dt, 1, 0.05, 0.4)
    X, xi = generate_trajectories(X0, steps, model)
    basis = PolynomialBasis(p)
    Y = np.zeros((N, steps + 1))
    Y[:, -1] = model.g(X[:, -1])
    t0 = time.time()
    V = ref_als.ALS(make_phis(X[:, -1], basis), Y[:, -1], n_iter=n_iter, ranks=ranks)
    for n in range(steps - 1, -1, -1):
        Z = multi_derivative(V, make_phis(X[:, n + 1], basis), make_phis(X[:, n + 1], basis, "grad"))
        h = model.h(X[:, n + 1], (n + 1) * dt, Y[:, n + 1], Z)
        V = ref_als.ALS(make_phis(X[:, n], basis), h * dt + Y[:, n + 1], n_iter=n_iter, ranks=ranks, init_tt=V)
        Y[:, n] = np.array(fast_contract(V, make_phis(X[:, n], basis)).view(np.ndarray)).squeeze()
        print("c2 step", n, "Y mean", Y[:, n].mean(), f"{time.time() - t0:.0f}s", flush=True)
    out = {"seed": seed, "Y0": Y[:, 0].mean(), "Ymean": Y.mean(axis=0), "Ysample": Y[:16], "X_last": X[:16, -1],
           "price": model.price(X0, 0), "seconds": time.time() - t0}
    np.savez_compressed(os.path.join(OUT, "pipeline_c2.npz"), **out)
    print("pipeline C2 Y0 =", repr(out["Y0"]), "analytic", out["price"], "reference time", out["seconds"])


if __name__ == "__main__":
    if len(sys.argv) > 1:          # python oracle/make_golden.py golden_pipeline_c2  -> only that fixture
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    golden_basis()
    golden_orthonormalize()
    golden_eval_grad()
    golden_als_kat()
    golden_als_c2()
    golden_als_c3()
    golden_als_legendre()
    golden_salsa()
    golden_paths()
    golden_explicit_small()
    golden_hessian()
    golden_pipeline_c1()
    print("done ->", OUT)
