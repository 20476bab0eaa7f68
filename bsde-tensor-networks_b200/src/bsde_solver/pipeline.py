"""Backward time loops of the BSDE solver with everything resident on the GPU
(reference: src/scripts/pipeline.py:66-141 — explicit scheme; src/scripts/pipeline_explicit.py:35-112 and
src/scripts/experiments/n_assets.py:128-198 — implicit scheme with fixed-point iterations).

Per time step the host only sequences library calls: gradient of the current value function (k_grad_step), target
formation (k_target), one TT regression (bsde_fit_als / bsde_fit_salsa) and the evaluation Y_n = V_n(X_n) (k_tt_eval).
Paths, targets, Y and Z never leave the device; the cores (a few KB) are the only data that crosses PCIe.

`paths="numpy"` draws the Brownian increments from the global numpy stream exactly as the reference does
(randn(batch, N+1, d), stochastic/path.py:26) and replays them on the GPU — with the same seed the solve consumes the
same random numbers as the reference, fit by fit.  `paths="philox"` draws them inside the path kernel.
"""
from __future__ import annotations

import time
from functools import partial

from bsde_solver import xp
from bsde_solver.core.calculus.derivative import multi_derivative
from bsde_solver.core.optimisation.als import ALS, SALSA
from bsde_solver.stochastic.path import generate_trajectories_device
from bsde_solver.utils import fast_contract


def simulate_paths(model, X0, n_steps, batch_size, paths="philox", seed=0, sample_offset=0, keep_noise=True):
    """x, xi as CUDA tensors [n_steps+1][d][batch] (xi is None with keep_noise=False in philox mode)."""
    from bsde_solver._backend import get_backend

    be = get_backend()
    from bsde_solver.stochastic.path import _x0_row

    x0, _ = _x0_row(X0)                       # one row when all samples start from the same point, else (batch, d)
    if paths == "numpy":
        xi_host = xp.random.randn(batch_size, n_steps + 1, x0.shape[-1])                   # path.py:26
        xi = be.to_device(xp.ascontiguousarray(xi_host.transpose(1, 2, 0)))
        return be.paths_simulate(model.model_id, model.params(), x0, batch_size, n_steps, model.T, xi=xi)
    if paths == "philox":
        return generate_trajectories_device(x0, n_steps, model, batch_size=batch_size, seed=seed, sample_offset=sample_offset,
                                            keep_noise=keep_noise)
    raise ValueError(f"unknown paths mode {paths}")


def _is_salsa(optimizer):
    """True for SALSA itself and for functools.partial wrappers around it (any depth)."""
    while isinstance(optimizer, partial):
        optimizer = optimizer.func
    return optimizer is SALSA


def _mean(v):
    """Mean over all samples of all ranks (sample-sharded runs combine the per-rank sums)."""
    from bsde_solver._backend import get_backend

    be = get_backend()
    m = be.mean(v)
    if be.world > 1:
        import torch
        import torch.distributed as dist

        t = torch.tensor([m * v.numel(), float(v.numel())], dtype=torch.float64, device=v.device)
        dist.all_reduce(t)
        m = float(t[0] / t[1])
    return m


def run_explicit(model, basis, X, n_iter, ranks, optimizer=ALS, apply_sigma=False, verbose=False, salsa_no_reg_at_0=False):
    """Explicit scheme: target_n = h(X_{n+1}, Y_{n+1}, Z_{n+1}) dt + Y_{n+1}, Z = grad V_{n+1}(X_{n+1})
    (pipeline.py:129-137; apply_sigma=True gives the sigma-scaled Z of experiments/n_assets.py:101-103).
    X: CUDA tensor [N+1][d][batch].  Returns (Y [N+1][batch] CUDA tensor, list of value functions, step times)."""
    from bsde_solver._backend import get_backend

    be = get_backend()
    n_steps = X.shape[0] - 1
    dt = model.T / n_steps
    Y = be.empty(n_steps + 1, X.shape[2])
    Y[n_steps] = model.g(X[n_steps])
    inp = basis.device_inputs(X[n_steps])
    V = optimizer(inp, Y[n_steps], n_iter=n_iter, ranks=ranks)
    Vs = [None] * (n_steps + 1)
    Vs[n_steps] = V
    step_times = []
    for n in range(n_steps - 1, -1, -1):
        t0 = time.perf_counter()
        Z = multi_derivative(V, basis.device_inputs(X[n + 1]))
        target = be.target(model.model_id, model.params(), X[n + 1], Y[n + 1], Y[n + 1], Z, None, dt, apply_sigma)
        opt = optimizer
        if n == 0 and salsa_no_reg_at_0:
            opt = partial(optimizer, do_reg=False)                                          # pipeline.py:135-136
        inp = basis.device_inputs(X[n])
        V = opt(inp, target, n_iter=n_iter, ranks=ranks, init_tt=V)
        Vs[n] = V                      # NB: as in the reference all entries alias one TensorTrain object (als.py:36-37)
        Y[n] = fast_contract(V, inp)
        step_times.append(time.perf_counter() - t0)
        if verbose:
            print("Step:", n, "Step time:", f"{step_times[-1]:.3f}s")
    return Y, Vs, step_times


def run_implicit(model, basis, X, xi, n_iter, n_iter_implicit, ranks, optimizer=SALSA, verbose=False, no_reg_at_0=True):
    """Implicit scheme (fixed point in Y_n, Z_n at every step):
        Z  = sigma(X_n) grad V_nk(X_n);  target = h(X_n, Y_nk, Z) dt + Y_{n+1} - sqrt(dt) Z . xi_{n+1}
        V_nk <- fit(phi(X_n), target, init_tt=V_nk);  Y_nk = V_nk(X_n)
    (pipeline_explicit.py:94-109 in the intended form of experiments/n_assets.py:172-181)."""
    from bsde_solver._backend import get_backend

    be = get_backend()
    n_steps = X.shape[0] - 1
    dt = model.T / n_steps
    Y = be.empty(n_steps + 1, X.shape[2])
    Y[n_steps] = model.g(X[n_steps])
    inp = basis.device_inputs(X[n_steps])
    V = optimizer(inp, Y[n_steps], n_iter=n_iter, ranks=ranks)
    step_times = []
    for n in range(n_steps - 1, -1, -1):
        t0 = time.perf_counter()
        inp = basis.device_inputs(X[n])
        Y_nk = Y[n + 1]
        # the reference passes do_reg=False at n == 0 whenever the optimiser is SALSA (pipeline_explicit.py:103-104,
        # experiments/n_assets.py:176-179) — also when it arrives wrapped in functools.partial(SALSA, max_rank=...)
        opt = partial(optimizer, do_reg=False) if (n == 0 and no_reg_at_0 and _is_salsa(optimizer)) else optimizer
        for _ in range(n_iter_implicit):
            grad = multi_derivative(V, inp)
            target = be.target(model.model_id, model.params(), X[n], Y_nk, Y[n + 1], grad, xi[n + 1], dt, True)
            V = opt(inp, target, n_iter=n_iter, ranks=ranks, init_tt=V)
            Y_nk = fast_contract(V, inp)
        Y[n] = Y_nk
        step_times.append(time.perf_counter() - t0)
        if verbose:
            print("Step:", n, "Step time:", f"{step_times[-1]:.3f}s")
    return Y, V, step_times


def predicted_price(Y):
    """Y0 = mean over samples of Y[:, 0] (pipeline.py:170)."""
    return _mean(Y[0])
