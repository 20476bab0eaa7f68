"""Forward Euler-Maruyama paths (reference: stochastic/path.py:6-34), simulated by the fused CUDA kernel k_paths.

generate_trajectories keeps the reference contract: increments xi = randn(batch, N+1, dim) are drawn from the global
numpy stream (so a seeded script consumes the RNG exactly as the reference does) and *replayed* on the GPU; x and xi
come back as host (batch, N+1, dim) arrays.  generate_trajectories_device is the production form: increments come
from the counter-based Philox generator inside the kernel and nothing leaves the GPU.
"""
from bsde_solver import xp
from bsde_solver.bsde import BackwardSDE


def _x0_row(X0):
    """(d,) -> itself; (batch, d) with identical rows (every reference script tiles one X0) -> that row, d doubles to
    ship; a general (batch, d) X0 -> itself, shipped sample by sample."""
    X0 = xp.asarray(X0, dtype=xp.float64)
    if X0.ndim == 1:
        return X0, None
    if not (X0 == X0[0]).all():
        return xp.ascontiguousarray(X0), X0.shape[0]
    return xp.ascontiguousarray(X0[0]), X0.shape[0]


def generate_trajectories(X0, N, model: BackwardSDE):
    """X0 (batch, dim), N time steps -> x (batch, N+1, dim), xi (batch, N+1, dim); xi[:, 0] is unused."""
    from bsde_solver._backend import get_backend

    batch_size, dim = X0.shape
    x0, _ = _x0_row(X0)
    xi = xp.random.randn(batch_size, N + 1, dim)                       # path.py:26
    be = get_backend()
    xi_dev = be.to_device(xp.ascontiguousarray(xi.transpose(1, 2, 0)))  # [N+1][dim][batch]
    x_dev, _ = be.paths_simulate(model.model_id, model.params(), x0, batch_size, N, model.T, xi=xi_dev)
    x = xp.ascontiguousarray(x_dev.permute(2, 0, 1).cpu().numpy())
    return x, xi


def generate_trajectories_device(X0, N, model: BackwardSDE, batch_size=None, seed=0, sample_offset=0, keep_noise=True):
    """Device-resident paths: x [N+1][dim][batch] (and xi) as CUDA tensors, increments from Philox4x32-10 keyed by
    `seed` and indexed by (sample_offset + sample, asset, step) — shards of one global sample set are reproducible
    independently of how many GPUs hold them."""
    from bsde_solver._backend import get_backend

    x0, n = _x0_row(X0)
    batch_size = batch_size if batch_size is not None else n
    if batch_size is None:
        raise ValueError("batch_size is required when X0 is a single row")
    return get_backend().paths_simulate(model.model_id, model.params(), x0, batch_size, N, model.T, seed=seed,
                                        sample_offset=sample_offset, keep_xi=keep_noise)
