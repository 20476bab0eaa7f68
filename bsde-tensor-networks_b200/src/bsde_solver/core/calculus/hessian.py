"""hessian — second derivatives of a tensor-train function at the samples
(reference: core/calculus/hessian.py:6-26): H[i, j](x_s) = d^2 V / dx_i dx_j.

Same signature as the reference: `hessian(tt, x, dx, ddx, batch=False)` with the per-asset lists of basis values, first and
second derivative values; `batch=True` returns (d, d, N) as the reference does, `batch=False` takes one point (each entry a
vector of p values) and returns (d, d).  A `DeviceInputs` handle for `x` (then dx / ddx are not needed: the basis is
evaluated in the kernel) returns a CUDA tensor [d][d][N].  The reference contracts the whole network once per index pair;
the GPU path is one right-interface pass and one row kernel per asset (bsde_tt_hessian)."""
from bsde_solver import xp


def hessian(tt, x, dx=None, ddx=None, batch=False):
    from bsde_solver._backend import DeviceInputs, get_backend
    from bsde_solver._inputs import as_inputs, phis_list
    from bsde_solver.core.tensor.tensor_train import pack_cores

    ranks, flat = pack_cores(tt)
    be = get_backend()
    if isinstance(x, DeviceInputs):
        return be.tt_hessian(x, ranks, flat)
    if dx is None or ddx is None:
        raise ValueError("hessian needs dx and ddx for explicit basis values")
    if not batch:        # one point: vectors of p values -> a batch of one
        x, dx, ddx = ([xp.asarray(v, dtype=xp.float64).reshape(1, -1) for v in phis_list(lst)] for lst in (x, dx, ddx))
    h = be.tt_hessian(as_inputs(x, dx, ddx), ranks, flat).cpu().numpy()
    return h if batch else h[:, :, 0]
