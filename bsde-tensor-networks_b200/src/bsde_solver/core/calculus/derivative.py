"""multi_derivative — gradient of a tensor-train function at every sample
(reference: core/calculus/derivative.py:43-85): dV/dx_i(x_s) = L_i[s] (C_i . dphi_i[s]) R_i[s].
One right-interface pass plus one fused left pass on the GPU (bsde_tt_grad -> k_grad_step)."""
from bsde_solver import xp


def multi_derivative(tt, phis, dphis=None):
    """Host lists in -> (N, d) numpy array as the reference; DeviceInputs in -> CUDA tensor Z [d][N]."""
    from bsde_solver._backend import DeviceInputs, get_backend
    from bsde_solver._inputs import as_inputs
    from bsde_solver.core.tensor.tensor_train import pack_cores

    on_device = isinstance(phis, DeviceInputs)
    if not on_device and dphis is None:
        raise ValueError("multi_derivative needs dphis for explicit basis values")
    ranks, flat = pack_cores(tt)
    z = get_backend().tt_grad(as_inputs(phis, dphis), ranks, flat)
    if on_device:
        return z
    return xp.ascontiguousarray(z.cpu().numpy().T)
