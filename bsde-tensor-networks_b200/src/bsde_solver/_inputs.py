"""Conversion of the reference's calling convention (lists of (N, p) basis-value arrays on the host) to the
device layout of the CUDA library (PHI [d][p][N], sample-contiguous)."""
from __future__ import annotations

import numpy as np

from bsde_solver._backend import BASIS_PHI, DeviceInputs, get_backend


def phis_list(phis):
    """A list of (N, p) arrays from a list or a TensorNetwork of TensorCores."""
    if hasattr(phis, "cores"):
        phis = list(phis.cores.values())
    return list(phis)


def stack_phis(phis) -> np.ndarray:
    """[(N, p)] * d  ->  contiguous float64 (d, p, N).  The reference's basis arrays are F-ordered (N, p) views
    (basis.py:21 builds (p, N) and transposes), so `.T` is already contiguous and this is one memcpy per asset."""
    phis = phis_list(phis)
    p = {int(np.shape(phi)[1]) for phi in phis}
    if len(p) != 1:
        raise ValueError("the CUDA backend needs the same number of basis functions for every asset")
    n = {int(np.shape(phi)[0]) for phi in phis}
    if len(n) != 1:
        raise ValueError("all basis arrays must have the same number of samples")
    out = np.empty((len(phis), p.pop(), n.pop()), dtype=np.float64)
    for i, phi in enumerate(phis):
        out[i] = np.asarray(phi).view(np.ndarray).T
    return out


def as_inputs(phis, dphis=None) -> DeviceInputs:
    """DeviceInputs are passed through; host lists are stacked and uploaded (explicit-values mode)."""
    if isinstance(phis, DeviceInputs):
        return phis
    be = get_backend()
    host = stack_phis(phis)
    ddata = None if dphis is None else be.to_device(stack_phis(dphis))
    return DeviceInputs(BASIS_PHI, host.shape[1], be.to_device(host), ddata=ddata)
