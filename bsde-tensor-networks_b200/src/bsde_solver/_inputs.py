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


def asset_blocks(phis):
    """The list as d contiguous (p, N) host blocks, without a stacked copy where the (N, p) arrays are F-ordered (the
    reference's own layout, basis.py:21); anything else is copied asset by asset."""
    phis = phis_list(phis)
    blocks = []
    for phi in phis:
        a = np.asarray(phi).view(np.ndarray)
        if a.ndim != 2:
            raise ValueError("basis arrays must be (N, p)")
        t = a.T
        blocks.append(t if (t.flags["C_CONTIGUOUS"] and t.dtype == np.float64) else np.ascontiguousarray(t, dtype=np.float64))
    if len({b.shape for b in blocks}) != 1:
        raise ValueError("the CUDA backend needs the same number of samples and of basis functions for every asset")
    return blocks


def tagged_points(phis):
    """(kind, p, coef, [x_j]) when every entry of the list is an untouched result of one of this package's basis objects
    (same basis for all assets), else None."""
    phis = phis_list(phis)
    tags = [getattr(phi, "basis_tag", None) for phi in phis]
    if not tags or any(t is None for t in tags):
        return None
    kind, p, coef, _ = tags[0]
    for phi, t in zip(phis, tags):
        if t[0] != kind or t[1] != p or np.shape(phi) != (t[3].shape[0], p):
            return None
        if (coef is None) != (t[2] is None) or (coef is not None and not np.array_equal(coef, t[2])):
            return None
    if len({t[3].shape[0] for t in tags}) != 1:
        return None
    return kind, p, coef, [t[3] for t in tags]


def as_inputs(phis, dphis=None, ddphis=None) -> DeviceInputs:
    """DeviceInputs are passed through; host lists are stacked and uploaded (explicit-values mode)."""
    if isinstance(phis, DeviceInputs):
        return phis
    be = get_backend()
    host = stack_phis(phis)
    ddata = None if dphis is None else be.to_device(stack_phis(dphis))
    d2data = None if ddphis is None else be.to_device(stack_phis(ddphis))
    return DeviceInputs(BASIS_PHI, host.shape[1], be.to_device(host), ddata=ddata, d2data=d2data)
