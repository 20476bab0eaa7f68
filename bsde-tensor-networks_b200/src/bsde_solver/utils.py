"""fast_contract and small helpers (reference: utils.py:3-39)."""
from functools import partial
from typing import Any, Callable

from bsde_solver import xp
from bsde_solver.core.tensor.tensor_core import TensorCore
from bsde_solver.core.tensor.tensor_network import TensorNetwork  # noqa: F401


def matricized(core, mode="left"):
    if mode == "right":
        return core.reshape((core.shape[0], -1))
    if mode == "left":
        return core.reshape((-1, core.shape[-1]))
    raise ValueError("mode must be either 'left' or 'right'")


def tensorized(core, shape):
    return core.reshape(shape)


def flatten(lst):
    return [item for sublist in lst for item in sublist]


def fast_contract(tt, x):
    """V(x_s) for every sample (reference: utils.py:19-27) — one fused chain kernel (bsde_tt_eval -> k_tt_eval).
    Host lists in -> TensorCore with the single index 'batch' out (as the reference); DeviceInputs in -> CUDA tensor."""
    from bsde_solver._backend import DeviceInputs, get_backend
    from bsde_solver._inputs import as_inputs
    from bsde_solver.core.tensor.tensor_train import pack_cores

    on_device = isinstance(x, DeviceInputs)
    ranks, flat = pack_cores(tt)
    out = get_backend().tt_eval(as_inputs(x), ranks, flat)
    if on_device:
        return out
    return TensorCore(out.cpu().numpy(), indices=("batch",))


def callable_name(any_callable: Callable[..., Any]) -> str:
    if isinstance(any_callable, partial):
        return any_callable.func.__name__
    return getattr(any_callable, "__name__", str(any_callable))
