"""ALS and SALSA — tensor-train regression of `result` on the per-asset basis values `phis`
(reference: core/optimisation/als.py:26-96 and :99-323).

Same signatures and the same contract as the reference: `init_tt` is re-randomised from the global numpy stream,
mutated in place and returned.  The whole optimisation — right-orthonormalisation, interface contractions,
normal-equation assembly on the FP64 tensor pipe, the small solves, QR/SVD steps and SALSA's rank adaptation —
runs on the GPU behind one C-ABI call (bsde_fit_als / bsde_fit_salsa).

`phis` is either the reference's list of (N, p) arrays (host; uploaded by the call) or a `DeviceInputs` handle to
data already on the GPU (then `result` may be a CUDA tensor too and nothing crosses PCIe but the cores).
"""
from __future__ import annotations

import warnings

import numpy as np

from bsde_solver import xp
from bsde_solver._backend import BASIS_PHI, DeviceInputs, get_backend
from bsde_solver._inputs import asset_blocks, phis_list, stack_phis, tagged_points  # noqa: F401
from bsde_solver.core.optimisation.solve import solver, solver_id  # noqa: F401  (solver re-exported as in the reference)
from bsde_solver.core.tensor.tensor_core import TensorCore  # noqa: F401
from bsde_solver.core.tensor.tensor_network import TensorNetwork  # noqa: F401
from bsde_solver.core.tensor.tensor_train import TensorTrain, left_unfold, pack_cores, right_unfold, unpack_cores  # noqa: F401

last_fit_info = {}
RANK_LIMIT = 8      # largest bond rank of the fused CUDA kernels (csrc/api.cu: kRankLimit)


class RankLimitWarning(RuntimeWarning):
    """SALSA wanted to grow a bond beyond RANK_LIMIT; set warnings.simplefilter("error", RankLimitWarning) to make it fatal."""



def _shape_of(phis):
    if isinstance(phis, DeviceInputs):
        return tuple([phis.p] * phis.d)
    return tuple(int(np.shape(phi)[1]) for phi in phis_list(phis))


def _is_device_tensor(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


def ALS(phis, result, n_iter=10, ranks=None, optimizer="lstsq", init_tt=None):
    shape = _shape_of(phis)
    tt = init_tt if init_tt else TensorTrain(shape, ranks)
    tt.randomize()                                   # als.py:37 — always, a warm start only carries the shapes
    sid = solver_id(optimizer)
    be = get_backend()
    tt_ranks, flat = pack_cores(tt)
    if isinstance(phis, DeviceInputs):
        y = result if _is_device_tensor(result) else be.to_device(np.asarray(result, dtype=np.float64).ravel())
        info = be.fit_als(phis, y, n_iter, tt_ranks, sid, flat)
    else:
        y = np.ascontiguousarray(np.asarray(result, dtype=np.float64).ravel())
        tagged = tagged_points(phis)
        if tagged is not None:
            # basis values of this package's own basis objects: ship the points, evaluate the basis in the kernels
            kind, p, coef, blocks = tagged
        else:
            # arbitrary basis values: the (N, p) arrays' transposes are consumed where they lie (no stacked host copy)
            kind, coef, blocks = BASIS_PHI, None, asset_blocks(phis)
            p = blocks[0].shape[0]
        if y.shape[0] != blocks[0].shape[-1]:
            raise ValueError(f"result has {y.shape[0]} samples, phis have {blocks[0].shape[-1]}")
        info = be.fit_als_host_assets(kind, p, blocks, coef, y, n_iter, tt_ranks, sid, flat)
        last_fit_info["inputs"] = "points (tagged basis values)" if tagged is not None else "explicit basis values"

    last_fit_info.update(cholesky=int(info[0]), jacobi=int(info[1]), lu=int(info[2]), min_rank=int(info[3]))
    return unpack_cores(tt, tt_ranks, flat)


def SALSA(phis, result, n_iter=10, ranks=None, omega=1.0, freq_omega=1.05, min_sv=0.2, init_tt=None, max_rank=10,
          optimizer="lstsq", do_reg: bool = True):
    shape = _shape_of(phis)
    tt = init_tt if init_tt else TensorTrain(shape, ranks)
    tt.randomize()                                   # als.py:116
    sid = solver_id(optimizer)
    be = get_backend()
    d, p = len(shape), shape[0]
    cur_ranks, flat = pack_cores(tt)
    # the reference computes its rank caps from the *constructor* ranks, which it never updates (als.py:122-123)
    init_ranks = np.asarray(list(tt.ranks), dtype=np.int32)
    cap = max(int(max_rank), int(cur_ranks.max()))
    buf = np.zeros(d * cap * p * cap, dtype=np.float64)
    buf[: flat.size] = flat
    from bsde_solver._inputs import as_inputs

    inp = as_inputs(phis)
    y = result if _is_device_tensor(result) else be.to_device(np.asarray(result, dtype=np.float64).ravel())
    ranks_io = np.array(cur_ranks, dtype=np.int32)
    state = be.fit_salsa(inp, y, n_iter, ranks_io, init_ranks, omega, freq_omega, min_sv, max_rank, sid, do_reg, buf)
    last_fit_info.update(eta=float(state[0]), omega=float(state[1]), min_sv=float(state[2]), rank_limited=bool(state[3]),
                         cholesky=int(state[4]), jacobi=int(state[5]), lu=int(state[6]))
    if state[3]:
        # the reference would have grown a bond beyond what the CUDA kernels hold (als.py:122-123, :249-253): the fit
        # that comes back is a valid SALSA fit with that bond capped, but NOT the reference's rank trajectory
        warnings.warn(f"SALSA: rank growth stopped at the CUDA backend's limit of {RANK_LIMIT} (max_rank={max_rank} allows "
                      f"more); ranks now {list(map(int, ranks_io))}", RankLimitWarning, stacklevel=2)
    return unpack_cores(tt, ranks_io, buf)
