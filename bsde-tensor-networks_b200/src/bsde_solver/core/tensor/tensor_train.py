"""TensorTrain — container of d cores (r_i, p, r_{i+1}) with the reference's names and index labels
(reference: core/tensor/tensor_train.py:16-126).  `orthonormalize` runs the QR chain on the GPU
(bsde_orthonormalize -> k_orthonormalize: in-CTA Householder QR with LAPACK's sign convention)."""
from __future__ import annotations

from copy import deepcopy

from bsde_solver import xp
from bsde_solver.core.tensor.tensor_core import TensorCore
from bsde_solver.core.tensor.tensor_network import TensorNetwork


def left_unfold(core: TensorCore):
    return core.unfold((0, 1), 2)


def right_unfold(core: TensorCore):
    return core.unfold(0, (1, 2))


def pack_cores(tt) -> tuple:
    """(ranks int32 (d+1), cores float64 flat) — the C ABI's wire format: cores concatenated in C order."""
    arrays = [xp.ascontiguousarray(xp.asarray(c).view(xp.ndarray), dtype=xp.float64) for c in tt.cores.values()]
    ranks = [a.shape[0] for a in arrays] + [arrays[-1].shape[2]]
    for k in range(len(arrays) - 1):
        if arrays[k].shape[2] != arrays[k + 1].shape[0]:
            raise ValueError(f"bond {k + 1}: core shapes {arrays[k].shape} / {arrays[k + 1].shape} do not match")
    return xp.asarray(ranks, dtype=xp.int32), xp.concatenate([a.ravel() for a in arrays])


def unpack_cores(tt, ranks, flat):
    """Write the flat cores back into tt.cores in place of the old TensorCores (same names and index labels)."""
    off = 0
    for i, (name, old) in enumerate(list(tt.cores.items())):
        shape = (int(ranks[i]), old.shape[1], int(ranks[i + 1]))
        n = shape[0] * shape[1] * shape[2]
        tt.cores[name] = TensorCore(xp.array(flat[off:off + n]).reshape(shape), indices=old.indices, name=old.name)
        off += n
    return tt


class TensorTrain(TensorNetwork):
    def __init__(self, shape: list[int], ranks: list[int]):
        cores = [
            TensorCore(xp.zeros((ranks[i], shape[i], ranks[i + 1])), indices=(f"r_{i}", f"m_{i+1}", f"r_{i+1}"))
            for i in range(len(shape))
        ]
        super().__init__(cores)
        self.shape = shape
        self.ranks = ranks
        self.order = len(shape)

    def orthonormalize(self, mode="right", start=0, stop=-1):
        """QR sweep making cores start..stop left- or right-orthonormal (tensor_train.py:34-87), on the device."""
        if mode not in ("left", "right"):
            raise ValueError("orthogonality must be either 'left' or 'right'")
        from bsde_solver._backend import get_backend

        shape = [int(c.shape[1]) for c in self.cores.values()]          # per-core basis sizes may differ (tensor_train.py:18-32)
        ranks, flat = pack_cores(self)
        get_backend().orthonormalize(shape[0] if len(set(shape)) == 1 else shape, ranks, flat, 1 if mode == "right" else 0, start, stop)
        unpack_cores(self, ranks, flat)

    @staticmethod
    def from_tensor(tensor, ranks: list[int]):
        """TT of a full tensor by successive truncated QRs (tensor_train.py:89-117).  Host-side constructor for
        small tensors (a d-way array is exponential in d): not part of the regression path."""
        shape = tensor.shape
        tt = TensorTrain(shape, ranks)
        rest = xp.asarray(tensor)
        for i in range(len(shape) - 1):
            q, r = xp.linalg.qr(rest.reshape(ranks[i] * shape[i], -1))
            tt.cores[f"core_{i}"] = TensorCore.like(q[:, :ranks[i + 1]], tt.cores[f"core_{i}"])
            rest = r[:ranks[i + 1], :]
        last = f"core_{len(shape) - 1}"
        tt.cores[last] = TensorCore.like(rest, tt.cores[last])
        return tt

    def randomize(self):
        for core in self.cores.values():
            core.randomize()
        return self

    def copy(self):
        return deepcopy(self)
