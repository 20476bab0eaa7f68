"""TensorNetwork — a named collection of TensorCores (reference: core/tensor/tensor_network.py:9-122).

In the reference every contraction of the regression path goes through `TensorNetwork.contract` ->
`opt_einsum.contract`.  Here the path's four contraction shapes are fixed-function CUDA kernels (interface updates,
normal-equation assembly, evaluation, gradient) reached through ALS / SALSA / fast_contract / multi_derivative, and
none of them calls `contract`.  The class is kept as the container the reference's callers build (`TensorTrain`
derives from it, scripts wrap `phis` in it); `contract` remains as a small host-side named-einsum utility for user
code that manipulates cores directly — it is not part of the accelerated path and is never used by it.
"""
from __future__ import annotations

from copy import deepcopy

from bsde_solver import xp
from bsde_solver.core.tensor.tensor_core import TensorCore

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


class TensorNetwork:
    def __init__(self, cores, names: list[str] = None):
        if isinstance(cores, dict):
            self.cores = cores
            return
        self.cores = {}
        for i, core in enumerate(cores):
            if core is None:
                continue
            name = names[i] if names else (core.name if core.name else f"core_{i}")
            if hasattr(core, "cores"):          # nested network: flatten with prefixed names
                for subname, sub in core.cores.items():
                    self.add_core(sub, f"{name}_{subname}")
            else:
                self.add_core(core, name)

    def add_core(self, core: TensorCore, name: str = None):
        if name in self.cores:
            raise ValueError(f"Tensor core with name {name} already exists.")
        core.name = name
        self.cores[name] = core

    def extract(self, core_names: list[str]):
        return TensorNetwork({name: self.cores[name] for name in core_names})

    def contract(self, tn=None, scheme=None, indices: list[str] = None, batch: bool = False, optimize="auto") -> TensorCore:
        """Named-index einsum over all cores (and `tn`): labels occurring once survive unless `indices` is given.
        Host utility only (see module docstring)."""
        operands = list(self.cores.values())
        if isinstance(tn, TensorNetwork):
            operands += list(tn.cores.values())
        elif isinstance(tn, TensorCore):
            operands.append(tn)
        labels = [list(op.indices) for op in operands]
        flat = [l for ls in labels for l in ls]
        if indices:
            out = list(indices)
            for idx in out:                      # unknown output labels become new unit axes on the last operand
                if idx not in flat:
                    operands[-1] = operands[-1].expand_dims(-1, name=idx, inplace=False)
                    labels[-1] = list(operands[-1].indices)
                    flat.append(idx)
        else:
            out = [l for l in flat if flat.count(l) == 1]
        if batch:
            out = list(dict.fromkeys(out))
            if "batch" not in out:
                out.append("batch")
        symbols = {}
        for l in flat + out:
            if l not in symbols:
                if len(symbols) >= len(_LETTERS):
                    raise NotImplementedError("TensorNetwork.contract: more than 52 distinct labels")
                symbols[l] = _LETTERS[len(symbols)]
        spec = ",".join("".join(symbols[l] for l in ls) for ls in labels) + "->" + "".join(symbols[l] for l in out)
        result = xp.einsum(spec, *[xp.asarray(op).view(xp.ndarray) for op in operands], optimize=True)
        return TensorCore(result, indices=out)

    def rename(self, old_name: str, new_name: str, inplace: bool = True):
        cores = self.cores if inplace else deepcopy(self.cores)
        for core in cores.values():
            core.rename(old_name, new_name, inplace=True)
        return self if inplace else TensorNetwork(cores)

    def copy(self):
        return TensorNetwork(deepcopy(self.cores))

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.cores[key]
        return self.cores[list(self.cores.keys())[key]]

    def __repr__(self):
        body = ",\n    ".join(str(core) for core in self.cores.values())
        return f"TensorNetwork(\n    {body}\n)"

    def sample(self, idx: int):
        """Slice sample `idx` out of every core when all of them carry a 'batch' axis (tensor_network.py:109-122)."""
        if not all("batch" in core.indices for core in self.cores.values()):
            return self
        picked = []
        for core in self.cores.values():
            b = core.indices.index("batch")
            arr = xp.moveaxis(core.view(xp.ndarray), b, 0)[idx, ...]
            picked.append(TensorCore(arr, indices=tuple(i for i in core.indices if i != "batch")))
        return TensorNetwork(picked, names=list(self.cores.keys()))
