"""TensorCore — ndarray with named axes, the boundary type of the regression path.

Signature-compatible with the reference class (core/tensor/tensor_core.py:6-164): construction from an array plus
`indices`/`name`, `unfold`, `like`, `randomize`, `rename`, `expand_dims`, `concatenate`, `size_at`, `copy`.
It lives on the host: cores are a few hundred bytes and the (N, p) basis lists are the reference's calling convention;
the CUDA library receives their memory through the C ABI.
"""
from __future__ import annotations

import re

from bsde_solver import xp


class TensorCore(xp.ndarray):
    def __new__(cls, input_array, indices=None, name=None):
        obj = xp.asarray(input_array).view(cls)
        obj.name = name
        obj.shape_info = tuple(obj.shape)
        obj.indices = tuple(indices) if indices is not None else cls._generate_indices(obj.shape_info)
        obj.base_shape = obj.shape
        obj.base_indices = obj.indices
        # basis values produced by this package's own basis objects say so (core/calculus/basis.py: BasisValues); the tag is
        # taken over by this wrapper only — not by views, slices or results of arithmetic (see __array_finalize__)
        obj.basis_tag = getattr(input_array, "basis_tag", None) if obj.shape == getattr(input_array, "shape", None) else None
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        for attr in ("shape_info", "indices", "name", "base_shape", "base_indices"):
            setattr(self, attr, getattr(obj, attr, None))
        self.basis_tag = None

    def _describe(self):
        return ", ".join(f"{i} {{{s}}}" for i, s in zip(self.indices, self.shape_info))

    def __str__(self):
        return (f"{self.name}: " if self.name is not None else "") + f"TensorCore({self._describe()})"

    def __repr__(self):
        return f"TensorCore({self._describe()})"

    @staticmethod
    def _generate_indices(shape_info):
        return tuple(f"axis_{i}" for i in range(len(shape_info)))

    @staticmethod
    def dummy(shape_info, indices=None, name=None):
        return TensorCore(xp.zeros(shape_info), indices=indices, name=name)

    def like(input_array, like_core: "TensorCore" = None, name=None):
        """TensorCore.like(array, core): `array` reshaped to core's shape with core's indices (tensor_core.py:50-53)."""
        return TensorCore(xp.asarray(input_array).reshape(like_core.shape), indices=like_core.indices,
                          name=like_core.name if name is None else name)

    def unfold(self, *groups):
        """Merge axes: each argument is an axis (name or position), a tuple of axes to merge, or -1 for
        "all remaining axes" (tensor_core.py:55-115).  Returns the transposed + reshaped core with '+'-joined names."""
        def resolve(ax):
            return self.indices[ax] if isinstance(ax, int) else ax

        names, used, rest_at = [], [], None
        for k, g in enumerate(groups):
            if isinstance(g, tuple):
                members = [resolve(a) for a in g]
            elif g == -1:
                rest_at = k
                names.append(None)
                continue
            else:
                members = [resolve(g)]
            used += members
            names.append(members)
        if rest_at is not None:
            names[rest_at] = [i for i in self.indices if i not in used]
        order = [self.indices.index(i) for members in names for i in members]
        sizes = []
        for members in names:
            n = 1
            for i in members:
                n *= self.shape_info[self.indices.index(i)]
            sizes.append(int(n))
        out = super().transpose(*order).reshape(*sizes)
        out.indices = tuple("+".join(members) for members in names)
        out.shape_info = tuple(sizes)
        return out

    def randomize(self):
        """U(-1, 1) fill from the *global* numpy stream (tensor_core.py:117-118) — same draws as the reference."""
        self[:] = (xp.random.rand(*self.shape_info) - 0.5) * 2

    def rename(self, old_name: str, new_name: str, inplace: bool = True):
        """Regex rename of axis labels, '*' matching an index number (tensor_core.py:120-135)."""
        pattern = old_name.replace("*", r"(\d+)") if "*" in old_name else old_name
        repl = new_name.replace("*", r"\1")
        new_indices = tuple(re.sub(pattern, repl, idx) for idx in self.indices)
        if not inplace:
            return TensorCore(super().copy(), indices=new_indices, name=self.name)
        self.indices = new_indices
        return self

    def expand_dims(self, axis, inplace=True, name=None):
        label = name if name else f"axis_{axis}"
        # the reference slices with the raw axis value (tensor_core.py:137-150), so -1 inserts before the last axis
        new_shape = (*self.shape_info[:axis], 1, *self.shape_info[axis:])
        new_indices = list(self.indices)
        new_indices.insert(axis, label)
        arr = super().reshape(*new_shape)
        if inplace:
            self[:] = arr
            self.indices = tuple(new_indices)
            return self
        return TensorCore(arr, indices=tuple(new_indices), name=self.name)

    @staticmethod
    def concatenate(tensors: list["TensorCore"]):
        return TensorCore(xp.stack([t for t in tensors], axis=0), indices=("batch",) + tuple(tensors[0].indices))

    def size_at(self, axe_name):
        return self.shape_info[self.indices.index(axe_name)]

    def copy(self, *args, **kwargs):
        return TensorCore(super().copy(*args, **kwargs), indices=self.indices, name=self.name)
