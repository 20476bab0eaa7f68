"""bsde_solver — drop-in for the tensor-train regression path of antoine311200/BSDE-tensor-networks on B200.

Same import paths and call signatures as the reference package (src/bsde_solver of the reference); every numerical
routine of the path is executed by the CUDA library libbsde_b200.so through its C ABI (include/bsde_b200.h).

`xp` is kept for signature compatibility (reference: bsde_solver/__init__.py:4-29): it is the array module used for
*host glue only* (shapes, seeds, the numpy RNG stream that initialises cores exactly as the reference does).  The
reference's numpy/cupy switch is gone — there is one backend, CUDA, and no CPU implementation of the path.
"""
import importlib

xp = None


def reload_backend(backend: str):
    """Reference: bsde_solver/__init__.py:4-16.  'numpy' and 'cupy' are accepted names; both select the single
    CUDA backend (host glue arrays stay numpy).  Anything else raises ValueError as the reference does."""
    global xp
    if backend not in ("numpy", "cupy"):
        raise ValueError(f"Unknown backend {backend}")
    xp = importlib.import_module("numpy")


reload_backend("numpy")
