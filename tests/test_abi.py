"""CPU: the C-ABI library builds, loads without a GPU, exports every symbol include/bsde_b200.h declares, and the
ctypes binding covers the same set.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bsde_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"BSDE_API\s+[\w\s\*]+?\b(bsde_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("bsde_fit_als", "bsde_fit_als_host", "bsde_fit_salsa", "bsde_tt_eval", "bsde_tt_grad", "bsde_paths_simulate",
                 "bsde_target", "bsde_normal_eq", "bsde_solve", "bsde_orthonormalize", "bsde_basis_eval", "bsde_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from bsde_solver import _lib

    lib = _lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in declared_symbols():
        assert name in exported, f"{name} declared in the header but not exported"
        assert hasattr(lib, name)
    # nothing but the C ABI leaks out of the library
    assert all(n.startswith("bsde_") for n in exported), sorted(n for n in exported if not n.startswith("bsde_"))[:5]


def test_ctypes_binding_matches_header():
    from bsde_solver import _lib

    bound = set(_lib.SIGNATURES) | {"bsde_last_error"}
    assert bound == set(declared_symbols())


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product path refuses to run — there is no CPU fallback to fall into."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bsde_solver import _lib
    from bsde_solver._backend import Backend, BackendError

    lib = _lib.load()
    ctx = ctypes.c_void_p()
    rc = lib.bsde_ctx_create(0, ctypes.byref(ctx))
    assert rc != 0 and b"CUDA" in lib.bsde_last_error()
    with pytest.raises(BackendError):
        Backend()
    from bsde_solver.core.optimisation.als import ALS
    import numpy as np

    phis = [np.ones((10, 3)) for _ in range(3)]
    with pytest.raises(BackendError):
        ALS(phis, np.ones(10), ranks=(1, 2, 2, 1))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "bsde-tensor-networks_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "tt_oracle" not in text, os.path.join(base, f)
