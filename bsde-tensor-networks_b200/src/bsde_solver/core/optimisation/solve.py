"""solver(A, b, method) — the small dense solve of the normal equations (reference: core/optimisation/solve.py:6-19),
executed on the GPU by the same in-CTA routines the ALS micro-step uses (bsde_solve -> k_micro_solve):
'lstsq' = minimum-norm least squares with numpy's rcond=None cutoff (eps * n * s_max), 'lu' / 'qr' / 'solve' =
Gaussian elimination with partial pivoting."""
from bsde_solver import xp
from bsde_solver._lib import SOLVER_IDS


def solver_id(method):
    if method not in SOLVER_IDS:
        raise ValueError(f"Unknown method {method}")
    return SOLVER_IDS[method]


def solver(A, b, method="lu"):
    from bsde_solver._backend import get_backend

    sid = solver_id(method)
    x, _ = get_backend().solve(xp.asarray(A).view(xp.ndarray), xp.asarray(b).view(xp.ndarray), sid)
    return x
