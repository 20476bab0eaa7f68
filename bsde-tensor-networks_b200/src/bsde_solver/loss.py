"""PDELoss — residual of the semilinear PDE the BSDE represents, evaluated along the paths as an accuracy diagnostic that
needs no reference solution (reference: loss.py:6-33; there it is only called from commented-out code, pipeline.py:155-161):

    d_t v + 1/2 Tr(sigma sigma^T D^2 v) + b . D v + h(x, t, v, D v) = 0

The reference's body multiplies `sigma.T @ sigma` of an (N, d, d) array (which does not broadcast) and prints its shape; this
class implements the equation of its docstring with the same call signature."""
from bsde_solver import xp


class PDELoss:
    def __init__(self, model):
        self.model = model

    def __call__(self, t, x, v, vt, vx, vxx):
        """t: time; x (N, d); v (N,); vt (N,) time derivative; vx (N, d) gradient; vxx (N, d, d) Hessian.  Returns (N,)."""
        x = xp.asarray(x, dtype=xp.float64)
        vx = xp.asarray(vx, dtype=xp.float64)
        vxx = xp.asarray(vxx, dtype=xp.float64)
        sigma = xp.asarray(self.model.sigma(x, t))                     # (N, d, d)
        a = xp.einsum("nik,njk->nij", sigma, sigma)                    # sigma sigma^T
        loss = xp.asarray(vt, dtype=xp.float64) + xp.sum(xp.asarray(self.model.b(x, t)) * vx, axis=1)
        loss = loss + 0.5 * xp.sum(a * vxx, axis=(1, 2))
        return loss + xp.asarray(self.model.h(x, t, xp.asarray(v, dtype=xp.float64), vx))


class ReferenceLoss:
    pass
