"""Per-asset feature maps (reference: core/calculus/basis.py:6-44).

`eval / grad / dgrad` keep the reference's signature — a 1-D array of N points in, an (N, p) array out (an F-ordered
view, as the reference's `xp.array([...]).T`) — and are evaluated by the same device function the fused regression
kernels inline (bsde_basis_eval -> k_basis_eval).  `device_inputs(X)` is the fast path: it wraps X [d][N] already on
the GPU so that ALS / SALSA / fast_contract / multi_derivative evaluate the basis on the fly and never store it.
"""
from bsde_solver import xp
from scipy.special import legendre

from bsde_solver._lib import BASIS_LEGENDRE, BASIS_POLY


class BasisValues(xp.ndarray):
    """The (N, p) array `basis.eval(x)` returns for a host x, remembering what it is: `basis_tag` = (kind, p, coef, x) with x
    the contiguous host copy of the points.  ALS / SALSA given a list of such arrays (directly or wrapped in TensorCore)
    ship x instead of the p basis rows and evaluate the basis inside the kernels (a quarter of the PCIe traffic at p = 4, and
    the fused kernels).  The values are READ-ONLY so that the tag cannot go stale; arithmetic on them gives plain arrays
    without a tag (the tag is not inherited by views or results)."""
    basis_tag = None


class Basis:
    kind = None
    degree = 0

    def _coef(self):
        return None

    def _eval(self, x, deriv):
        from bsde_solver._backend import get_backend

        be = get_backend()
        on_device = hasattr(x, "is_cuda")
        xh = None if on_device else xp.array(x, dtype=xp.float64).ravel()       # always a private copy (it is made read-only below)
        xd = x if on_device else be.to_device(xh)
        out = be.basis_eval(self.kind, self.degree, xd, self._coef(), deriv)      # [p][N]
        if on_device:
            return out.T
        rows = out.cpu().numpy()                  # (p, N) C-order; the reference returns its transpose, an F-ordered (N, p) view
        if deriv != 0:
            return rows.T
        rows.flags.writeable = False
        xh.flags.writeable = False
        vals = rows.T.view(BasisValues)
        vals.basis_tag = (self.kind, self.degree, self._coef(), xh)
        return vals

    def eval(self, x):
        raise NotImplementedError("eval")

    def grad(self, x):
        raise NotImplementedError("grad")

    def dgrad(self, x):
        raise NotImplementedError("dgrad")

    def device_inputs(self, X):
        """X: CUDA tensor [d][N] (one time slice of the paths) -> DeviceInputs evaluated on the fly."""
        from bsde_solver._backend import DeviceInputs

        return DeviceInputs(self.kind, self.degree, X, coef=self._coef())


class PolynomialBasis(Basis):
    """Monomials x^0 .. x^(degree-1); `degree` is the number of basis functions, as in the reference."""
    kind = BASIS_POLY

    def __init__(self, degree: int):
        self.degree = degree

    def eval(self, x):
        return self._eval(x, 0)

    def grad(self, x):
        return self._eval(x, 1)

    def dgrad(self, x):
        return self._eval(x, 2)


class LegendreBasis(Basis):
    """Legendre polynomials P_0 .. P_(degree-1), evaluated as numpy.polyval does on scipy's coefficients
    (Horner from the leading coefficient), so the device values follow the reference's rounding sequence."""
    kind = BASIS_LEGENDRE

    def __init__(self, degree: int):
        self.degree = degree
        self.coefs = [legendre(i) for i in range(degree)]
        self.grad_coefs = [c.deriv() for c in self.coefs]
        table = xp.zeros((3, degree, degree))
        for i in range(degree):
            for k, poly in enumerate((self.coefs[i], self.grad_coefs[i], self.grad_coefs[i].deriv())):
                c = xp.atleast_1d(xp.asarray(poly.coeffs, dtype=xp.float64))
                table[k, i, degree - len(c):] = c          # descending powers, leading zeros are exact no-ops in Horner
        self._table = table

    def _coef(self):
        return self._table

    def eval(self, x):
        return self._eval(x, 0)

    def grad(self, x):
        return self._eval(x, 1)

    def dgrad(self, x):
        return self._eval(x, 2)
