"""BSDE models (reference: bsde.py:5-102): drift b, diffusion sigma, driver h, terminal condition g, reference price.

BlackScholes and HJB — the two models the regression path is specified on — evaluate g and h with the CUDA kernels
the device-resident time loop uses (k_terminal, k_target), for host arrays (reference signature) and CUDA tensors
alike.  `b` / `sigma` return the reference's dense host arrays for callers that want them; the path kernel itself
uses the closed diagonal form.  `price` is the reference's ground-truth diagnostic, not part of the solve.
"""
from math import ceil

from bsde_solver import xp
from bsde_solver._lib import MODEL_BLACKSCHOLES, MODEL_HJB


class BackwardSDE:
    model_id = None

    def __init__(self, X0, delta_t, T) -> None:
        self.dim = X0.shape[-1]
        self.N = int(ceil(T / delta_t))
        self.delta_t = delta_t
        self.time = xp.linspace(0, T, self.N + 1)
        self.T = T

    def params(self):
        raise NotImplementedError

    def b(self, x, t):
        pass

    def sigma(self, x, t):
        pass

    def h(self, x, t, y, z):
        pass

    def g(self, x):
        pass

    def price(self, X, t, n_sims=1000):
        pass

    # ---- device evaluation shared by the two supported models ----
    def _to_dn(self, a):
        """host (N, d) -> device [d][N]"""
        from bsde_solver._backend import get_backend

        return get_backend().to_device(xp.ascontiguousarray(xp.asarray(a, dtype=xp.float64).T))

    def _g(self, x):
        from bsde_solver._backend import get_backend

        be = get_backend()
        if hasattr(x, "is_cuda"):
            return be.terminal(self.model_id, self.params(), x)
        return be.terminal(self.model_id, self.params(), self._to_dn(x)).cpu().numpy()

    def _h(self, x, y, z):
        from bsde_solver._backend import get_backend

        be = get_backend()
        on_device = hasattr(x, "is_cuda")
        if not on_device:
            x, z = self._to_dn(x), self._to_dn(z)
            y = be.to_device(xp.asarray(y, dtype=xp.float64).ravel())
        zero = be.empty(x.shape[1]).zero_()
        out = be.target(self.model_id, self.params(), x, y, zero, z, None, 1.0, False)   # h*1 + 0 == h exactly
        return out if on_device else out.cpu().numpy()


class BlackScholes(BackwardSDE):
    model_id = MODEL_BLACKSCHOLES

    def __init__(self, X0, delta_t, T, r, sigma) -> None:
        super().__init__(X0, delta_t, T)
        self.r = r
        self.sigma_ = sigma

    def params(self):
        return xp.array([self.r, self.sigma_], dtype=xp.float64)

    def b(self, x, t):
        return xp.zeros_like(x)

    def sigma(self, x, t):
        out = xp.zeros((x.shape[0], x.shape[1], x.shape[1]))
        idx = xp.arange(x.shape[1])
        out[:, idx, idx] = self.sigma_ * x
        return out

    def h(self, x, t, y, z):
        return self._h(x, y, z)

    def g(self, x):
        return self._g(x)

    def price(self, X, t, n_sims=1000):
        from bsde_solver._backend import get_backend

        gx = self._g(X if hasattr(X, "is_cuda") else self._to_dn(X))
        return xp.exp((self.r + self.sigma_**2) * (self.T - t)) * get_backend().mean(gx)


class HJB(BackwardSDE):
    model_id = MODEL_HJB

    def __init__(self, X0, delta_t, T, sigma) -> None:
        super().__init__(X0, delta_t, T)
        self.sigma_ = sigma

    def params(self):
        return xp.array([self.sigma_], dtype=xp.float64)

    def b(self, x, t):
        return xp.zeros_like(x)

    def sigma(self, x, t):
        return self.sigma_ * xp.broadcast_to(xp.eye(self.dim), (x.shape[0], self.dim, self.dim))

    def h(self, x, t, y, z):
        return self._h(x, y, z)

    def g(self, x):
        return self._g(x)

    def price(self, x, t, n_sims=1000):
        """-log E[exp(-g(x + sqrt(T-t) sigma xi))] by Monte Carlo over n_sims draws per sample (bsde.py:98-102).
        Ground truth for the printed error only; evaluated on the host in sample chunks."""
        x = xp.asarray(x.T.cpu().numpy() if hasattr(x, "is_cuda") else x, dtype=xp.float64)
        out = xp.empty(x.shape[0])
        step = max(1, int(2**22 // max(1, x.shape[1] * n_sims)))
        for s in range(0, x.shape[0], step):
            xs = x[s:s + step]
            noise = xp.random.randn(xs.shape[0], xs.shape[1], n_sims)
            pts = xs[:, :, None] + xp.sqrt(self.T - t) * self.sigma_ * noise
            value = 1.0 / (0.5 + 0.5 * xp.sum(pts**2, axis=1))        # exp(-log(.)) = 1/(.)
            out[s:s + step] = -xp.log(xp.mean(value, axis=-1))
        return out
