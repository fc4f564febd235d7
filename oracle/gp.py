"""CPU oracle: george.GP + george.kernels semantics as alabi uses them.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**:
george (dfm/george 0.4.x, unpinned in /root/reference/setup.py:15) is not
vendored and not installed; this file restates its published algorithm
(BasicSolver = dense kernel matrix + LAPACK Cholesky) and is anchored on the
reference's call sites:

* kernel construction         alabi/core.py:998-1014
* amplitude ``kernel * var``  alabi/gp_utils.py:230-231, core.py:1136-1139
* ``george.GP(...)`` ctor     alabi/gp_utils.py:233, core.py:1141
* ``gp.compute``              alabi/gp_utils.py:243, core.py:1158,1430
* ``gp.log_likelihood``       alabi/core.py:1248
* ``gp.grad_log_likelihood``  alabi/core.py:1261
* ``gp.predict``              alabi/core.py:85,95,1441,1601,1812
* parameter order / names     docs/source/save_reload.py:117-120,
                              alabi/core.py:1050-1076

Everything is FP64, row-major, and deliberately written for clarity, not
speed (N^2 P memory in ``grad_log_likelihood`` exactly like george).
"""
import numpy as np
from scipy.linalg import cholesky, cho_solve, solve_triangular

KERNEL_IDS = {"ExpSquaredKernel": 0, "Matern32Kernel": 1, "Matern52Kernel": 2}


# ---------------------------------------------------------------------------
# stationary kernels with an axis-aligned metric (george.kernels, recalled)
# ---------------------------------------------------------------------------
def scaled_sqdist(x1, x2, log_M):
    """r^2_ij = sum_k (x1_ik - x2_jk)^2 / M_k with M_k = exp(log_M_k).

    george's axis-aligned ``Metric`` stores the *squared* length scale M
    (alabi/core.py:987-993 passes ``metric=exp(log l)`` as M)."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
    x2 = np.atleast_2d(np.asarray(x2, dtype=np.float64))
    inv_M = np.exp(-np.asarray(log_M, dtype=np.float64))
    r2 = np.zeros((x1.shape[0], x2.shape[0]))
    for k in range(x1.shape[1]):            # per dimension: no (M, N, d) temporary
        dk = x1[:, k, None] - x2[None, :, k]
        r2 += dk * dk * inv_M[k]
    return r2


def radial(kind, r2):
    """k(r^2) for the three kernels north_star names (george kernel defs)."""
    kind = KERNEL_IDS.get(kind, kind)
    if kind == 0:                       # ExpSquaredKernel
        return np.exp(-0.5 * r2)
    if kind == 1:                       # Matern32Kernel
        r = np.sqrt(3.0 * r2)
        return (1.0 + r) * np.exp(-r)
    if kind == 2:                       # Matern52Kernel
        r = np.sqrt(5.0 * r2)
        return (1.0 + r + r * r / 3.0) * np.exp(-r)
    raise ValueError(f"unknown kernel {kind}")


def radial_grad(kind, r2):
    """dk/d(r^2) (george's ``radial_gradient``)."""
    kind = KERNEL_IDS.get(kind, kind)
    if kind == 0:
        return -0.5 * np.exp(-0.5 * r2)
    if kind == 1:
        r = np.sqrt(3.0 * r2)
        return -3.0 * 0.5 * np.exp(-r)
    if kind == 2:
        r = np.sqrt(5.0 * r2)
        return -5.0 * (1.0 + r) * np.exp(-r) / 6.0
    raise ValueError(f"unknown kernel {kind}")


def kernel_value(kind, x1, x2, log_M, log_const=None):
    """amp * k(x1, x2); amp = exp(log_const) (ConstantKernel product) or 1."""
    k = radial(kind, scaled_sqdist(x1, x2, log_M))
    if log_const is not None:
        k = np.exp(log_const) * k
    return k


# ---------------------------------------------------------------------------
# GP object with the george protocol alabi consumes
# ---------------------------------------------------------------------------
class OracleGP:
    """Restatement of ``george.GP(kernel, fit_mean, mean, white_noise,
    fit_white_noise)`` with a ``BasicSolver``.

    Hyper-parameter vector order (unfrozen entries only):
    ``[mean:value][white_noise:value][kernel:k1:log_constant]
    kernel:k2:metric:log_M_0_0 ...`` (docs/source/save_reload.py:117-120).
    Without an amplitude the scale names are ``kernel:metric:log_M_i_i``.
    """

    def __init__(self, kind, ndim, log_M, log_const=None, mean=0.0, fit_mean=False,
                 white_noise=-12.0, fit_white_noise=False):
        self.kind = KERNEL_IDS.get(kind, kind)
        self.ndim = int(ndim)
        self.log_M = np.array(log_M, dtype=np.float64).reshape(self.ndim)
        self.log_const = None if log_const is None else float(log_const)
        self.mean = float(mean)
        self.white_noise = float(white_noise)
        self.fit_mean = bool(fit_mean)
        self.fit_white_noise = bool(fit_white_noise)
        self.computed = False
        self._x = None
        self._alpha = None
        self._y = None

    # -- parameter protocol -------------------------------------------------
    def get_parameter_names(self, include_frozen=False):
        names = []
        if self.fit_mean or include_frozen:
            names.append("mean:value")
        if self.fit_white_noise or include_frozen:
            names.append("white_noise:value")
        if self.log_const is not None:
            names.append("kernel:k1:log_constant")
            pre = "kernel:k2:metric:"
        else:
            pre = "kernel:metric:"
        names += [f"{pre}log_M_{i}_{i}" for i in range(self.ndim)]
        return tuple(names)

    def get_parameter_vector(self):
        v = []
        if self.fit_mean:
            v.append(self.mean)
        if self.fit_white_noise:
            v.append(self.white_noise)
        if self.log_const is not None:
            v.append(self.log_const)
        return np.array(v + list(self.log_M), dtype=np.float64)

    def set_parameter_vector(self, v):
        v = np.asarray(v, dtype=np.float64)
        if len(v) != len(self.get_parameter_vector()):
            raise ValueError("dimension mismatch")
        n = 0
        if self.fit_mean:
            self.mean = float(v[n]); n += 1
        if self.fit_white_noise:
            self.white_noise = float(v[n]); n += 1
        if self.log_const is not None:
            self.log_const = float(v[n]); n += 1
        self.log_M = v[n:n + self.ndim].copy()
        self.computed = False            # george marks the model dirty
        self._alpha = None

    def get_parameter_dict(self):
        return dict(zip(self.get_parameter_names(), self.get_parameter_vector()))

    # -- kernel matrix + factorisation (BasicSolver.compute) -----------------
    def get_matrix(self, x1, x2=None):
        x2 = x1 if x2 is None else x2
        return kernel_value(self.kind, x1, x2, self.log_M, self.log_const)

    def compute(self, x, yerr=0.0):
        """K = amp k(X,X) + (yerr^2 + exp(white_noise)) I; upper Cholesky;
        logdet = 2 sum log diag; raises ``np.linalg.LinAlgError`` if not SPD."""
        self._x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        n = len(self._x)
        self._yerr2 = float(yerr) ** 2 * np.ones(n)
        K = self.get_matrix(self._x)
        K[np.diag_indices_from(K)] += self._yerr2 + np.exp(self.white_noise)
        self._factor = (cholesky(K, overwrite_a=True, lower=False), False)
        self.log_determinant = 2.0 * np.sum(np.log(np.diag(self._factor[0])))
        self._const = -0.5 * (n * np.log(2.0 * np.pi) + self.log_determinant)
        self.computed = True
        self._alpha = None
        return self

    def recompute(self, quiet=False):
        if not self.computed:
            if self._x is None:
                raise RuntimeError("You need to compute the model first")
            try:
                self.compute(self._x, np.sqrt(self._yerr2[0]))
            except (ValueError, np.linalg.LinAlgError):
                if quiet:
                    return False
                raise
        return True

    def apply_inverse(self, b):
        return cho_solve(self._factor, b)

    def get_inverse(self):
        return self.apply_inverse(np.eye(len(self._x)))

    def _compute_alpha(self, y):
        r = np.ascontiguousarray(np.asarray(y, dtype=np.float64) - self.mean)
        self._y = np.asarray(y, dtype=np.float64)
        self._alpha = self.apply_inverse(r)
        return self._alpha

    # -- a3 / a4 / a6 ---------------------------------------------------------
    def log_likelihood(self, y, quiet=False):
        if not self.recompute(quiet=quiet):
            return -np.inf
        r = np.asarray(y, dtype=np.float64) - self.mean
        ll = self._const - 0.5 * np.dot(r, self.apply_inverse(r))
        return ll if np.isfinite(ll) else -np.inf

    def grad_log_likelihood(self, y, quiet=False):
        npar = len(self.get_parameter_vector())
        if not self.recompute(quiet=quiet):
            return np.zeros(npar)
        alpha = self._compute_alpha(y)
        A = np.outer(alpha, alpha) - self.get_inverse()
        g = []
        if self.fit_mean:
            g.append(np.sum(alpha))
        if self.fit_white_noise:
            g.append(0.5 * np.exp(self.white_noise) * np.trace(A))
        x = self._x
        r2 = scaled_sqdist(x, x, self.log_M)
        amp = 1.0 if self.log_const is None else np.exp(self.log_const)
        if self.log_const is not None:
            g.append(0.5 * np.sum(A * (amp * radial(self.kind, r2))))
        dk = amp * radial_grad(self.kind, r2)
        inv_M = np.exp(-self.log_M)
        for k in range(self.ndim):
            d2 = (x[:, None, k] - x[None, :, k]) ** 2 * inv_M[k]
            g.append(0.5 * np.sum(A * dk * (-d2)))
        return np.array(g, dtype=np.float64)

    def predict(self, y, t, return_cov=False, return_var=False):
        self.recompute()
        alpha = self._compute_alpha(y)
        xs = np.atleast_2d(np.asarray(t, dtype=np.float64))
        Kxs = self.get_matrix(xs, self._x)
        mu = Kxs @ alpha + self.mean
        if not (return_var or return_cov):
            return mu
        KinvKxs = self.apply_inverse(Kxs.T)
        if return_var:
            amp = 1.0 if self.log_const is None else np.exp(self.log_const)
            var = amp * np.ones(len(xs))          # k(x*, x*) ; white noise NOT added
            var -= np.sum(Kxs.T * KinvKxs, axis=0)
            return mu, var
        cov = self.get_matrix(xs) - Kxs @ KinvKxs
        return mu, cov

    # -- gradients of the prediction w.r.t. the query point -----------------------------
    def predict_grad(self, y, t, h=None):
        """d mu / dx and d sigma^2 / dx (M x d) as alabi/utility.py:558-621 defines
        them: ``grad_k^T alpha`` and ``-2 grad_k^T K^-1 k*``.  ``h=None`` uses the
        analytic kernel derivative dk/dx_k = amp k'(r^2) 2 (x_k - x_jk) / M_k; a float
        reproduces the reference's central difference of the kernel
        (``numerical_kernel_gradient``, h = 1e-6 there)."""
        self.recompute()
        alpha = self._compute_alpha(y)
        xs = np.atleast_2d(np.asarray(t, dtype=np.float64))
        amp = 1.0 if self.log_const is None else np.exp(self.log_const)
        inv_M = np.exp(-self.log_M)
        dmu = np.zeros_like(xs)
        dvar = np.zeros_like(xs)
        for q in range(len(xs)):
            x = xs[q]
            if h is None:
                diff = x[None, :] - self._x                                   # (N, d)
                r2 = np.sum(diff * diff * inv_M[None, :], axis=1)
                gk = (amp * radial_grad(self.kind, r2))[:, None] * 2.0 * diff * inv_M[None, :]
            else:
                gk = np.zeros_like(self._x)
                for i in range(self.ndim):
                    xp, xm = x.copy(), x.copy()
                    xp[i] += h
                    xm[i] -= h
                    gk[:, i] = (self.get_matrix(xp[None, :], self._x).ravel()
                                - self.get_matrix(xm[None, :], self._x).ravel()) / (2 * h)
            ks = self.get_matrix(x[None, :], self._x).ravel()
            dmu[q] = gk.T @ alpha
            dvar[q] = -2.0 * gk.T @ self.apply_inverse(ks)
        return dmu, dvar

    # extended-precision style variance used to measure conditioning effects
    def predict_var_via_L(self, t):
        """sigma^2 = k** - ||L^{-1} k*||^2 (the form the GPU path uses)."""
        xs = np.atleast_2d(np.asarray(t, dtype=np.float64))
        Kxs = self.get_matrix(xs, self._x)
        v = solve_triangular(self._factor[0], Kxs.T, trans="T", lower=False)
        amp = 1.0 if self.log_const is None else np.exp(self.log_const)
        return amp - np.sum(v * v, axis=0)


def make_gp(kind, x, y, log_M, amp=None, mean=None, white_noise=-12.0,
            fit_mean=True, fit_white_noise=True, compute=True):
    """``gp_utils.configure_gp`` composition (alabi/gp_utils.py:222-248):
    ``kernel *= var(y)`` builds ``ConstantKernel(log(var/ndim)) * k`` (george
    ``Kernel.__mul__`` divides a float by ndim — recalled), ``mean=median(y)``."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    ndim = x.shape[1]
    log_const = None if amp is None else np.log(float(amp) / ndim)
    gp = OracleGP(kind, ndim, log_M, log_const=log_const,
                  mean=np.median(y) if mean is None else mean, fit_mean=fit_mean,
                  white_noise=white_noise, fit_white_noise=fit_white_noise)
    if compute:
        gp.compute(x)
    return gp
