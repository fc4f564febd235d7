"""Benchmark likelihoods with the names of alabi/benchmarks.py (input generators
for BASELINE.json's configs): ``rosenbrock``, ``gaussian_shells``, ``eggbox``,
``gaussian_2d`` dicts with ``fn`` / ``bounds`` and ``random_gaussian_covariance``."""
import math

import numpy as np
from scipy.optimize import rosen
from scipy.stats import multivariate_normal

__all__ = ["rosenbrock", "gaussian_shells", "eggbox", "gaussian_2d", "multimodal", "random_gaussian_covariance",
           "gaussian_nd"]


def rosenbrock_fn(x):
    return -rosen(x) / 100.0


rosenbrock = {"fn": rosenbrock_fn, "bounds": [(-5, 5), (-5, 5)]}


def _logcirc(theta, c, r=2.0, w=0.1):
    const = math.log(1.0 / math.sqrt(2.0 * math.pi * w ** 2))
    d = np.sqrt(np.sum((theta - c) ** 2, axis=-1))
    return const - (d - r) ** 2 / (2.0 * w ** 2)


def gaussian_shells_fn(theta):
    theta = np.asarray(theta).flatten()
    return np.logaddexp(_logcirc(theta, np.array([-3.5, 0.0])), _logcirc(theta, np.array([3.5, 0.0])))


gaussian_shells = {"fn": gaussian_shells_fn, "bounds": [(-6, 6), (-6, 6)]}


def eggbox_fn(x):
    x = np.asarray(x).flatten()
    tmax = 5.0 * np.pi
    t = 2.0 * tmax * x - tmax
    return -(2.0 + np.cos(t[0] / 2.0) * np.cos(t[1] / 2.0)) ** 5.0


eggbox = {"fn": eggbox_fn, "bounds": [(0, 1), (0, 1)]}


def multimodal_fn(x):
    x = np.asarray(x).flatten()
    return -(np.sin(x[0]) ** 10 + np.cos(10 + x[1] * x[0]) * np.cos(x[0]))


multimodal = {"fn": multimodal_fn, "bounds": [(0, 5), (0, 5)]}


def gaussian_2d_fn(theta):
    theta = np.asarray(theta).flatten()
    return multivariate_normal.logpdf(theta, mean=np.array([0.5, 0.5]), cov=np.array([[0.1, 0.0], [0.0, 0.1]]))


gaussian_2d = {"fn": gaussian_2d_fn, "bounds": [(0, 1), (0, 1)]}


def random_gaussian_covariance(n_dims, rng=None):
    """Q diag(lambda) Q^T, lambda ~ Exp(1), Q from the QR of a normal matrix."""
    rng = np.random if rng is None else rng
    lam = rng.exponential(scale=1.0, size=n_dims)
    Q = rng.standard_normal((n_dims, n_dims)) if hasattr(rng, "standard_normal") else rng.randn(n_dims, n_dims)
    Q, _ = np.linalg.qr(Q)
    return Q @ np.diag(lam) @ Q.T


def gaussian_nd(ndim, rng=None, bound=3.0):
    """Correlated N-d Gaussian log-density on (-bound, bound)^ndim (docs/source/plot_gaussian_nd.py)."""
    cov = random_gaussian_covariance(ndim, rng)
    mean = np.zeros(ndim)

    def fn(theta):
        return multivariate_normal.logpdf(np.asarray(theta).flatten(), mean=mean, cov=cov)
    return {"fn": fn, "bounds": [(-bound, bound)] * ndim, "cov": cov}
