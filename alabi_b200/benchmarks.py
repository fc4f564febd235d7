"""Benchmark likelihoods with the names of alabi/benchmarks.py (input generators
for BASELINE.json's configs): ``rosenbrock``, ``gaussian_shells``, ``eggbox``,
``gaussian_2d`` dicts with ``fn`` / ``bounds`` and ``random_gaussian_covariance``."""
import math

import numpy as np
from scipy.optimize import rosen
from scipy.stats import multivariate_normal

__all__ = ["test1d", "rosenbrock", "rosenbrock_nd", "gaussian_shells", "eggbox", "gaussian_2d", "multimodal",
           "random_gaussian_covariance", "multimodal_gaussian_nd", "gaussian_nd"]
# (the reference's ``logo`` benchmark reads ../benchmark/logo.txt through scipy's removed
# interp2d; it is a plotting demo, not an input generator of BASELINE.json's configs)


def test1d_fn(theta):
    """1-D Bayesian-optimisation toy (alabi/benchmarks.py:27-34)."""
    theta = np.asarray(theta)
    return -np.sin(3 * theta) - theta ** 2 + 0.7 * theta


test1d = {"fn": test1d_fn, "bounds": [(-2, 1)]}


def rosenbrock_fn(x):
    return -rosen(x) / 100.0


rosenbrock = {"fn": rosenbrock_fn, "bounds": [(-5, 5), (-5, 5)]}


def rosenbrock_nd(x, a, b):
    """N-d Rosenbrock log-density of Pagani et al. 2020 as the reference evaluates it
    (alabi/benchmarks.py:59-94): ``b`` is an (n1, n2) coefficient matrix, ndim = (n1-1) n2 + 1;
    ``x`` is (ndim,) or (nsamples, ndim)."""
    x = np.asarray(x, dtype=float)
    n1, n2 = b.shape
    ndim = (n1 - 1) * n2 + 1
    one = x.ndim == 1
    xs = x.reshape(1, -1) if one else x
    col_weight = b[:, 2:].sum(axis=0)
    ll = -a * (xs[:, 0] - 1) ** 2
    ll = ll - (((xs[:, 2:n1] - xs[:, 1:n1 - 1] ** 2) ** 2) * col_weight).sum(axis=1)
    norm = np.sqrt(a / np.pi) * np.pi ** ndim * np.sqrt(np.prod(b[:, 2:]))
    ll = ll - np.log(norm)
    return ll[0] if one else ll


def logcirc(theta, c):
    """One Gaussian shell, radius 2, width 0.1 (alabi/benchmarks.py:100-105)."""
    return _logcirc(theta, c)


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


def multimodal_gaussian_nd(x, means, covs, amps):
    """alabi/benchmarks.py:209-215, kept as the reference computes it: exp of the summed
    exponentials of the amplitude-scaled mode log-densities."""
    parts = np.array([amps[i] * multivariate_normal.logpdf(x, mean=means[i], cov=covs[i]) for i in range(len(means))])
    return np.exp(np.sum(np.exp(parts), axis=0))


def gaussian_nd(ndim, rng=None, bound=3.0):
    """Correlated N-d Gaussian log-density on (-bound, bound)^ndim (docs/source/plot_gaussian_nd.py)."""
    cov = random_gaussian_covariance(ndim, rng)
    mean = np.zeros(ndim)

    def fn(theta):
        return multivariate_normal.logpdf(np.asarray(theta).flatten(), mean=mean, cov=cov)
    return {"fn": fn, "bounds": [(-bound, bound)] * ndim, "cov": cov}
