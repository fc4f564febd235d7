"""CPU oracle: acquisition utilities, priors and the length-scale regulariser.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED against golden
vectors produced by the reference's own functions
(``tests/golden/make_golden.py`` -> ``tests/golden/utility_golden.npz``).

Vectorised restatements (one value per candidate row) of:

* ``lnprior_uniform``           alabi/utility.py:218-275  (strict inequalities)
* ``prior_transform_uniform``   alabi/utility.py:278-345
* ``lnprior_normal``            alabi/utility.py:370-378  (uniform box + independent normals)
* ``prior_transform_normal``    alabi/utility.py:381-486  (untruncated inverse normal CDF)
* ``logsubexp``                 alabi/utility.py:489-504
* ``agp_utility``               alabi/utility.py:629-701
* ``bape_utility``              alabi/utility.py:729-810
* ``jones_utility``             alabi/utility.py:853-946
* ``regularization_term/gradient``  alabi/gp_utils.py:30-108
* ``estimate_burnin``           alabi/mcmc_utils.py:15-72
"""
import numpy as np
from scipy.special import ndtr

UTILITY_IDS = {"bape": 0, "agp": 1, "jones": 2}


def in_bounds(theta, bounds):
    """True where every coordinate is STRICTLY inside its (lo, hi) pair."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    b = np.asarray(bounds, dtype=np.float64)
    return np.all((theta > b[:, 0]) & (theta < b[:, 1]), axis=1)


def lnprior_uniform(theta, bounds):
    """0 inside the open box, -inf outside; one value per row."""
    return np.where(in_bounds(theta, bounds), 0.0, -np.inf)


def prior_transform_uniform(u, bounds):
    u = np.asarray(u, dtype=np.float64)
    b = np.asarray(bounds, dtype=np.float64)
    return (b[:, 1] - b[:, 0]) * u + b[:, 0]


LOG_SQRT_2PI = 0.9189385332046727      # scipy.stats.norm's _norm_pdf_logC = log(sqrt(2 pi))


def lnprior_normal(theta, bounds, mu, sd):
    """Box prior plus independent normals; ``mu[k]`` NaN marks a uniform dimension.  The
    normal terms are added in dimension order with scipy's ``norm.logpdf`` expression
    ((-z*z/2 - log sqrt(2 pi)) - log sd, z = (x - mu)/sd), one value per row."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    lnp = lnprior_uniform(theta, bounds).copy()
    for k, (m, s) in enumerate(zip(mu, sd)):
        if not np.isnan(m):
            z = (theta[:, k] - m) / s
            lnp = lnp + ((-(z * z) / 2.0 - LOG_SQRT_2PI) - np.log(s))
    return lnp


def prior_transform_normal(u, bounds, mu, sd):
    from scipy.special import ndtri
    u = np.asarray(u, dtype=np.float64)
    b = np.asarray(bounds, dtype=np.float64)
    out = (b[:, 1] - b[:, 0]) * u + b[:, 0]
    for k, (m, s) in enumerate(zip(mu, sd)):
        if not np.isnan(m):
            out[..., k] = ndtri(u[..., k]) * s + m
    return out


def logsubexp(x1, x2):
    """log(exp(x1) - exp(x2)) = x1 + log(1 - exp(x2 - x1)); -inf when x1 <= x2."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    with np.errstate(all="ignore"):
        out = x1 + np.log(1.0 - np.exp(x2 - x1))
    return np.where(x1 <= x2, -np.inf, out)


def bape(mu, var, inside=None):
    """-((2 mu + var) + logsubexp(var, 0)); +inf outside the prior box."""
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    with np.errstate(all="ignore"):
        u = -((2.0 * mu + var) + logsubexp(var, 0.0))
    return u if inside is None else np.where(inside, u, np.inf)


def agp(mu, var, inside=None):
    """-(mu + 0.5 log(2 pi e var)); NaN for var < 0; +inf outside the box."""
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    with np.errstate(all="ignore"):
        u = -(mu + 0.5 * np.log(2.0 * np.pi * np.e * var))
    return u if inside is None else np.where(inside, u, np.inf)


def jones(mu, var, y_best, zeta=0.01, inside=None):
    """-((mu - y_best - zeta) Phi(z) + sigma phi(z)), z = (mu-y_best-zeta)/sigma;
    exactly 0.0 when sigma is not > 0 (this includes NaN sigma from var < 0)."""
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    with np.errstate(all="ignore"):
        std = np.sqrt(var)
        d = mu - y_best - zeta
        z = d / std
        pdf = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
        u = -(d * ndtr(z) + std * pdf)
    u = np.where(std > 0, u, 0.0)
    return u if inside is None else np.where(inside, u, np.inf)


def utility(kind, mu, var, inside=None, y_best=0.0, zeta=0.01):
    kind = UTILITY_IDS.get(kind, kind)
    if kind == 0:
        return bape(mu, var, inside)
    if kind == 1:
        return agp(mu, var, inside)
    if kind == 2:
        return jones(mu, var, y_best, zeta, inside)
    raise ValueError(kind)


def first_argmin(u):
    """Lowest index of the smallest non-NaN value (np.nanargmin tie rule);
    -1 if every entry is NaN."""
    u = np.asarray(u, dtype=np.float64)
    if np.all(np.isnan(u)):
        return -1
    return int(np.nanargmin(u))


def regularization_term(hp, length_idx, amp_0=1.0, mu_0=1.0, sigma_0=2.0):
    """LogNormal(mu_0 + 0.5 log(len(hp)), sigma_0) negative log prior summed
    over the length-scale entries of the hyper-vector (the reference uses
    ``len(hparams)``, not the problem dimension — reproduced)."""
    hp = np.asarray(hp, dtype=np.float64)
    ll = hp[list(length_idx)]
    mu = mu_0 + 0.5 * np.log(len(hp))
    return amp_0 * np.sum(ll + 0.5 * np.log(2.0 * np.pi * sigma_0 ** 2)
                          + (ll - mu) ** 2 / (2.0 * sigma_0 ** 2))


def regularization_gradient(hp, length_idx, amp_0=1.0, mu_0=1.0, sigma_0=2.0):
    """Reference's "gradient": (1 + (l - mu)/sigma_0^2) / exp(l) on the
    length-scale slots (d/d l_linear, not d/d log l — reproduced)."""
    hp = np.asarray(hp, dtype=np.float64)
    idx = list(length_idx)
    g = np.zeros_like(hp)
    mu = mu_0 + 0.5 * np.log(len(hp))
    g[idx] = (1.0 + (hp[idx] - mu) / sigma_0 ** 2) / np.exp(hp[idx])
    return amp_0 * g


def burnin_thin(tau):
    """burn = int(2 max tau), thin = max(int(0.5 min tau), 1); NaNs dropped,
    tau = 1 if nothing finite is left."""
    tau = np.atleast_1d(np.asarray(tau, dtype=np.float64))
    if np.any(~np.isfinite(tau)):
        tau = tau[np.isfinite(tau)]
        if len(tau) < 1:
            tau = np.array([1.0])
    return int(2.0 * np.max(tau)), int(max(int(0.5 * np.min(tau)), 1))
