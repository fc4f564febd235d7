"""Synthetic inputs of BASELINE.json's five configurations: the training set a
``SurrogateModel`` would hold, the kernel north_star names for it, and the hyper-parameters the
timings and the parity tests use.  Input generators only — no arithmetic of the hot path.

    c1  2-D Rosenbrock      ExpSquared  N = 150   (50 initial + 100 active-learning points)
    c2  2-D Gaussian shells Matern-3/2  N = 1000
    c3  2-D eggbox          Matern-5/2  N = 4000
    c4  10-D Gaussian       ExpSquared  N = 8192
    c5  20-D Gaussian       ExpSquared  N = 16384

The likelihoods restate alabi/benchmarks.py:46-52 (Rosenbrock), :100-116 (shells), :123-132
(eggbox), :195-206 (random covariance) and docs/source/plot_gaussian_nd.py (correlated N-d
Gaussian on (-3, 3)^d), vectorised over rows.  Seeds are fixed (seed = config number) and the
draws use ``numpy.random.default_rng`` only, so every process / rank / test rebuilds the same
arrays bit for bit (``tests/test_host_logic.py`` pins them against the oracle's generator).
Hyper-parameters: mean = median(y), amplitude = var(y) (george stores log(var / ndim)),
white noise = the reference default -12 (alabi/core.py:741), log_M per config below.
"""
import numpy as np

__all__ = ["make_config", "CONFIGS"]

CONFIGS = {
    "c1": dict(kind="ExpSquaredKernel", ndim=2, lo=-5.0, hi=5.0, n=150, fn="rosenbrock", utility="bape", seed=1,
               log_M=float(np.log(1.5 ** 2))),
    "c2": dict(kind="Matern32Kernel", ndim=2, lo=-6.0, hi=6.0, n=1000, fn="gaussian_shells", utility="agp", seed=2,
               log_M=float(np.log(0.5 ** 2))),
    "c3": dict(kind="Matern52Kernel", ndim=2, lo=0.0, hi=1.0, n=4000, fn="eggbox", utility="bape", seed=3,
               log_M=float(np.log(0.1 ** 2))),
    "c4": dict(kind="ExpSquaredKernel", ndim=10, lo=-3.0, hi=3.0, n=8192, fn="gaussian_nd", utility="bape", seed=4,
               log_M=2.0),
    "c5": dict(kind="ExpSquaredKernel", ndim=20, lo=-3.0, hi=3.0, n=16384, fn="gaussian_nd", utility="bape", seed=5,
               log_M=2.0),
}


def _rosenbrock(x):
    r = np.sum(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2, axis=1)
    return -r / 100.0


def _gaussian_shells(x, r=2.0, w=0.1, c=3.5):
    const = np.log(1.0 / np.sqrt(2.0 * np.pi * w ** 2))

    def shell(cx):
        dist = np.sqrt((x[:, 0] - cx) ** 2 + x[:, 1] ** 2)
        return const - (dist - r) ** 2 / (2.0 * w ** 2)
    return np.logaddexp(shell(-c), shell(c))


def _eggbox(x):
    tmax = 5.0 * np.pi
    t = 2.0 * tmax * x - tmax
    return -(2.0 + np.cos(t[:, 0] / 2.0) * np.cos(t[:, 1] / 2.0)) ** 5.0


def _gaussian_nd(ndim, rng):
    lam = rng.exponential(scale=1.0, size=ndim)
    Q, _ = np.linalg.qr(rng.standard_normal((ndim, ndim)))
    cov = Q @ np.diag(lam) @ Q.T + 0.5 * np.eye(ndim)
    L = np.linalg.cholesky(cov)

    def fn(x):
        z = np.linalg.solve(L, np.atleast_2d(x).T)
        return -0.5 * np.sum(z * z, axis=0) - np.sum(np.log(np.diag(L))) - 0.5 * ndim * np.log(2.0 * np.pi)
    return fn


def make_config(name, n=None, white_noise=-12.0):
    """dict(name, kind, ndim, bounds, X, y, fn, utility, hp) for config ``name``; ``n`` overrides
    the number of training points (same stream: the first ``n`` rows agree only for the
    default ``n``).  ``hp`` = dict(log_M (ndim,), amp, mean, white_noise)."""
    c = CONFIGS[name]
    d = c["ndim"]
    rng = np.random.default_rng(c["seed"])
    fn = {"rosenbrock": _rosenbrock, "gaussian_shells": _gaussian_shells, "eggbox": _eggbox}.get(c["fn"])
    if fn is None:
        fn = _gaussian_nd(d, rng)
    n = int(c["n"] if n is None else n)
    bounds = np.array([(c["lo"], c["hi"])] * d, dtype=np.float64)
    X = rng.uniform(bounds[:, 0], bounds[:, 1], size=(n, d))
    y = fn(X)
    hp = dict(log_M=np.full(d, c["log_M"]), amp=float(np.var(y)), mean=float(np.median(y)),
              white_noise=float(white_noise))
    return dict(name=name, kind=c["kind"], ndim=d, bounds=bounds, X=X, y=y, fn=lambda x: fn(np.atleast_2d(x)),
                utility=c["utility"], hp=hp, rng=rng)


def build_gp(cfg, device=None):
    """``alabi_b200.GP`` with the config's kernel and hyper-parameters (not yet computed)."""
    from . import kernels
    from .gp import GP
    hp = cfg["hp"]
    k = getattr(kernels, cfg["kind"])(metric=np.exp(hp["log_M"]), ndim=cfg["ndim"]) * hp["amp"]
    return GP(kernel=k, fit_mean=True, mean=hp["mean"], white_noise=hp["white_noise"], fit_white_noise=True,
              device=device)
