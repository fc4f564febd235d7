"""CPU oracle: the synthetic benchmark likelihoods of BASELINE.json's configs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED against the
reference's own functions via ``tests/golden/benchmarks_golden.npz``.
Vectorised restatements (rows = points) of alabi/benchmarks.py:46-52
(Rosenbrock), :100-116 (Gaussian shells), :123-132 (eggbox), :177-188
(2-D Gaussian) and :195-206 (random covariance) + the correlated N-d Gaussian
of docs/source/plot_gaussian_nd.py.
"""
import numpy as np


def rosenbrock(x):
    """-rosen(x)/100 with scipy's rosen: sum 100 (x_{i+1}-x_i^2)^2 + (1-x_i)^2."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    r = np.sum(100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2, axis=1)
    return -r / 100.0


ROSENBROCK_BOUNDS = [(-5, 5), (-5, 5)]


def gaussian_shells(x, r=2.0, w=0.1, c=3.5):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    const = np.log(1.0 / np.sqrt(2.0 * np.pi * w ** 2))

    def shell(cx):
        d = np.sqrt((x[:, 0] - cx) ** 2 + x[:, 1] ** 2)
        return const - (d - r) ** 2 / (2.0 * w ** 2)
    return np.logaddexp(shell(-c), shell(c))


GAUSSIAN_SHELLS_BOUNDS = [(-6, 6), (-6, 6)]


def eggbox(x):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    tmax = 5.0 * np.pi
    t = 2.0 * tmax * x - tmax
    return -(2.0 + np.cos(t[:, 0] / 2.0) * np.cos(t[:, 1] / 2.0)) ** 5.0


EGGBOX_BOUNDS = [(0, 1), (0, 1)]


def random_gaussian_covariance(ndim, rng):
    """Q diag(lambda) Q^T with lambda ~ Exp(1) and Q from the QR of a normal
    matrix (same construction as the reference; ``rng`` is a seeded
    ``numpy.random.Generator`` instead of the global state)."""
    lam = rng.exponential(scale=1.0, size=ndim)
    Q, _ = np.linalg.qr(rng.standard_normal((ndim, ndim)))
    return Q @ np.diag(lam) @ Q.T


def mvn_logpdf(x, mean, cov):
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    d = x.shape[1]
    L = np.linalg.cholesky(cov)
    z = np.linalg.solve(L, (x - mean).T)
    return -0.5 * np.sum(z * z, axis=0) - np.sum(np.log(np.diag(L))) - 0.5 * d * np.log(2.0 * np.pi)


def make_config(name, seed=None, n_override=None):
    """Synthetic training set for BASELINE.json config c1..c5.

    Returns dict(kind, bounds, X, y, utility, fn)."""
    spec = {
        "c1": dict(kind="ExpSquaredKernel", bounds=ROSENBROCK_BOUNDS, n=50, fn=rosenbrock, utility="bape", seed=1),
        "c2": dict(kind="Matern32Kernel", bounds=GAUSSIAN_SHELLS_BOUNDS, n=1000, fn=gaussian_shells, utility="agp", seed=2),
        "c3": dict(kind="Matern52Kernel", bounds=EGGBOX_BOUNDS, n=4000, fn=eggbox, utility="bape", seed=3),
        "c4": dict(kind="ExpSquaredKernel", bounds=[(-3, 3)] * 10, n=8192, fn=None, utility="bape", seed=4),
        "c5": dict(kind="ExpSquaredKernel", bounds=[(-3, 3)] * 20, n=16384, fn=None, utility="bape", seed=5),
    }[name]
    rng = np.random.default_rng(spec["seed"] if seed is None else seed)
    b = np.asarray(spec["bounds"], dtype=np.float64)
    d = len(b)
    fn = spec["fn"]
    if fn is None:
        cov = random_gaussian_covariance(d, rng) + 0.5 * np.eye(d)
        mean = np.zeros(d)
        fn = lambda x, m=mean, c=cov: mvn_logpdf(x, m, c)
    n = spec["n"] if n_override is None else n_override
    X = rng.uniform(b[:, 0], b[:, 1], size=(n, d))
    return dict(kind=spec["kind"], bounds=b, X=X, y=fn(X), utility=spec["utility"], fn=fn, rng=rng)
