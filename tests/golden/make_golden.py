"""Generate golden vectors from the REAL reference (run in the build container).

    python tests/golden/make_golden.py

/root/reference is pure Python but imports george / emcee / dynesty / skopt /
matplotlib / corner, none of which are installed (no network).  Those modules
are stubbed in ``sys.modules`` so that ``import alabi`` succeeds; only the
reference's pure-NumPy functions are then executed (acquisition utilities,
priors, regulariser, burn-in heuristic, benchmark likelihoods).  The outputs
are frozen under tests/golden/*.npz and pin ``oracle.utility`` and
``oracle.benchmarks``.  The GPU box has no /root/reference: nothing in the
test-suite calls this script.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ALABI_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    _stub("george", kernels=_stub("george.kernels"), GP=_Any)
    _stub("skopt")
    _stub("skopt.space", Space=_Any)
    _stub("skopt.space.space", Real=_Any)
    _stub("skopt.sampler", Sobol=_Any, Lhs=_Any, Halton=_Any, Hammersly=_Any, Grid=_Any)
    mpl = _stub("matplotlib", rcParams={}, rc=lambda *a, **k: None, use=lambda *a, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors = _stub("matplotlib.colors", LogNorm=_Any, Normalize=_Any)
    mpl.cm = _stub("matplotlib.cm")
    mpl.gridspec = _stub("matplotlib.gridspec")
    mpl.ticker = _stub("matplotlib.ticker")
    mpl.patches = _stub("matplotlib.patches")
    mpl.lines = _stub("matplotlib.lines")
    _stub("corner")
    _stub("emcee")
    dy = _stub("dynesty")
    dy.plotting = _stub("dynesty.plotting")
    dy.utils = _stub("dynesty.utils")
    _stub("h5py")
    sys.path.insert(0, REF)
    import alabi  # noqa: F401
    from alabi import utility, gp_utils, mcmc_utils, benchmarks
    return utility, gp_utils, mcmc_utils, benchmarks


def main():
    ut, gpu, mcu, bm = import_reference()
    rng = np.random.default_rng(20261018)

    # ---- acquisition utilities ------------------------------------------------
    d, M = 3, 400
    bounds = np.array([(-1.0, 2.0), (0.0, 1.0), (-5.0, 5.0)])
    theta = rng.uniform(bounds[:, 0] - 0.3, bounds[:, 1] + 0.3, size=(M, d))
    theta[:4] = [[-1.0, 0.5, 0.0], [2.0, 0.5, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 5.0]]  # on the edge
    mu = rng.normal(0.0, 3.0, size=M)
    var = np.exp(rng.uniform(-30.0, 3.0, size=M))
    var[10:14] = [0.0, -1e-12, 1e-300, 750.0]          # degenerate variances
    y_best = 1.25
    out = {k: np.empty(M) for k in ("bape", "agp", "jones", "lnprior")}
    with np.errstate(all="ignore"):
        for i in range(M):
            pg = lambda x, i=i: (np.array([mu[i]]), np.array([var[i]]))
            out["bape"][i] = ut.bape_utility(theta[i], pg, bounds)
            out["agp"][i] = ut.agp_utility(theta[i], pg, bounds)
            out["jones"][i] = ut.jones_utility(theta[i], pg, bounds, y_best)
            out["lnprior"][i] = ut.lnprior_uniform(theta[i], bounds)
    x1 = rng.normal(0, 5, size=200)
    x2 = rng.normal(0, 5, size=200)
    x2[:5] = x1[:5]
    with np.errstate(all="ignore"):
        lse = np.array([ut.logsubexp(a, b) for a, b in zip(x1, x2)])
    u = rng.uniform(size=(50, d))
    pt = ut.prior_transform_uniform(u, bounds)
    pt1 = ut.prior_transform_uniform(u[0], bounds)

    # ---- regulariser -------------------------------------------------------------
    hp = rng.normal(0, 1.5, size=(20, 7))
    lidx = [3, 4, 5, 6]
    reg = np.array([gpu.regularization_term(h, lidx, amp_0=1.3, mu_0=0.7, sigma_0=1.9) for h in hp])
    regg = np.array([gpu.regularization_gradient(h, lidx, amp_0=1.3, mu_0=0.7, sigma_0=1.9) for h in hp])
    reg_d = np.array([gpu.regularization_term(h, lidx) for h in hp])

    # ---- burn-in heuristic -------------------------------------------------------
    taus = [np.array([12.3, 40.9, 7.7]), np.array([np.nan, 55.5]), np.array([np.nan, np.nan]), np.array([0.4, 0.9])]
    burn = []
    for t in taus:
        class S:
            def get_autocorr_time(self, tol=0, t=t):
                return t.copy()
        burn.append(mcu.estimate_burnin(S()))

    np.savez(os.path.join(HERE, "utility_golden.npz"), bounds=bounds, theta=theta, mu=mu, var=var,
             y_best=y_best, bape=out["bape"], agp=out["agp"], jones=out["jones"], lnprior=out["lnprior"],
             x1=x1, x2=x2, logsubexp=lse, u=u, prior_transform=pt, prior_transform_1d=pt1,
             hp=hp, lidx=np.array(lidx), reg=reg, reg_grad=regg, reg_default=reg_d,
             taus=np.array([np.pad(t, (0, 3 - len(t)), constant_values=-1.0) for t in taus]),
             tau_len=np.array([len(t) for t in taus]), burn=np.array(burn))

    # ---- benchmark likelihoods ---------------------------------------------------
    xr = rng.uniform(-5, 5, size=(64, 2))
    xs = rng.uniform(-6, 6, size=(64, 2))
    xe = rng.uniform(0, 1, size=(64, 2))
    np.savez(os.path.join(HERE, "benchmarks_golden.npz"),
             xr=xr, rosenbrock=np.array([bm.rosenbrock_fn(x) for x in xr]),
             xs=xs, shells=np.array([bm.gaussian_shells_fn(x) for x in xs]),
             xe=xe, eggbox=np.array([bm.eggbox_fn(x) for x in xe]),
             xg=xe, gaussian_2d=np.array([bm.gaussian_2d_fn(x) for x in xe]))
    # ---- normal / mixed priors (own generator so the files above keep their bits) ----
    rng2 = np.random.default_rng(20261019)
    pb = np.array([(-2.0, 2.0), (0.0, 10.0), (-1.0, 1.0), (3.0, 4.0)])
    pdata = [(None, None), (5.0, 1.5), (0.2, 0.05), (None, None)]
    px = rng2.uniform(pb[:, 0] - 0.2, pb[:, 1] + 0.2, size=(300, 4))
    px[:3] = [[-2.0, 5.0, 0.2, 3.5], [0.0, 10.0, 1.0, 4.0], [0.0, 5.0, 0.2, 3.5]]
    with np.errstate(all="ignore"):
        lnpn = np.array([ut.lnprior_normal(x, pb, pdata) for x in px])
    pu = rng2.uniform(size=(64, 4))
    ptn = ut.prior_transform_normal(pu, pb, pdata)
    ptn1 = ut.prior_transform_normal(pu[0], pb, pdata)
    np.savez(os.path.join(HERE, "priors_golden.npz"), bounds=pb,
             data=np.array([[np.nan if v is None else v for v in t] for t in pdata]),
             x=px, lnprior_normal=lnpn, u=pu, prior_transform_normal=ptn, prior_transform_normal_1d=ptn1)
    # ---- remaining benchmark functions -------------------------------------------
    rng3 = np.random.default_rng(20261020)
    x1d = rng3.uniform(-2, 1, size=40)
    bmat = rng3.uniform(0.5, 2.0, size=(4, 3))
    xnd = rng3.normal(0, 1, size=(16, 10))
    xm = rng3.uniform(0, 5, size=(32, 2))
    means = [np.array([0.0, 0.0, 0.0]), np.array([1.0, -1.0, 0.5])]
    covs = [np.eye(3) * 0.5, np.diag([0.2, 0.4, 0.9])]
    xg3 = rng3.normal(0, 1, size=(12, 3))
    np.savez(os.path.join(HERE, "benchmarks2_golden.npz"),
             x1d=x1d, test1d=bm.test1d_fn(x1d), bmat=bmat, xnd=xnd,
             rosenbrock_nd=bm.rosenbrock_nd(xnd, 0.7, bmat), rosenbrock_nd_1=bm.rosenbrock_nd(xnd[0], 0.7, bmat),
             xm=xm, multimodal=np.array([bm.multimodal_fn(x) for x in xm]),
             logcirc=bm.logcirc(xm, np.array([3.5, 0.0])),
             xg3=xg3, mmg=bm.multimodal_gaussian_nd(xg3, means, covs, [0.3, 0.6]),
             wy=xnd[:, 0], wp=xnd[:, 1],
             wmse=np.array([gpu.weighted_mse_by_probability(xnd[:, 0], xnd[:, 1], weight_method=m, temperature=1.7)
                            for m in ("exponential", "linear", "softmax", "rank")]))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
