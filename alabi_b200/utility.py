"""Host-side helpers with the names and semantics of alabi/utility.py.

* scalers ``no_scaler / nlog_scaler / log_scaler``      alabi/utility.py:45-72
* ``prior_sampler``                                      alabi/utility.py:79-199
  (scikit-optimize is replaced by ``scipy.stats.qmc``: same designs, different
  sequences — the inputs are random draws anyway)
* ``lnprior_uniform`` / ``prior_transform_uniform``      alabi/utility.py:218-345
* ``logsubexp``                                          alabi/utility.py:489-504
* ``agp_utility / bape_utility / jones_utility``         alabi/utility.py:629-946
  (scalar, one-point forms for scipy optimisers; the batched evaluation lives on
  the GPU: ``GP.utility_argmin`` -> K4)
* ``assign_utility`` / ``minimize_objective``            alabi/utility.py:949-1163
"""
import time
import warnings

import numpy as np
from scipy.optimize import minimize
from scipy.stats import norm, qmc
from sklearn import preprocessing

__all__ = ["agp_utility", "bape_utility", "jones_utility", "assign_utility", "minimize_objective",
           "prior_sampler", "prior_sampler_normal", "lnprior_uniform", "lnprior_normal",
           "prior_transform_uniform", "prior_transform_normal", "logsubexp",
           "NewFunctionTransformer", "nlog_scaler", "log_scaler", "no_scaler", "scaler_affine"]


# ---------------------------------------------------------------------------
# scalers
# ---------------------------------------------------------------------------
class NewFunctionTransformer(preprocessing.FunctionTransformer):
    def __init__(self, name, func=None, inverse_func=None, *, validate=False, accept_sparse=False,
                 check_inverse=True, feature_names_out=None, kw_args=None, inv_kw_args=None):
        super().__init__(func=func, inverse_func=inverse_func, validate=validate, accept_sparse=accept_sparse,
                         check_inverse=check_inverse, feature_names_out=feature_names_out, kw_args=kw_args,
                         inv_kw_args=inv_kw_args)
        self.name = name

    def __str__(self):
        return self.name

    __repr__ = __str__


def nlog(x): return np.log10(-x)
def nlog_inv(x): return -10 ** x
def log_scale(x): return np.log10(x)
def log_scale_inv(logx): return 10 ** logx
def no_scale(x): return x


nlog_scaler = NewFunctionTransformer(name="nlog_scaler", func=nlog, inverse_func=nlog_inv)
log_scaler = NewFunctionTransformer(name="log_scaler", func=log_scale, inverse_func=log_scale_inv)
no_scaler = NewFunctionTransformer(name="no_scaler", func=no_scale, inverse_func=no_scale)


def scaler_affine(scaler, ndim, inverse=False):
    """Express a fitted scaler as what the device sampler understands.

    forward (theta): returns (scale, offset) with x_scaled = x * scale + offset.
    inverse (y): returns (kind, scale, offset): y = ys * scale + offset (kind 0),
    -10**ys (1), 10**ys (2).  Raises for scalers with no closed form."""
    name = getattr(scaler, "name", None)
    if not inverse:
        if name == "no_scaler":
            return np.ones(ndim), np.zeros(ndim)
        if isinstance(scaler, preprocessing.MinMaxScaler):
            return np.asarray(scaler.scale_, dtype=float), np.asarray(scaler.min_, dtype=float)
        if isinstance(scaler, preprocessing.StandardScaler):
            s = np.asarray(scaler.scale_, dtype=float) if scaler.with_std else np.ones(ndim)
            m = np.asarray(scaler.mean_, dtype=float) if scaler.with_mean else np.zeros(ndim)
            return 1.0 / s, -m / s
        raise NotImplementedError(f"theta_scaler {scaler!r} has no affine form for the device sampler")
    if name == "no_scaler":
        return 0, 1.0, 0.0
    if name == "nlog_scaler":
        return 1, 1.0, 0.0
    if name == "log_scaler":
        return 2, 1.0, 0.0
    if isinstance(scaler, preprocessing.MinMaxScaler):
        return 0, 1.0 / float(scaler.scale_[0]), -float(scaler.min_[0]) / float(scaler.scale_[0])
    if isinstance(scaler, preprocessing.StandardScaler):
        s = float(scaler.scale_[0]) if scaler.with_std else 1.0
        m = float(scaler.mean_[0]) if scaler.with_mean else 0.0
        return 0, s, m
    raise NotImplementedError(f"y_scaler {scaler!r} has no closed form for the device sampler")


# ---------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------
def prior_sampler(bounds=None, nsample=1, sampler="uniform", random_state=None):
    """Draw ``nsample`` points inside ``bounds`` ((min, max) per dimension).

    sampler: 'uniform', 'sobol', 'lhs', 'halton', 'hammersly', 'grid'."""
    b = np.asarray(bounds, dtype=np.float64)
    ndim = len(b)
    lo, hi = b[:, 0], b[:, 1]
    nsample = int(nsample)
    if random_state is None:
        random_state = int(time.time() * 1000000) % (2 ** 32)
    if sampler == "uniform":
        u = np.random.default_rng(random_state).uniform(size=(nsample, ndim))
    elif sampler == "sobol":
        u = qmc.Sobol(d=ndim, scramble=False).random(nsample + 1)[1:]
    elif sampler == "lhs":
        u = qmc.LatinHypercube(d=ndim, seed=random_state).random(nsample)
    elif sampler == "halton":
        u = qmc.Halton(d=ndim, scramble=False).random(nsample + 1)[1:]
    elif sampler == "hammersly":
        h = qmc.Halton(d=max(ndim - 1, 1), scramble=False).random(nsample + 1)[1:]
        first = (np.arange(nsample) + 0.5) / nsample
        u = np.column_stack([first, h])[:, :ndim]
    elif sampler == "grid":
        per = int(np.ceil(nsample ** (1.0 / ndim)))
        axes = [np.linspace(0.0, 1.0, per)] * ndim
        u = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, ndim)[:nsample]
    else:
        raise ValueError(f"Sampler method '{sampler}' not implemented. Valid options for 'sampler' are: "
                         "uniform, sobol, lhs, halton, hammersly, grid.")
    return lo + (hi - lo) * u


def lnprior_uniform(x, bounds):
    """0.0 if every coordinate is strictly inside its bounds, else -inf."""
    ndim = len(bounds)
    x = np.array([x]) if ndim == 1 else np.array(x).squeeze()
    x = np.atleast_1d(x).reshape(-1)
    for i in range(ndim):
        if not ((x[i] > bounds[i][0]) and (x[i] < bounds[i][1])):
            return -np.inf
    return 0


def prior_sampler_normal(prior_data, bounds, nsample=1, random_state=None):
    """Truncated-normal prior draws (alabi/utility.py:202-215): dimensions whose
    ``prior_data[i]`` is ``(mean, std)`` are drawn with ``scipy.stats.truncnorm`` inside
    ``bounds``, ``(None, None)`` dimensions uniformly.  ``random_state`` (seed or Generator)
    makes the draw reproducible; None uses NumPy's global state like the reference."""
    from scipy.stats import truncnorm
    b = np.asarray(bounds, dtype=np.float64)
    rng = np.random.default_rng(random_state) if random_state is not None else None
    out = np.zeros((len(b), int(nsample)))
    for i in range(len(b)):
        mu, sd = prior_data[i]
        if mu is not None:
            lb, ub = (b[i, 0] - mu) / sd, (b[i, 1] - mu) / sd
            out[i] = truncnorm.rvs(lb, ub, loc=mu, scale=sd, size=int(nsample), random_state=rng)
        elif rng is not None:
            out[i] = rng.uniform(b[i, 0], b[i, 1], size=int(nsample))
        else:
            out[i] = np.random.uniform(low=b[i, 0], high=b[i, 1], size=int(nsample))
    return out.T


def lnprior_normal(x, bounds, data):
    """Uniform box times independent normals (alabi/utility.py:370-378); ``data[i]`` is
    ``(mean, std)`` or ``(None, None)``.  The normal terms use scipy's ``norm.logpdf``
    expression, added in dimension order (the device sampler adds them the same way)."""
    lnp = lnprior_uniform(x, bounds)
    x = np.asarray(x, dtype=np.float64).flatten()
    for i in range(len(x)):
        if data[i][0] is not None:
            z = (x[i] - data[i][0]) / data[i][1]
            lnp += (-(z * z) / 2.0 - 0.9189385332046727) - np.log(data[i][1])
    return lnp


def prior_transform_normal(x, bounds, data):
    """Unit cube -> prior draws for nested sampling (alabi/utility.py:381-486): uniform
    dimensions map linearly onto ``bounds``; normal ones through the (untruncated, as in
    the reference) inverse normal CDF.  1-D or (nsamples, ndim)."""
    from scipy.stats import norm as _norm
    x = np.asarray(x, dtype=float)
    bounds = np.asarray(bounds)
    if x.ndim not in (1, 2):
        raise ValueError(f"x must be 1D or 2D array, got {x.ndim}D array with shape {x.shape}")
    ndim = x.shape[-1]
    if len(bounds) != ndim or len(data) != ndim:
        raise ValueError(f"Bounds length ({len(bounds)}) and data length ({len(data)}) "
                         f"must match x dimensions ({ndim})")
    out = np.zeros(x.shape)
    for i, (lo, hi) in enumerate(bounds):
        if data[i][0] is None:
            out[..., i] = (hi - lo) * x[..., i] + lo
        else:
            out[..., i] = _norm.ppf(x[..., i], data[i][0], data[i][1])
    return out


def prior_transform_uniform(theta, bounds):
    """Unit hypercube -> box: (hi - lo) * u + lo; 1-D or (nsamples, ndim)."""
    theta = np.asarray(theta, dtype=float)
    b = np.asarray(bounds, dtype=float)
    if theta.ndim not in (1, 2):
        raise ValueError("theta must be 1D or 2D")
    return (b[:, 1] - b[:, 0]) * theta + b[:, 0]


def logsubexp(x1, x2):
    """log(exp(x1) - exp(x2)), -inf when x1 <= x2."""
    if x1 <= x2:
        return -np.inf
    return x1 + np.log(1.0 - np.exp(x2 - x1))


# ---------------------------------------------------------------------------
# one-point acquisition utilities (host callables for scipy)
# ---------------------------------------------------------------------------
def agp_utility(theta, predict_gp, bounds):
    if not np.isfinite(lnprior_uniform(theta, bounds)):
        return np.inf
    mu, var = predict_gp(np.asarray(theta).reshape(1, -1))
    with np.errstate(all="ignore"):
        util = -(mu + 0.5 * np.log(2.0 * np.pi * np.e * var))
    return np.asarray(util).item()


def bape_utility(theta, predict_gp, bounds):
    theta = np.asarray(theta).flatten()
    if not np.isfinite(lnprior_uniform(theta, bounds)):
        return np.inf
    mu, var = predict_gp(theta.reshape(1, -1))
    mu, var = np.asarray(mu).item(), np.asarray(var).item()
    with np.errstate(all="ignore"):
        util = -((2.0 * mu + var) + logsubexp(var, 0.0))
    return float(util)


def jones_utility(theta, predict_gp, bounds, y_best, zeta=0.01):
    if not np.isfinite(lnprior_uniform(theta, bounds)):
        return np.inf
    mu, var = predict_gp(np.asarray(theta).reshape(1, -1))
    mu, var = np.asarray(mu).item(), np.asarray(var).item()
    with np.errstate(all="ignore"):
        std = np.sqrt(var)
    if not (std > 0):
        return 0.0
    z = (mu - y_best - zeta) / std
    return float(-((mu - y_best - zeta) * norm.cdf(z) + std * norm.pdf(z)))


def numerical_kernel_gradient(xs, x_train, gp, h=1e-6):
    """Central difference of k(xs, x_train) w.r.t. the single query point ``xs``
    (alabi/utility.py:511-556), shape (n_train, d).  Kept for API compatibility; the
    acquisition gradients below use the analytic derivative on the device."""
    xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
    if xs.shape[0] != 1:
        raise ValueError("This function handles single query point only")
    x_train = np.atleast_2d(np.asarray(x_train, dtype=np.float64))
    out = np.zeros((x_train.shape[0], x_train.shape[1]))
    for i in range(x_train.shape[1]):
        step = np.zeros(x_train.shape[1])
        step[i] = h
        kp = gp.kernel.get_value(xs + step, x_train).ravel()
        km = gp.kernel.get_value(xs - step, x_train).ravel()
        out[:, i] = (kp - km) / (2.0 * h)
    return out


def grad_gp_mean_prediction(xs, gp):
    """d mu / dx at one query point (alabi/utility.py:558-583), from ``GP.predict_grad``."""
    return gp.predict_grad(gp._y, np.asarray(xs, dtype=np.float64).reshape(1, -1))[2][0]


def grad_gp_var_prediction(xs, gp):
    """d sigma^2 / dx at one query point (alabi/utility.py:586-621), from ``GP.predict_grad``."""
    return gp.predict_grad(gp._y, np.asarray(xs, dtype=np.float64).reshape(1, -1))[3][0]


def grad_agp_utility(theta, gp, bounds):
    """alabi/utility.py:704-726: -(d mu + 0.5 d sigma^2) (the reference's expression,
    kept as is), inf outside the prior box.  d mu and d sigma^2 come from the device
    (``GP.predict_grad``, analytic kernel derivative; the reference differences the
    kernel with h = 1e-6 and forms a dense K^-1 per call)."""
    theta = np.asarray(theta, dtype=np.float64).flatten()
    if not np.isfinite(lnprior_uniform(theta, bounds)):
        return np.full(len(theta), np.inf)
    _, _, d_mu, d_var = gp.predict_grad(gp._y, theta.reshape(1, -1))
    return (-(d_mu[0] + 0.5 * d_var[0])).flatten()


def grad_bape_utility(theta, gp, bounds):
    """alabi/utility.py:813-850: chain rule on -((2 mu + s2) + log(exp(s2) - 1))."""
    theta = np.asarray(theta, dtype=np.float64).flatten()
    if not np.isfinite(lnprior_uniform(theta, bounds)):
        return np.full(len(theta), np.inf)
    _, var, d_mu, d_var = gp.predict_grad(gp._y, theta.reshape(1, -1))
    with np.errstate(all="ignore"):
        exp_var = np.exp(var[0])
        d_bape_d_var = -(1.0 + exp_var / (exp_var - 1.0))
    return -2.0 * d_mu[0] + d_bape_d_var * d_var[0]


def assign_utility(algorithm):
    """(utility, gradient) pair (alabi/utility.py:949-966); jones has no gradient."""
    table = {"bape": (bape_utility, grad_bape_utility), "agp": (agp_utility, grad_agp_utility),
             "jones": (jones_utility, None)}
    if algorithm not in table:
        print(f"ERROR: Unknown utility function: {algorithm}. Defaulting to BAPE.")
        return bape_utility, grad_bape_utility
    return table[algorithm]


def _minimize_single(obj_fn, bounds, x0, method, options, grad_obj_fn=None):
    res = minimize(fun=obj_fn, x0=np.array(x0).flatten(), jac=grad_obj_fn, bounds=bounds, method=method,
                   options=options)
    x_opt, f_opt = res.x, res.fun
    if np.all(np.isfinite(x_opt)) and np.all(np.isfinite(f_opt)):
        if np.isfinite(lnprior_uniform(x_opt, bounds)):
            if res.nit > 5:
                return x_opt, f_opt
            print(f"Warning: Aquisition function ran for {res.nit} iterations. Optimizer success: {res.success}")
            return (np.nan, np.nan) if res.nit <= 1 else (x_opt, f_opt)
        print("Warning: Acquisition function optimization prior fail", x_opt)
        return np.nan, np.nan
    print("Warning: Acquisition function optimization infinite fail", x_opt, f_opt)
    return np.nan, np.nan


def minimize_objective_single(idx, obj_fn, bounds, starting_point, method, options, grad_obj_fn=None):
    """One restart of ``minimize_objective`` (alabi/utility.py:969-1027)."""
    return _minimize_single(obj_fn, bounds, starting_point, method, options, grad_obj_fn)


def minimize_objective(obj_fn, bounds=None, nopt=1, method="l-bfgs-b", ps=None, options=None,
                       grad_obj_fn=None, pool=None, starting_points=None):
    """Multi-restart minimisation of an acquisition function; best finite,
    in-prior result wins, (nan, nan) if every restart fails."""
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    m = str(method).lower()
    if options is None:
        options = {"l-bfgs-b": {"maxiter": 100, "ftol": 1e-6, "gtol": 1e-5},
                   "nelder-mead": {"maxiter": 200, "xatol": 1e-6, "fatol": 1e-6}}.get(m, {"maxiter": 100})
    else:
        options = options.copy()
        for old, new in (("max_iter", "maxiter"), ("max_eval", "maxfev"), ("max_fun", "maxfun")):
            if old in options:
                options[new] = options.pop(old)
    if m == "nelder-mead":
        options["adaptive"] = True
        grad_obj_fn = None
    elif m == "l-bfgs-b":
        options.setdefault("maxcor", 10)
    if starting_points is None:
        starting_points = prior_sampler(bounds, nsample=nopt, sampler="lhs") if ps is None else ps(nsample=nopt)
    starting_points = np.array([np.asarray(pt).flatten() for pt in starting_points])
    results = [_minimize_single(obj_fn, bounds, sp, method, options, grad_obj_fn) for sp in starting_points]
    valid = [(t, o) for t, o in results if np.all(np.isfinite(t)) and np.isfinite(o)]
    if not valid:
        print(f"Warning: All {nopt} optimization attempts failed. Returning NaN.")
        return np.nan, np.nan
    best = int(np.argmin([o for _, o in valid]))
    return valid[best]
