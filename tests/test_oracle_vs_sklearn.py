"""Independent cross-check of the CPU oracle's GP arithmetic against
scikit-learn's GaussianProcessRegressor (a third implementation of the same
published model, installed in this image; george itself is not).

This does NOT discharge "parity unpinned" for george-specific conventions
(parameter order, the ``c/ndim`` amplitude, white-noise handling in the
predictive variance); it pins the maths those conventions wrap: kernel values,
log marginal likelihood, its gradient, predictive mean and variance.

Mapping (alabi/core.py:987-1014, gp_utils.py:222-248):
  george metric M_k = l_k^2            <->  sklearn length_scale l_k
  exp(log_constant)                    <->  ConstantKernel(constant_value)
  exp(white_noise) on the diagonal     <->  WhiteKernel(noise_level)
  d/d log_M_k = 1/2 d/d log l_k ; sklearn's gradient is w.r.t. log theta.
"""
import numpy as np
import pytest

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel  # noqa: E402

from oracle import gp as ogp  # noqa: E402


def _sk_kernel(kind, amp, ell, wn):
    if kind == "ExpSquaredKernel":
        base = RBF(length_scale=ell)
    elif kind == "Matern32Kernel":
        base = Matern(length_scale=ell, nu=1.5)
    else:
        base = Matern(length_scale=ell, nu=2.5)
    return ConstantKernel(amp) * base + WhiteKernel(noise_level=wn)


@pytest.mark.parametrize("kind", ["ExpSquaredKernel", "Matern32Kernel", "Matern52Kernel"])
@pytest.mark.parametrize("n,d", [(60, 1), (200, 2), (350, 5)])
def test_oracle_matches_sklearn(kind, n, d):
    rng = np.random.default_rng(1000 * n + d)
    X = rng.uniform(-2, 2, size=(n, d))
    y = np.sin(X.sum(axis=1)) - 0.3 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
    ell = rng.uniform(0.6, 1.8, size=d)
    amp, wn, mean = 1.7 * np.var(y), np.exp(-7.0), float(np.median(y))
    o = ogp.OracleGP(kind, d, 2.0 * np.log(ell), log_const=np.log(amp), mean=mean, fit_mean=True,
                     white_noise=np.log(wn), fit_white_noise=True)
    o.compute(X)
    kern = _sk_kernel(kind, amp, ell, wn)
    gpr = sk.GaussianProcessRegressor(kernel=kern, alpha=0.0, optimizer=None, normalize_y=False)
    gpr.fit(X, y - mean)

    # kernel matrix (a1/a2)
    K_o = o.get_matrix(X)
    K_o[np.diag_indices(n)] += wn
    np.testing.assert_allclose(K_o, kern(X), rtol=1e-12, atol=1e-14)

    # log marginal likelihood and gradient (a3/a4)
    ll_s, g_s = gpr.log_marginal_likelihood(kern.theta, eval_gradient=True)
    ll_o = o.log_likelihood(y)
    assert abs(ll_o - ll_s) <= 1e-9 * abs(ll_s)
    g_o = o.grad_log_likelihood(y)             # [mean, white_noise, log_constant, log_M...]
    # sklearn theta order: [log constant, log l_0.., log noise]
    scale = np.max(np.abs(g_s))
    np.testing.assert_allclose(g_o[2], g_s[0], rtol=1e-7, atol=1e-9 * scale)
    np.testing.assert_allclose(g_o[3:], 0.5 * g_s[1:1 + d], rtol=1e-7, atol=1e-9 * scale)
    np.testing.assert_allclose(g_o[1], g_s[-1], rtol=1e-7, atol=1e-9 * scale)
    # d/d mean = sum(alpha): finite-difference check (sklearn has no mean parameter)
    h = 1e-5
    op, om = (ogp.OracleGP(kind, d, 2.0 * np.log(ell), log_const=np.log(amp), mean=mean + s, fit_mean=True,
                           white_noise=np.log(wn), fit_white_noise=True).compute(X) for s in (h, -h))
    fd = (op.log_likelihood(y) - om.log_likelihood(y)) / (2 * h)
    assert abs(g_o[0] - fd) <= 1e-5 * max(1.0, abs(fd))

    # prediction (a6).  sklearn adds the WhiteKernel to k(x*, x*); george does not.
    t = rng.uniform(-2.2, 2.2, size=(400, d))
    mu_s, sd_s = gpr.predict(t, return_std=True)
    mu_o, var_o = o.predict(y, t, return_var=True)
    np.testing.assert_allclose(mu_o, mu_s + mean, rtol=1e-9, atol=1e-9 * np.max(np.abs(mu_s)))
    np.testing.assert_allclose(var_o + wn, sd_s ** 2, rtol=0, atol=2e-8 * amp)
