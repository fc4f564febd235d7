"""Self-consistency of the (unpinned) george / emcee restatements."""
import numpy as np
import pytest
from scipy import stats

from oracle import gp as ogp, emcee as oem, philox, benchmarks as ob


def _data(n=60, d=3, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(n, d))
    y = np.sin(X.sum(axis=1)) + 0.1 * rng.normal(size=n)
    return X, y


@pytest.mark.parametrize("kind", ["ExpSquaredKernel", "Matern32Kernel", "Matern52Kernel"])
def test_grad_loglike_matches_finite_differences(kind):
    X, y = _data()
    gp = ogp.make_gp(kind, X, y, log_M=[-0.3, 0.2, 0.5], amp=np.var(y), white_noise=-5.0)
    p = gp.get_parameter_vector()
    assert gp.get_parameter_names() == ("mean:value", "white_noise:value", "kernel:k1:log_constant",
                                        "kernel:k2:metric:log_M_0_0", "kernel:k2:metric:log_M_1_1",
                                        "kernel:k2:metric:log_M_2_2")
    g = gp.grad_log_likelihood(y)
    for i in range(len(p)):
        h = 1e-6
        pp, pm = p.copy(), p.copy()
        pp[i] += h
        pm[i] -= h
        gp.set_parameter_vector(pp)
        lp = gp.log_likelihood(y)
        gp.set_parameter_vector(pm)
        lm = gp.log_likelihood(y)
        assert abs((lp - lm) / (2 * h) - g[i]) < 2e-5 * max(1.0, abs(g[i]))


def test_closed_forms_and_interpolation():
    r2 = np.array([0.0, 1e-30, 400.0])
    for kind in (0, 1, 2):
        k = ogp.radial(kind, r2)
        assert k[0] == 1.0 and abs(k[1] - 1.0) < 1e-14 and k[2] < 1e-8
    X, y = _data(40, 2)
    gp = ogp.make_gp("Matern32Kernel", X, y, log_M=[-2.0, -2.0], amp=np.var(y), white_noise=-25.0)
    mu, var = gp.predict(y, X, return_var=True)
    np.testing.assert_allclose(mu, y, atol=1e-5)
    assert np.all(var > -1e-6) and np.all(var < 1e-4)
    np.testing.assert_allclose(gp.predict_var_via_L(X[:5] + 0.3), gp.predict(y, X[:5] + 0.3, return_var=True)[1], rtol=1e-6)


def test_not_spd_is_quiet_minus_inf():
    X, y = _data(20, 2)
    X[1] = X[0]
    gp = ogp.make_gp("ExpSquaredKernel", X, y, log_M=[0.0, 0.0], amp=2.0, white_noise=-80.0, compute=False)
    gp._x = X
    gp._yerr2 = np.zeros(len(X))
    assert gp.log_likelihood(y, quiet=True) == -np.inf
    assert np.all(gp.grad_log_likelihood(y, quiet=True) == 0.0)


def test_replay_equals_explicit_draw_loop_and_targets_gaussian():
    """Stretch move with the device draw layout leaves N(0, diag(1, 4)) invariant."""
    sig = np.array([1.0, 2.0])
    lp = lambda q: -0.5 * np.sum((np.atleast_2d(q) / sig) ** 2, axis=1)
    rng = np.random.default_rng(3)
    p0 = rng.normal(size=(64, 2)) * sig
    for rs in (True, False):
        chain, lps, nacc, rec = oem.replay_device_chain(p0, lp, 600, seed=1234, randomize_split=rs)
        flat = chain[200:].reshape(-1, 2)
        tau = oem.integrated_time(chain[200:], tol=0)
        neff = flat.shape[0] / tau.max()
        assert abs(flat[:, 0].mean()) < 5 * sig[0] / np.sqrt(neff)
        assert abs(flat[:, 1].std() - 2.0) < 0.15
        thin = chain[200::int(2 * tau.max()) + 1].reshape(-1, 2)
        assert stats.kstest(thin[:, 0], "norm").pvalue > 1e-3
        assert 0.3 < nacc.mean() / 600 < 0.95
        # per-proposal record is consistent
        r = rec[10]
        ok = r["partner"] >= 0
        assert np.all(r["sets"][ok] != r["sets"][r["partner"][ok]])
        np.testing.assert_allclose(r["lp"][ok], lp(r["q"][ok]))


def test_box_times_normal_prior_posterior():
    """Gaussian likelihood N(0, 0.6) x prior N(0.5, 0.3) on theta_0 (ut.lnprior_normal's oracle
    restatement) in a box: the stretch move with the device draw layout samples the analytic
    posterior N(0.4, 0.268) on theta_0 and the truncated likelihood on theta_1."""
    from oracle import utility as ou
    b = np.array([(-3.0, 3.0), (-3.0, 3.0)])
    mu, sd = np.array([0.5, np.nan]), np.array([0.3, np.nan])
    like = lambda q: -0.5 * (np.atleast_2d(q)[:, 0] / 0.6) ** 2 - 0.5 * (np.atleast_2d(q)[:, 1] / 0.9) ** 2
    with np.errstate(invalid="ignore"):
        lp = lambda q: like(q) + ou.lnprior_normal(q, b, mu, sd)
        rng = np.random.default_rng(8)
        p0 = rng.uniform(-1, 1, size=(48, 2))
        chain, lps, nacc, rec = oem.replay_device_chain(p0, lp, 900, seed=77)
    flat = chain[300:].reshape(-1, 2)
    assert abs(flat[:, 0].mean() - 0.4) < 0.05 and abs(flat[:, 0].std() - 0.2683) < 0.04
    assert abs(flat[:, 1].std() - 0.9) < 0.12 and np.all(np.abs(flat) < 3.0)
    assert np.all(np.isfinite(lps[300:]))


def test_emcee_order_sampler_targets_gaussian():
    lp = lambda q: -0.5 * np.sum(np.atleast_2d(q) ** 2, axis=1)
    s = oem.StretchEnsemble(40, 2, lp, seed=5, vectorize=True)
    s.run_mcmc(np.random.default_rng(0).normal(size=(40, 2)), 500)
    flat = s.get_chain(discard=100, flat=True)
    assert abs(flat.std() - 1.0) < 0.1
    assert s.get_chain(discard=10, thin=7).shape[0] == len(range(10 + 6, 500, 7))


def test_move_draws_layout():
    d = philox.move_draws(99, 33, 7, randomize_split=False)
    assert np.array_equal(d["sets"], np.arange(33) % 2)
    assert np.all(d["partner"] % 2 != d["sets"]) and d["partner"].max() < 33
    d2 = philox.move_draws(99, 33, 7, randomize_split=True, walker_offset=1000)
    assert np.all((d2["u_z"] >= 0) & (d2["u_z"] < 1)) and set(np.unique(d2["sets"])) == {0, 1}


def test_config_generators():
    for name, n in (("c1", 50), ("c2", 64), ("c4", 32)):
        c = ob.make_config(name, n_override=n)
        assert c["X"].shape == (n, len(c["bounds"])) and np.all(np.isfinite(c["y"]))
