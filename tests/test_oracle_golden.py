"""Pin the oracle against vectors produced by the reference's own functions
(tests/golden/make_golden.py) and against published known answers."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import utility as ou, benchmarks as ob, philox


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(a, b, equal_nan=True)


def test_utilities_bit_exact_vs_reference():
    g = np.load(os.path.join(GOLDEN, "utility_golden.npz"))
    inside = ou.in_bounds(g["theta"], g["bounds"])
    assert _eq(ou.lnprior_uniform(g["theta"], g["bounds"]), g["lnprior"])
    assert inside.sum() > 50 and (~inside).sum() > 50
    assert _eq(ou.bape(g["mu"], g["var"], inside), g["bape"])
    assert _eq(ou.agp(g["mu"], g["var"], inside), g["agp"])
    assert _eq(ou.jones(g["mu"], g["var"], float(g["y_best"]), 0.01, inside), g["jones"])
    # the edge rows lie ON the box: strict inequalities put them outside
    assert np.all(np.isinf(g["bape"][:4]))
    # same argmin as the reference values, lowest index on ties
    for k in ("bape", "agp", "jones"):
        assert ou.first_argmin(ou.utility(k, g["mu"], g["var"], inside, float(g["y_best"]))) == int(np.nanargmin(g[k]))


def test_logsubexp_and_prior_transform():
    g = np.load(os.path.join(GOLDEN, "utility_golden.npz"))
    assert _eq(ou.logsubexp(g["x1"], g["x2"]), g["logsubexp"])
    assert _eq(ou.prior_transform_uniform(g["u"], g["bounds"]), g["prior_transform"])
    assert _eq(ou.prior_transform_uniform(g["u"][0], g["bounds"]), g["prior_transform_1d"])


def test_normal_priors_vs_reference():
    g = np.load(os.path.join(GOLDEN, "priors_golden.npz"))
    mu, sd = g["data"][:, 0], g["data"][:, 1]
    assert _eq(ou.lnprior_normal(g["x"], g["bounds"], mu, sd), g["lnprior_normal"])
    assert np.isinf(g["lnprior_normal"]).sum() > 20 and np.isfinite(g["lnprior_normal"]).sum() > 20
    assert _eq(ou.prior_transform_normal(g["u"], g["bounds"], mu, sd), g["prior_transform_normal"])
    assert _eq(ou.prior_transform_normal(g["u"][0], g["bounds"], mu, sd), g["prior_transform_normal_1d"])


def test_regulariser_and_burnin():
    g = np.load(os.path.join(GOLDEN, "utility_golden.npz"))
    lidx = list(g["lidx"])
    for h, r, rg, rd in zip(g["hp"], g["reg"], g["reg_grad"], g["reg_default"]):
        np.testing.assert_allclose(ou.regularization_term(h, lidx, 1.3, 0.7, 1.9), r, rtol=1e-14)
        np.testing.assert_allclose(ou.regularization_gradient(h, lidx, 1.3, 0.7, 1.9), rg, rtol=1e-14)
        np.testing.assert_allclose(ou.regularization_term(h, lidx), rd, rtol=1e-14)
    for t, n, b in zip(g["taus"], g["tau_len"], g["burn"]):
        assert ou.burnin_thin(t[:n]) == (int(b[0]), int(b[1]))


def test_benchmark_functions():
    g = np.load(os.path.join(GOLDEN, "benchmarks_golden.npz"))
    np.testing.assert_allclose(ob.rosenbrock(g["xr"]), g["rosenbrock"], rtol=1e-13)
    np.testing.assert_allclose(ob.gaussian_shells(g["xs"]), g["shells"], rtol=1e-13)
    np.testing.assert_allclose(ob.eggbox(g["xe"]), g["eggbox"], rtol=1e-13)
    mean, cov = np.array([0.5, 0.5]), np.diag([0.1, 0.1])
    np.testing.assert_allclose(ob.mvn_logpdf(g["xg"], mean, cov), g["gaussian_2d"], rtol=1e-12)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32(*[np.array([x]) for x in c], *k)
        assert tuple(int(g[0]) for g in got) == want
