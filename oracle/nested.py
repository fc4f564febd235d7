"""CPU oracle: replay of the device's constrained random walks (ab_nested_walk).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The algorithm is dynesty's
``sample="rwalk"`` replacement step as the reference configures it (alabi/core.py:2600-2641,
dynesty unpinned in /root/reference/setup.py:17 and not installed: **parity unpinned** against
dynesty itself); the draw layout is this repo's own (include/alabi_b200.h, ``ab_nested_config``)
and is restated here independently of the CUDA source:

    Philox4x32-10, key = seed, counter = (chain_offset + chain, step, pair p, launch counter)
    u1 = 1 - u53(w0, w1), u2 = u53(w2, w3);  z[2p] = sqrt(-2 ln u1) cos(2 pi u2), z[2p+1] = ... sin
    u' = u + scale * C z   (C lower triangular);  accepted iff u' in (0, 1)^d and L(T(u')) > L_min
"""
import numpy as np

from .philox import philox4x32, u53


def walk_normals(seed, counter, chains, step, ndim, chain_offset=0):
    """z (nchains, ndim) of one walk step."""
    c = np.asarray(chains, dtype=np.uint64) + np.uint64(chain_offset)
    z = np.zeros((len(c), ndim))
    for p in range((ndim + 1) // 2):
        r = philox4x32(c, step, p, counter, seed & 0xFFFFFFFF, seed >> 32)
        u1 = 1.0 - u53(r[0], r[1])
        u2 = u53(r[2], r[3])
        rad = np.sqrt(-2.0 * np.log(u1))
        z[:, 2 * p] = rad * np.cos(2.0 * np.pi * u2)
        if 2 * p + 1 < ndim:
            z[:, 2 * p + 1] = rad * np.sin(2.0 * np.pi * u2)
    return z


def replay_walk(u0, theta0, logl0, lmin, scale, chol, walks, seed, counter, transform, loglike):
    """End points (u, theta, logl, naccept per chain) of ``walks`` steps from (u0, theta0, logl0)."""
    u, theta, logl = np.array(u0, dtype=np.float64), np.array(theta0, dtype=np.float64), np.array(logl0, dtype=np.float64)
    k, d = u.shape
    C = np.tril(np.asarray(chol, dtype=np.float64))
    nacc = np.zeros(k, dtype=np.int64)
    margins = []
    for step in range(int(walks)):
        z = walk_normals(seed, counter, np.arange(k), step, d)
        prop = u + scale * (z @ C.T)
        inside = np.all((prop > 0.0) & (prop < 1.0), axis=1)
        if inside.any():
            t_in = transform(prop[inside])
            l_in = np.asarray(loglike(t_in), dtype=np.float64)
            margins.append(np.abs(l_in - lmin))
            ok = l_in > lmin
            idx = np.nonzero(inside)[0][ok]
            u[idx], theta[idx], logl[idx] = prop[idx], t_in[ok], l_in[ok]
            nacc[idx] += 1
    return u, theta, logl, nacc, (np.min(np.concatenate(margins)) if margins else np.inf)
