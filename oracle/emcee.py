"""CPU oracle: emcee 3.x ensemble stretch move as alabi drives it.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  **Parity unpinned**:
emcee (>=3.0, /root/reference/setup.py:16) is not vendored and not installed;
this restates its published algorithm (Goodman & Weare 2010 stretch move in
the red-blue form of ``emcee.moves.RedBlueMove`` / ``StretchMove``) anchored on
the reference's call sites:

* ``emcee.EnsembleSampler(nwalkers, ndim, self.lnprob, pool, **kw)``  alabi/core.py:2319
* ``run_mcmc(p0, nsteps, progress=True)``                              alabi/core.py:2325
* ``lnprob = like_fn(theta) + prior_fn(theta)``                        alabi/core.py:2073-2100
* ``get_autocorr_time(tol=0)`` / ``get_chain(discard, thin, flat)``    alabi/mcmc_utils.py:45, core.py:2341
"""
import numpy as np

from . import philox


def stretch_half_step(coords, logp, in_S, partner, zz, log_u, logp_fn):
    """One red/blue half update with explicit random draws.

    coords (n, d), logp (n,) are the CURRENT state; ``in_S`` marks the walkers
    being updated; for each of them ``partner`` indexes a walker of the
    complement, ``zz`` is the stretch factor and ``log_u`` the log accept
    uniform.  q = c - (c - s) z ; ln p_acc = (d-1) ln z + lp(q) - lp(s);
    accept iff ln p_acc > ln u (NaN compares False).  Returns the proposals,
    their log-probs and the accept mask (all indexed like the S walkers)."""
    idx = np.flatnonzero(in_S)
    ndim = coords.shape[1]
    s = coords[idx]
    c = coords[partner[idx]]
    q = c - (c - s) * zz[idx, None]
    lp_q = np.asarray(logp_fn(q), dtype=np.float64)
    with np.errstate(invalid="ignore"):
        lnpdiff = (ndim - 1.0) * np.log(zz[idx]) + lp_q - logp[idx]
        acc = lnpdiff > log_u[idx]
    return idx, q, lp_q, acc


def replay_device_chain(p0, logp_fn, nsteps, seed, a=2.0, randomize_split=True,
                        walker_offset=0, first_step=0):
    """Replay the device sampler's chain on the CPU from the same Philox draws.

    Returns chain (nsteps, n, d), logp (nsteps, n), accepted counts (n,) and
    the per-step proposal record (list of dicts) for per-proposal parity."""
    coords = np.array(p0, dtype=np.float64)
    n, d = coords.shape
    logp = np.asarray(logp_fn(coords), dtype=np.float64)
    chain = np.empty((nsteps, n, d))
    lps = np.empty((nsteps, n))
    nacc = np.zeros(n, dtype=np.int64)
    record = []
    for t in range(nsteps):
        dr = philox.move_draws(seed, n, first_step + t, randomize_split, walker_offset)
        zz = ((a - 1.0) * dr["u_z"] + 1.0) ** 2 / a
        with np.errstate(divide="ignore"):
            log_u = np.log(dr["u_acc"])
        prop_q = np.full((n, d), np.nan)
        prop_lp = np.full(n, np.nan)
        for split in (0, 1):
            in_S = (dr["sets"] == split) & (dr["partner"] >= 0)
            if not in_S.any():
                continue
            idx, q, lp_q, acc = stretch_half_step(coords, logp, in_S, dr["partner"], zz, log_u, logp_fn)
            prop_q[idx] = q
            prop_lp[idx] = lp_q
            coords[idx[acc]] = q[acc]
            logp[idx[acc]] = lp_q[acc]
            nacc[idx[acc]] += 1
        chain[t] = coords
        lps[t] = logp
        record.append(dict(q=prop_q, lp=prop_lp, **dr))
    return chain, lps, nacc, record


class StretchEnsemble:
    """emcee.EnsembleSampler restatement driven by ``numpy.random.RandomState``
    in emcee's own consumption order (shuffle of the label vector, ``rand(Ns)``
    for z, ``randint(Nc)`` for partners, one ``rand()`` per walker for the
    accept test).  Used as the CPU baseline and for posterior KS checks.

    ``vectorize=True`` hands the whole proposal block to ``log_prob_fn``
    (emcee's ``vectorize`` flag); otherwise one call per walker like the
    reference's ``lnprob`` path (alabi/core.py:2096)."""

    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, seed=None, vectorize=False):
        if nwalkers < 2 * ndim:
            raise ValueError("It is unadvisable to use a red-blue move with fewer walkers than twice the number of dimensions.")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn = log_prob_fn
        self.vectorize = vectorize
        self.random = np.random.RandomState(seed)
        self.chain = np.empty((0, nwalkers, ndim))
        self.log_prob = np.empty((0, nwalkers))
        self.naccepted = np.zeros(nwalkers)
        self.iteration = 0

    def _lp(self, q):
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(q), dtype=np.float64)
        else:
            lp = np.array([float(self.log_prob_fn(qi)) for qi in q])
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def run_mcmc(self, p0, nsteps):
        coords = np.array(p0, dtype=np.float64)
        logp = self._lp(coords)
        n = self.nwalkers
        chain = np.empty((nsteps, n, self.ndim))
        lps = np.empty((nsteps, n))
        all_inds = np.arange(n)
        for t in range(nsteps):
            inds = all_inds % 2
            self.random.shuffle(inds)
            for split in range(2):
                S1 = inds == split
                s = coords[S1]
                c = coords[~S1]
                Ns, Nc = len(s), len(c)
                zz = ((self.a - 1.0) * self.random.rand(Ns) + 1) ** 2.0 / self.a
                factors = (self.ndim - 1.0) * np.log(zz)
                rint = self.random.randint(Nc, size=(Ns,))
                q = c[rint] - (c[rint] - s) * zz[:, None]
                new_lp = self._lp(q)
                lnpdiff = factors + new_lp - logp[S1]
                acc = lnpdiff > np.log(self.random.rand(Ns))
                j = all_inds[S1][acc]
                coords[j] = q[acc]
                logp[j] = new_lp[acc]
                self.naccepted[j] += 1
            chain[t] = coords
            lps[t] = logp
        self.chain = np.concatenate([self.chain, chain])
        self.log_prob = np.concatenate([self.log_prob, lps])
        self.iteration += nsteps
        return coords

    def get_chain(self, discard=0, thin=1, flat=False):
        v = self.chain[discard + thin - 1::thin]
        return v.reshape(-1, self.ndim) if flat else v

    @property
    def acceptance_fraction(self):
        return self.naccepted / self.iteration


# ---------------------------------------------------------------------------
# emcee.autocorr restatement (Sokal windowing, c = 5)
# ---------------------------------------------------------------------------
def _acf_1d(x):
    n = 1
    while n < len(x):
        n <<= 1
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0]


def integrated_time(x, c=5, tol=50, quiet=False):
    """x: (nsteps, nwalkers, ndim) -> tau per dimension (walker-averaged ACF,
    tau(M) = 2 cumsum(acf) - 1 at the first M with M >= c tau(M))."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, None, None]
    if x.ndim == 2:
        x = x[:, :, None]
    nt, nw, nd = x.shape
    tau = np.empty(nd)
    for d in range(nd):
        f = np.zeros(nt)
        for k in range(nw):
            f += _acf_1d(x[:, k, d])
        f /= nw
        taus = 2.0 * np.cumsum(f) - 1.0
        m = np.arange(len(taus)) < c * taus
        window = int(np.argmin(m)) if np.any(m) else len(taus) - 1
        tau[d] = taus[window]
    if np.any(tol * tau > nt) and not quiet and tol > 0:
        raise RuntimeError("The chain is shorter than %d times the integrated autocorrelation time" % tol)
    return tau
