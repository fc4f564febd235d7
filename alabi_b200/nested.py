"""Batched nested sampler used by ``SurrogateModel.run_dynesty`` when dynesty is
not importable.

The reference hands ``self.like_fn`` to ``dynesty.(Dynamic)NestedSampler`` and
reads ``results.{samples, logwt, logz, logzerr, niter}`` plus
``dynesty.utils.resample_equal`` (alabi/core.py:2629-2706); dynesty calls the
likelihood one point at a time.  This sampler exposes the same result fields
and keeps the GPU busy instead: replacement points are produced by many
constrained random-walk chains advanced in lock-step, so every likelihood call
is ONE batched surrogate predict (K3) over all chains.

Algorithm: static nested sampling (Skilling 2006) with ln X_i = -i / nlive,
replacement by ``walks`` Metropolis steps from a random live point inside the
hard constraint L > L_min (the "rwalk" scheme of dynesty), proposal scale
adapted towards 50 % acceptance; chains are generated ``nbatch`` at a time and
consumed while they still satisfy the current constraint.
"""
import numpy as np

__all__ = ["BatchedNestedSampler", "NestedResults", "resample_equal"]


class NestedResults(dict):
    __getattr__ = dict.get


def resample_equal(samples, weights, rstate=None):
    """Systematic resampling to equally weighted samples (dynesty.utils.resample_equal)."""
    rstate = np.random.default_rng() if rstate is None else rstate
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    n = len(w)
    positions = (rstate.random() + np.arange(n)) / n
    idx = np.minimum(np.searchsorted(np.cumsum(w), positions), n - 1)
    out = np.asarray(samples)[idx]
    return out[rstate.permutation(n)]


class BatchedNestedSampler:
    def __init__(self, loglike_batch, prior_transform, ndim, nlive=500, walks=25, nbatch=None, rstate=None,
                 **unused):
        self.loglike = loglike_batch          # (m, ndim) -> (m,)
        self.prior_transform = prior_transform
        self.ndim = int(ndim)
        self.nlive = int(nlive)
        self.walks = int(walks)
        self.nbatch = int(nbatch) if nbatch else max(self.nlive // 4, 8)
        self.rng = np.random.default_rng(rstate)
        self.results = None
        self.ncall = 0

    def _like(self, u):
        theta = np.atleast_2d(self.prior_transform(u))
        self.ncall += len(theta)
        return theta, np.asarray(self.loglike(theta), dtype=np.float64).reshape(-1)

    def _new_points(self, live_u, live_l, lmin, scale):
        """nbatch constrained random walks advanced together; returns the end
        points (u, theta, logl) and the mean acceptance."""
        k, d = self.nbatch, self.ndim
        start = self.rng.integers(0, self.nlive, size=k)
        u = live_u[start].copy()
        logl = live_l[start].copy()
        theta = np.atleast_2d(self.prior_transform(u)).copy()
        cov = np.cov(live_u.T) + 1e-12 * np.eye(d) if d > 1 else np.atleast_2d(np.var(live_u) + 1e-12)
        chol = np.linalg.cholesky(cov)
        nacc = 0
        for _ in range(self.walks):
            prop = u + scale * (self.rng.standard_normal((k, d)) @ chol.T)
            inside = np.all((prop > 0.0) & (prop < 1.0), axis=1)
            pl = np.full(k, -np.inf)
            pt = theta.copy()
            if inside.any():
                t_in, l_in = self._like(prop[inside])
                pl[inside] = l_in
                pt[inside] = t_in
            ok = inside & (pl > lmin)
            u[ok], logl[ok], theta[ok] = prop[ok], pl[ok], pt[ok]
            nacc += ok.sum()
        return u, theta, logl, nacc / float(k * self.walks)

    def run_nested(self, dlogz=0.01, maxiter=None, maxcall=None, print_progress=False, **unused):
        nl, d = self.nlive, self.ndim
        live_u = self.rng.random((nl, d))
        live_t, live_l = self._like(live_u)
        samples, logls, logwts, logzs, logzerrs = [], [], [], [], []
        logz, h, logvol = -1e300, 0.0, 0.0
        dlv = np.log1p(-np.exp(-1.0 / nl))                 # ln(1 - e^{-1/nlive})
        scale, queue = 1.0, []
        it = 0
        maxiter = int(maxiter) if maxiter else 10 ** 9
        while it < maxiter:
            worst = int(np.argmin(live_l))
            lmin = live_l[worst]
            logwt = logvol + dlv + lmin                    # L_i * (X_{i-1} - X_i)
            logz_new = np.logaddexp(logz, logwt)
            if np.isfinite(lmin):
                h = (np.exp(logwt - logz_new) * lmin + np.exp(logz - logz_new) * (h + logz) - logz_new)
            logz = logz_new
            samples.append(live_t[worst].copy())
            logls.append(lmin)
            logwts.append(logwt)
            logzs.append(logz)
            logzerrs.append(np.sqrt(max(h, 0.0) / nl))
            logvol -= 1.0 / nl
            it += 1
            # replacement point
            new = None
            while new is None:
                while queue:
                    cu, ct, cl = queue.pop()
                    if cl > lmin:
                        new = (cu, ct, cl)
                        break
                if new is None:
                    qu, qt, ql, acc = self._new_points(live_u, live_l, lmin, scale)
                    scale = float(np.clip(scale * np.exp((acc - 0.5) / max(d, 1)), 1e-6, 10.0))
                    queue = [(qu[i], qt[i], ql[i]) for i in range(len(ql)) if ql[i] > lmin]
                    if maxcall and self.ncall > maxcall:
                        break
            if new is None:
                break
            live_u[worst], live_t[worst], live_l[worst] = new
            # remaining evidence in the live points
            dz = np.logaddexp(logz, np.max(live_l) + logvol) - logz
            if print_progress and it % 500 == 0:
                print(f"iter {it} logz {logz:.3f} dlogz {dz:.4f} ncall {self.ncall}")
            if dz < dlogz:
                break
        # add the final live points, each with X_final / nlive
        order = np.argsort(live_l)
        for i in order:
            logwt = logvol - np.log(nl) + live_l[i]
            logz_new = np.logaddexp(logz, logwt)
            h = (np.exp(logwt - logz_new) * live_l[i] + np.exp(logz - logz_new) * (h + logz) - logz_new)
            logz = logz_new
            samples.append(live_t[i].copy())
            logls.append(live_l[i])
            logwts.append(logwt)
            logzs.append(logz)
            logzerrs.append(np.sqrt(max(h, 0.0) / nl))
        self.results = NestedResults(samples=np.array(samples), logl=np.array(logls), logwt=np.array(logwts),
                                     logz=np.array(logzs), logzerr=np.array(logzerrs), niter=it,
                                     ncall=self.ncall, eff=100.0 * it / max(self.ncall, 1), nlive=nl)
        return self.results
