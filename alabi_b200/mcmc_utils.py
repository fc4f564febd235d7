"""MCMC post-processing (host side): burn-in / thinning heuristic of
alabi/mcmc_utils.py:15-72 and the integrated autocorrelation time that
``EnsembleSampler.get_autocorr_time`` needs (emcee.autocorr semantics: Sokal
window with c = 5 on the walker-averaged autocorrelation function)."""
import numpy as np

__all__ = ["estimate_burnin", "integrated_time", "AutocorrError"]


class AutocorrError(Exception):
    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super().__init__(*args, **kwargs)


def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def _autocorr_1d(x):
    x = np.asarray(x, dtype=np.float64)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - x.mean(), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    acf /= acf[0]
    return acf


def integrated_time(x, c=5, tol=50, quiet=False):
    """tau per parameter for a chain of shape (nsteps, nwalkers, ndim)."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, np.newaxis, np.newaxis]
    elif x.ndim == 2:
        x = x[:, :, np.newaxis]
    if x.ndim != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += _autocorr_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        m = np.arange(len(taus)) < c * taus
        windows[d] = np.argmin(m) if np.any(m) else len(taus) - 1
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag):
        msg = ("The chain is shorter than {0} times the integrated autocorrelation time for {1} parameter(s). "
               "Use this estimate with caution and run a longer chain!\n").format(tol, np.sum(flag))
        msg += "N/{0} = {1:.0f};\ntau: {2}".format(tol, n_t / tol, tau_est)
        if not quiet:
            raise AutocorrError(tau_est, msg)
    return tau_est


def estimate_burnin(sampler, est_burnin=True, thin_chains=True, verbose=False):
    """burn-in = int(2 max tau), thin = max(int(0.5 min tau), 1) with
    tau = sampler.get_autocorr_time(tol=0); non-finite taus are dropped and
    tau = 1 is used when nothing finite remains."""
    tau = np.atleast_1d(sampler.get_autocorr_time(tol=0))
    if np.any(~np.isfinite(tau)):
        tau = tau[np.isfinite(tau)]
        if len(tau) < 1:
            if verbose:
                print("Failed to compute integrated autocorrelation length, tau.")
                print("Setting tau = 1")
            tau = np.array([1.0])
    iburn = int(2.0 * np.max(tau)) if est_burnin else 0
    ithin = int(np.max((int(0.5 * np.min(tau)), 1))) if thin_chains else 1
    if verbose:
        print("burn-in estimate: %d" % iburn)
        print("thin estimate: %d\n" % ithin)
    return iburn, ithin
