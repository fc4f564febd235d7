"""Small-size launch of every kernel family for compute-sanitizer (tools/sanitize.sh):
ragged training-set sizes (not multiples of the 128-row blocks), several block rows for the
dataflow factorisation / solves, narrow and wide sampler units, resident and streamed
training sets, normal priors.  Sizes are small because the sanitizer slows kernels 10-100x."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

n = int(os.environ.get("TOUR_N", "333"))
d = int(os.environ.get("TOUR_D", "3"))
rng = np.random.default_rng(4)
for kind in ("ExpSquaredKernel", "Matern32Kernel", "Matern52Kernel"):
    X = rng.uniform(-1, 1, size=(n, d))
    y = -0.5 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
    g = ab.GP(kernel=getattr(ab.kernels, kind)(metric=np.full(d, 2.0), ndim=d) * np.var(y), fit_mean=True,
              mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    print(kind, "logL", g.log_likelihood(y), "grad", g.grad_log_likelihood(y)[:3])
    t = rng.uniform(-1.1, 1.1, size=(700, d))
    mu = g.predict(y, t, return_cov=False)
    mu, var = g.predict(y, t, return_var=True)                   # split variance kernel (few tiles)
    mu1, var1 = g.predict(y, t[:1], return_var=True)
    b = np.array([(-1.0, 1.0)] * d)
    for alg in ("bape", "agp", "jones"):
        print(alg, g.utility_argmin(y, t, b, algorithm=alg))
    _, _, dmu, dvar = g.predict_grad(y, t[:37])
    g.append_point(rng.uniform(-1, 1, size=d))
    y2 = np.append(y, -0.3)
    print("appended logL", g.log_likelihood(y2), g.predict(y2, t[:5], return_var=True)[1][:2])
    Kinv = g.solver.get_inverse()
    print("Kinv", Kinv.shape, float(np.abs(Kinv).max()))

# big-tile variance kernel (>= 74 query tiles) on a small factor
tq = rng.uniform(-1, 1, size=(74 * 128 + 5, d))
mu, var = g.predict(y2, tq, return_var=True)
print("var range", var.min(), var.max())
# device-buffer path (NumPy inputs above take the host-buffer entry point with its second stream)
mud, vard = g.predict(y2, torch.from_numpy(tq[:3000]).cuda(), return_var=True)
print("device predict", float(mud[0]), float(vard[0]))

# sampler: narrow units (resident training set), wide units, normal prior, streamed training set
lp = SurrogateLogProb(g, y2, b)
s = EnsembleSampler(64, d, lp, seed=1)
s.run_mcmc(rng.uniform(-0.9, 0.9, size=(64, d)), 4, record_proposals=True)
lpn = SurrogateLogProb(g, y2, b, prior_data=[(0.1, 0.3)] + [(None, None)] * (d - 1))
s = EnsembleSampler(4800, d, lpn, seed=2)
s.run_mcmc(rng.uniform(-0.9, 0.9, size=(4800, d)), 2)
print("ensemble acc", s.acceptance_fraction.mean())
if os.environ.get("TOUR_STREAMED", "1") == "1":
    n3, d3 = 2100, 12                       # 2100 * 13 * 8 B > 160 KB: cp.async chunk ring
    X3 = rng.uniform(-1, 1, size=(n3, d3))
    y3 = -0.5 * np.sum(X3 ** 2, axis=1)
    g3 = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d3, 4.0), ndim=d3) * np.var(y3), fit_mean=True,
               mean=np.median(y3), white_noise=-6.0, fit_white_noise=True)
    g3.compute(X3)
    print("logL3", g3.log_likelihood(y3), g3.grad_log_likelihood(y3)[:2])
    s = EnsembleSampler(48, d3, SurrogateLogProb(g3, y3, [(-1, 1)] * d3), seed=3)
    s.run_mcmc(rng.uniform(-0.9, 0.9, size=(48, d3)), 2)
    print("streamed ensemble acc", s.acceptance_fraction.mean())
torch.cuda.synchronize()
print("TOUR DONE")
