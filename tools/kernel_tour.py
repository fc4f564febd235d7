"""One launch of every kernel family at a representative size, for ncu captures
(`tools/ncu_tour.sh`).  Sizes: c4-like GP (N = 8192, d = 10, ExpSquared) for K1-K4 and
the gradient; c2 GP (N = 1000, d = 2, Matern-3/2) for the sampler."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

n = int(os.environ.get("TOUR_N", "8192"))
d = 10
rng = np.random.default_rng(4)
X = rng.uniform(-1, 1, size=(n, d))
y = -0.5 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y), fit_mean=True,
          mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
g.compute(X)                                   # K1 cov_kernel + K2 chol_dataflow_kernel
print("logL", g.log_likelihood(y))             # trsv_coop_kernel
print("grad", g.grad_log_likelihood(y)[:3])    # trinv, kinv, grad_tiles
t = rng.uniform(-1, 1, size=(1 << 18, d))
mu = g.predict(y, t, return_cov=False)         # predict_mean_kernel (mean only)
mu, var = g.predict(y, t[:148 * 128 * 2], return_var=True)     # panel + predict_var_kernel
idx, val = g.utility_argmin(y, t[:148 * 128 * 2], np.array([(-1.0, 1.0)] * d), algorithm="bape")
print("argmin", idx, val)
_, _, dmu, dvar = g.predict_grad(y, t[:256])    # tri_gemm_kernel x2 + predict_grad_kernel
print("grad", dmu[0, :2], dvar[0, :2])
g.append_point(rng.uniform(-1, 1, size=d))      # bordered update: forward trsv_dataflow + row kernels

X2 = rng.uniform(-6, 6, size=(1000, 2))
y2 = -0.5 * np.sum((X2 / 2.0) ** 2, axis=1)
g2 = ab.GP(kernel=ab.kernels.Matern32Kernel(metric=np.full(2, 9.0), ndim=2) * np.var(y2), fit_mean=True,
           mean=np.median(y2), white_noise=-8.0, fit_white_noise=True)
g2.compute(X2)
s = EnsembleSampler(1000, 2, SurrogateLogProb(g2, y2, [(-6, 6), (-6, 6)]), seed=1)
s.run_mcmc(rng.uniform(-5, 5, size=(1000, 2)), 200, store=True)
print("ensemble acc", s.acceptance_fraction.mean())
torch.cuda.synchronize()
