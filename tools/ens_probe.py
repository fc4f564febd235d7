import os, sys, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb
rng = np.random.default_rng(2)
X = rng.uniform(-6, 6, size=(1000, 2))
y = -0.5 * np.sum((X / 2.0) ** 2, axis=1)
k = ab.kernels.Matern32Kernel(metric=np.full(2, 9.0), ndim=2) * np.var(y)
g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-8.0, fit_white_noise=True)
g.compute(X)
lp = SurrogateLogProb(g, y, [(-6, 6), (-6, 6)])
for wpu in (0, 4, 8):
    for dbg in (1, 0, 4):
        s = EnsembleSampler(1000, 2, lp, seed=1, warps_per_unit=wpu)
        s.debug_timing = dbg
        s.run_mcmc(rng.uniform(-5, 5, size=(1000, 2)), 50, store=False)
        s.run_mcmc(None, 2000, store=False)
        print(json.dumps({"wpu": wpu, "dbg": dbg, "walker_steps_per_s": 1000 * 2000 / s.last_run_device_seconds}), flush=True)

# c5-like per-GPU share: N = 16384, d = 20, 8192 walkers
import time
d, n, nw = 20, 16384, int(os.environ.get('ENS_NW', '8192'))
X = rng.uniform(0, 1, size=(n, d))
y = -0.5 * np.sum(((X - 0.5) / 0.2) ** 2, axis=1)
k = ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y)
g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
g.compute(X)
lp = SurrogateLogProb(g, y, [(0, 1)] * d)
for pp in (4, 32, 0):
    s = EnsembleSampler(nw, d, lp, seed=1)
    s.debug_timing = pp
    s.run_mcmc(rng.uniform(0.3, 0.7, size=(nw, d)), 2, store=False)
    s.run_mcmc(None, 20, store=False)
    ws = nw * 20 / s.last_run_device_seconds
    print(json.dumps({"c5_like": True, "P": pp, "walker_steps_per_s": ws, "fp64_instr_rate_frac": ws * n * (2 * d + 22) / 1.8e13,
                      "acc": float(s.acceptance_fraction.mean())}), flush=True)
