"""Development: per-phase cycle counters of the spread mode (small ensemble, large training set)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb
rng = np.random.default_rng(3)
for n, d, nw in ((8192, 10, 100), (16384, 20, 200), (2000, 10, 100)):
    X = rng.uniform(0, 1, size=(n, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.2) ** 2, axis=1)
    k = ab.kernels.ExpSquaredKernel(metric=np.full(d, 0.3 * d), ndim=d) * np.var(y)
    g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    lp = SurrogateLogProb(g, y, [(0, 1)] * d)
    s = EnsembleSampler(nw, d, lp, seed=1)
    s.debug_timing = 1
    s.run_mcmc(rng.uniform(0.3, 0.7, size=(nw, d)), 20, store=False)
    s.run_mcmc(None, 500, store=False)
    print(json.dumps({"n": n, "d": d, "nw": nw, "us_per_step": s.last_run_device_seconds / 500 * 1e6}), flush=True)
