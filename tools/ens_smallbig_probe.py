"""Development: small ensembles (emcee's default 10 x ndim walkers) on LARGE training sets, where the
training points do not fit shared memory and every unit streams them."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb
rng = np.random.default_rng(3)
for n, d, nw in ((2000, 10, 100), (8192, 10, 100), (8192, 10, 400), (16384, 20, 200), (4000, 4, 100), (1900, 10, 100)):
    X = rng.uniform(0, 1, size=(n, d))
    y = -0.5 * np.sum(((X - 0.5) / 0.2) ** 2, axis=1)
    k = ab.kernels.ExpSquaredKernel(metric=np.full(d, 0.3 * d), ndim=d) * np.var(y)
    g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    lp = SurrogateLogProb(g, y, [(0, 1)] * d)
    p0 = rng.uniform(0.3, 0.7, size=(nw, d))
    for sched in (3, 0):        # 3: 2-proposal units streaming the whole training set (the previous path), 0: automatic (spread)
        s = EnsembleSampler(nw, d, lp, seed=1, schedule=sched)
        s.run_mcmc(p0, 20, store=False)
        st = s.run_mcmc(None, 500, store=False)
        print(json.dumps({"n": n, "d": d, "nw": nw, "schedule": sched, "us_per_step": s.last_run_device_seconds / 500 * 1e6,
                          "walker_steps_per_s": nw * 500 / s.last_run_device_seconds, "state_sum": float(st.coords.sum())}), flush=True)
