"""Development: time of both triangular sweeps + sums (ab_gp_log_likelihood on a factorised model)
and of the factorisation, per training-set size."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import alabi_b200 as ab
from alabi_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(5)
for n in (512, 1024, 2048, 4096, 8192, 16384):
    d = 4
    X = rng.uniform(0, 1, size=(n, d))
    y = np.sin(X.sum(axis=1) * 3.0)
    k = ab.kernels.ExpSquaredKernel(metric=np.full(d, 0.05), ndim=d) * np.var(y)
    g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-8.0, fit_white_noise=True)
    g.compute(X)
    h = g._hd.h
    yd = torch.from_numpy(y).cuda()
    ll = ctypes.c_double()

    def ev(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best
    t_solve = ev(lambda: lib.ab_gp_log_likelihood(h, _lib.ptr(yd), ctypes.byref(ll)))
    t_fac = ev(lambda: lib.ab_gp_factor(h), reps=3)
    print(json.dumps({"n": n, "solves_loglike_us": t_solve * 1e6, "us_per_block_step": t_solve * 1e6 / (2 * ((n + 127) // 128)),
                      "hbm_frac": 8.0 * n * n / t_solve * 1e-9 / 6458.1, "factor_ms": t_fac * 1e3, "loglike": ll.value}), flush=True)
