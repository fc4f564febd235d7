"""Development: time of the covariance build inside ab_gp_factor (profiling family 0) at the c4 / c5 sizes
and a 2-D case."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import alabi_b200 as ab
from alabi_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(5)
for kind, n, d in (("ExpSquaredKernel", 8192, 10), ("ExpSquaredKernel", 16384, 20), ("Matern32Kernel", 8192, 2), ("ExpSquaredKernel", 4096, 4)):
    X = rng.uniform(0, 1, size=(n, d))
    y = np.sin(X.sum(axis=1) * 3.0)
    k = getattr(ab.kernels, kind)(metric=np.full(d, 0.05 * d), ndim=d) * np.var(y)
    g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    h = g._hd.h
    lib.ab_gp_set_profiling(h, 1)
    for _ in range(5):
        lib.ab_gp_factor(h)
    torch.cuda.synchronize()
    cms, ccnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h, 0, ctypes.byref(cms), ctypes.byref(ccnt))
    lib.ab_gp_set_profiling(h, 0)
    t = cms.value * 1e-3 / max(ccnt.value, 1)
    print(json.dumps({"kind": kind, "n": n, "d": d, "cov_build_us": t * 1e6, "lower_triangle_gb_s": 4.0 * n * n / t * 1e-9,
                      "frac_hbm": 4.0 * n * n / t * 1e-9 / 6458.1, "launches": ccnt.value}), flush=True)
