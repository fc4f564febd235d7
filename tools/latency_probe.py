"""Per-call latency of the one-point calls the acquisition polish and host samplers make."""
import os, sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200 import _lib, utility as ut

def per_call(f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6

for n, d in ((150, 2), (1000, 2), (4000, 2)):
    rng = np.random.default_rng(n)
    X = rng.uniform(-5, 5, size=(n, d)); y = -0.5 * np.sum((X / 2) ** 2, axis=1)
    g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y), fit_mean=True,
              mean=np.median(y), white_noise=-8.0, fit_white_noise=True)
    g.compute(X)
    x1 = rng.uniform(-4, 4, size=(1, d)); b = [(-5, 5)] * d
    g.predict(y, x1, return_var=True)
    hd = g._hd
    mu, var = np.empty(1), np.empty(1)
    xd = torch.from_numpy(x1).cuda(); mud = torch.empty(1, dtype=torch.float64, device="cuda"); vard = torch.empty_like(mud)
    print(f"N={n}: GP.predict(mean+var) {per_call(lambda: g.predict(y, x1, return_var=True)):.0f} us | "
          f"GP.predict(mean) {per_call(lambda: g.predict(y, x1, return_cov=False)):.0f} us | "
          f"C host call mean+var {per_call(lambda: hd.lib.ab_gp_predict_host(hd.h, _lib.ptr(x1), 1, _lib.ptr(mu), _lib.ptr(var))):.0f} us | "
          f"C host call mean {per_call(lambda: hd.lib.ab_gp_predict_host(hd.h, _lib.ptr(x1), 1, _lib.ptr(mu), None)):.0f} us | "
          f"C device call mean+var (no sync) {per_call(lambda: hd.lib.ab_gp_predict(hd.h, _lib.ptr(xd), 1, _lib.ptr(mud), _lib.ptr(vard))):.0f} us | "
          f"predict_grad {per_call(lambda: g.predict_grad(y, x1), 500):.0f} us | "
          f"bape_utility {per_call(lambda: ut.bape_utility(x1[0], lambda q: g.predict(y, q, return_var=True), b)):.0f} us")
