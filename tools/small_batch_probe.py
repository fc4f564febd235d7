"""Development probe: latency of small-batch predict (mean+var) and predict_grad calls."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
for n, d in ((1000, 2), (4000, 2), (8192, 10)):
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, size=(n, d)); y = -0.5 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
    g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 1.0), ndim=d) * np.var(y), fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    for m in (1, 64, 1024):
        t = rng.uniform(-1, 1, size=(m, d))
        g.predict(y, t, return_var=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): g.predict(y, t, return_var=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        g.predict_grad(y, t); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): g.predict_grad(y, t)
        torch.cuda.synchronize(); dg = (time.perf_counter() - t0) / 10
        print(f"N={n} d={d} M={m}: predict(mean+var) {dt*1e3:.3f} ms, predict_grad {dg*1e3:.3f} ms", flush=True)
