"""A handful of one-point predict / predict_grad calls at N = 4000 for an ncu launch list of the
few-query kernels (few_cross / few_gemv / few_finish / few_gemv_t / few_grad)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
n, d = int(os.environ.get("FEW_N", "4000")), 2
rng = np.random.default_rng(n)
X = rng.uniform(-5, 5, size=(n, d)); y = -0.5 * np.sum((X / 2) ** 2, axis=1)
g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y), fit_mean=True,
          mean=np.median(y), white_noise=-8.0, fit_white_noise=True)
g.compute(X)
x1 = rng.uniform(-4, 4, size=(1, d))
for _ in range(4):
    mu, var = g.predict(y, x1, return_var=True)
    out = g.predict_grad(y, x1)
torch.cuda.synchronize()
print("few probe", float(mu[0]), float(var[0]), out[2][0])
