"""Development: per-task time stamps of the dataflow Cholesky (critical-path breakdown)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200 import _lib
lib = _lib.load()
n = int(os.environ.get("TRACE_N", "2048")); d = 10
rng = np.random.default_rng(n)
X = rng.uniform(-1, 1, size=(n, d)); y = -0.5 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y), fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
g.compute(X); h = g._hd.h
T = (n + 127) // 128; nt = T * (T + 1) // 2
buf = torch.zeros(nt * 6, dtype=torch.int64, device="cuda")
lib.ab_gp_debug_stamps(h, _lib.ptr(buf))
for _ in range(3): lib.ab_gp_factor(h)
torch.cuda.synchronize()
s = buf.cpu().numpy().reshape(nt, 6).astype(np.float64)
t0 = s[:, 0].min(); s = (s - t0) / 1e3
tasks = [(i, j) for j in range(T) for i in range(j, T)]
idx = {t: k for k, t in enumerate(tasks)}
print("total us", s[:, 5].max())
print("col | diag: pop accEnd staged potf2End pub | sub(j+1,j): accEnd flagSeen gemmEnd pub | dt(col)")
prev = None
for j in range(min(T, 14)):
    dg = s[idx[(j, j)]]
    line = f"{j:3d} | {dg[0]:7.1f} {dg[1]:7.1f} {dg[2]:7.1f} {dg[4]:7.1f} {dg[5]:7.1f}"
    if j + 1 < T:
        sb = s[idx[(j + 1, j)]]
        line += f" | {sb[1]:7.1f} {sb[3]:7.1f} {sb[4]:7.1f} {sb[5]:7.1f}"
    if prev is not None: line += f" | {dg[5] - prev:6.1f}"
    prev = dg[5]
    print(line)
