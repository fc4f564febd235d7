"""Where the end-to-end time of run_mcmc(store=True) goes at the c2 size (1000 walkers x 5000
steps, 120 MB of chain + log-prob): device time, host staging, copies."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

rng = np.random.default_rng(2)
X = rng.uniform(-6, 6, size=(1000, 2))
y = -0.5 * np.sum((X / 2.0) ** 2, axis=1)
g = ab.GP(kernel=ab.kernels.Matern32Kernel(metric=np.full(2, 9.0), ndim=2) * np.var(y), fit_mean=True,
          mean=np.median(y), white_noise=-8.0, fit_white_noise=True)
g.compute(X)
lp = SurrogateLogProb(g, y, [(-6, 6), (-6, 6)])
p0 = rng.uniform(-5, 5, size=(1000, 2))
EnsembleSampler(1000, 2, lp, seed=1).run_mcmc(p0, 100, store=False)
torch.cuda.synchronize()
t = time.perf_counter(); h = torch.empty((5000, 1000, 2), dtype=torch.float64, pin_memory=True); t1 = time.perf_counter() - t
print(f"fresh pinned 80 MB alloc: {t1 * 1e3:.1f} ms")
d = torch.empty((5000, 1000, 2), dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
t = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); print(f"D2H 80 MB pinned: {(time.perf_counter() - t) * 1e3:.1f} ms")
del h, d
for rep in range(3):
    s = EnsembleSampler(1000, 2, lp, seed=10 + rep)
    torch.cuda.synchronize(); t = time.perf_counter()
    s.run_mcmc(p0, 5000)
    el = time.perf_counter() - t
    print(f"run {rep}: e2e {el * 1e3:.1f} ms ({5e6 / el / 1e6:.1f} M walker-steps/s), device {s.last_run_device_seconds * 1e3:.1f} ms, "
          f"chain {s.get_chain().shape}")
    ch = s.get_chain().copy()
    del s
s2 = EnsembleSampler(1000, 2, lp, seed=12)
s2.run_mcmc(p0, 2500); s2.run_mcmc(None, 2500)
print("two halves equal one run:", np.array_equal(s2.get_chain(), ch))
