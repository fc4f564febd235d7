"""Where init_gp's first seconds go: CUDA context, library load, first launches, hyper-opt."""
import os, sys, time, tempfile
t0 = time.time()
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t1 = time.time(); print(f"import numpy+torch {t1 - t0:.2f}s")
torch.zeros(1, device="cuda"); torch.cuda.synchronize(); t2 = time.time(); print(f"CUDA context {t2 - t1:.2f}s")
import alabi_b200 as ab
from alabi_b200 import benchmarks as bm, _lib
_lib.load(); t3 = time.time(); print(f"import alabi_b200 + dlopen {t3 - t2:.2f}s")
rng = np.random.default_rng(0)
X = rng.uniform(-5, 5, size=(50, 2)); y = np.array([bm.rosenbrock["fn"](x) for x in X])
g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.ones(2), ndim=2) * np.var(y), fit_mean=True, mean=np.median(y),
          white_noise=-12, fit_white_noise=False)
g.compute(X); t4 = time.time(); print(f"first GP.compute (N=50) {t4 - t3:.2f}s")
ll = g.log_likelihood(y); t5 = time.time(); print(f"first log_likelihood {t5 - t4:.3f}s")
gr = g.grad_log_likelihood(y); t6 = time.time(); print(f"first grad_log_likelihood {t6 - t5:.3f}s")
mu, var = g.predict(y, X[:1], return_var=True); t7 = time.time(); print(f"first predict {t7 - t6:.3f}s")
p = g.get_parameter_vector()
t = time.time()
for i in range(100):
    g.set_parameter_vector(p + 1e-3 * i); g.log_likelihood(y, quiet=True); g.grad_log_likelihood(y, quiet=True)
print(f"100 x (set + logL + grad) at N=50: {(time.time() - t) * 10:.2f} ms each")
np.random.seed(1)
for hyp in ("ml", "cv"):
    sm = ab.SurrogateModel(lnlike_fn=bm.rosenbrock["fn"], bounds=bm.rosenbrock["bounds"], savedir=tempfile.mkdtemp(), cache=False, verbose=False)
    sm.init_samples(ntrain=50, ntest=20)
    t = time.time()
    sm.init_gp(kernel="ExpSquaredKernel", fit_amp=True, fit_mean=True, white_noise=-12, hyperopt_method=hyp, gp_nopt=3)
    print(f"init_gp({hyp}) warm: {time.time() - t:.2f}s")
