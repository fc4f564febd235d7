"""Default k-fold CV hyper-parameter search of init_gp (100 + 50 + 25 candidates x 5 folds,
alabi/gp_utils.py:640-1231) on the c1 and c2 training sets: the batched device job against the
one-by-one device path (same candidates, same folds)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alabi_b200 import gp_utils, utility as ut, workloads

for name, n in (("c1", 150), ("c2", 1000), ("c3", 2000)):
    cfg = workloads.make_config(name, n=n)
    X, y = cfg["X"], cfg["y"]
    g = workloads.build_gp(cfg)
    g.compute(X)
    base = g.get_parameter_vector()
    rng = np.random.default_rng(1)
    lo = base - np.array([np.std(y), 3.0, 2.0] + [2.0] * cfg["ndim"])
    hi = base + np.array([np.std(y), 3.0, 2.0] + [2.0] * cfg["ndim"])
    cands = rng.uniform(lo, hi, size=(100, len(base)))
    cands[0] = base
    out = {}
    for batched in (True, False, True):
        np.random.seed(3)
        torch.cuda.synchronize()
        t0 = time.time()
        r = gp_utils.optimize_gp_kfold_cv(g, X, y, cands, ut.no_scaler, k_folds=5, scoring="mse", stage2_candidates=50,
                                          stage3_candidates=25, verbose=False, batched=batched, random_state=11)
        torch.cuda.synchronize()
        out[batched] = (time.time() - t0, r.get_parameter_vector())
    print(f"{name} N={n}: 875 jobs batched {out[True][0]:.3f} s, one by one {out[False][0]:.3f} s, "
          f"speed-up {out[False][0] / out[True][0]:.1f}x, same winner {np.allclose(out[True][1], out[False][1])}", flush=True)

# the batched call alone: 100 candidates x 5 folds on the c2 training set
from sklearn.model_selection import KFold
cfg = workloads.make_config("c2")
X, y = cfg["X"], cfg["y"]
g = workloads.build_gp(cfg)
base = g.get_parameter_vector()
cands = base + np.random.default_rng(2).normal(0, 0.2, size=(100, len(base)))
folds = [f for c in range(100) for f in KFold(n_splits=5, shuffle=True, random_state=c).split(X)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    preds, ll, st = g.cv_batch(X, y, cands, folds)
    torch.cuda.synchronize(); print(f"cv_batch 500 jobs, nt = 800: {(time.time() - t0) * 1e3:.1f} ms, ok {int((st == 0).sum())}", flush=True)
