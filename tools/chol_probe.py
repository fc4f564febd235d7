"""Development probe (GPU box): Cholesky / inverse timings only."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200 import _lib
from gpu_probe import ev_time

lib = _lib.load()
for n in [int(s) for s in os.environ.get("PROBE_SIZES", "2048,4096,8192,16384").split(",")]:
    d = int(os.environ.get("PROBE_D", "10"))
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, size=(n, d))
    y = -0.5 * np.sum(X ** 2, axis=1) + 0.01 * rng.normal(size=n)
    k = ab.kernels.ExpSquaredKernel(metric=np.full(d, 4.0), ndim=d) * np.var(y)
    g = ab.GP(kernel=k, fit_mean=True, mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
    g.compute(X)
    h = g._hd.h
    res = {"n": n}
    Ls = {}
    for la in (1, 2):
        lib.ab_gp_set_lookahead(h, la)
        rc = lib.ab_gp_factor(h)
        if rc != 0:
            res[f"mode{la}_rc"] = rc
            continue
        t = ev_time(lambda: lib.ab_gp_factor(h), reps=5)
        res[f"mode{la}_ms"] = round(t * 1e3, 3)
        res[f"mode{la}_tflops"] = round(n ** 3 / 3 / t * 1e-12, 2)
        g._targets_pushed = False
        g._set_targets(y)
        L, alpha = g.export_state()
        Ls[la] = (L[:n, :n].clone(), alpha.clone())
    import ctypes
    lib.ab_gp_set_lookahead(h, 2)
    lib.ab_gp_set_profiling(h, 1)
    for _ in range(3):
        lib.ab_gp_factor(h)
    torch.cuda.synchronize()
    ms, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h, 0, ctypes.byref(ms), ctypes.byref(cnt))
    res["cov_lower_ms"] = round(ms.value / max(cnt.value, 1), 4)
    res["cov_lower_gbs"] = round(4.0 * n * n / (ms.value / max(cnt.value, 1) * 1e-3) * 1e-9, 1)
    lib.ab_gp_set_profiling(h, 0)
    if len(Ls) == 2:
        res["max_abs_dL"] = float((Ls[1][0] - Ls[2][0]).abs().max())
        res["rel_dalpha"] = float((Ls[1][1] - Ls[2][1]).abs().max() / Ls[1][1].abs().max())
    print(json.dumps(res), flush=True)
    del g
