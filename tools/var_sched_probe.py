"""Variance-GEMM schedules on full panels (ab_gp_set_variance_schedule): one CTA per query tile
against row-block pairs on adjacent CTAs; time per 18944-query wave and bit identity, c4 and c3."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alabi_b200 import _lib, workloads
lib = _lib.load()
for name in (sys.argv[1:] or ["c4", "c3"]):
    cfg = workloads.make_config(name)
    g = workloads.build_gp(cfg)
    g.compute(cfg["X"])
    m = 148 * 128 * 2
    b = cfg["bounds"]
    t = torch.from_numpy(np.random.default_rng(0).uniform(b[:, 0], b[:, 1], size=(m, cfg["ndim"]))).cuda()
    out = {}
    for mode in (1, 2):
        lib.ab_gp_set_variance_schedule(g._handle().h, mode)
        mu, var = g.predict(cfg["y"], t, return_var=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mu, var = g.predict(cfg["y"], t, return_var=True)
        e1.record(); e1.synchronize()
        out[mode] = (e0.elapsed_time(e1) / 3, var.clone())
    print(f"{name}: N={len(cfg['X'])} {m} queries: one-CTA-per-tile {out[1][0]:.3f} ms, row-block pairs {out[2][0]:.3f} ms, "
          f"identical bits {bool(torch.equal(out[1][1], out[2][1]))}", flush=True)
