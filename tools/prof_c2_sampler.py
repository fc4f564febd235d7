"""The small-ensemble sampler at the c2 sizes (1000 walkers, N = 1000, Matern-3/2), one short run per schedule,
for `ncu --set full` (tools/gpu_ncu_c2.sh)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alabi_b200 import workloads
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb
c2 = workloads.make_config("c2")
g = workloads.build_gp(c2)
g.compute(c2["X"])
lp = SurrogateLogProb(g, c2["y"], c2["bounds"])
rng = np.random.default_rng(0)
b = np.asarray(c2["bounds"], dtype=float)
p0 = rng.uniform(b[:, 0], b[:, 1], size=(1000, 2))
for sched in (1, 0):
    s = EnsembleSampler(1000, 2, lp, seed=1, schedule=sched)
    s.run_mcmc(p0, 400, store=False)
    torch.cuda.synchronize()
    print("schedule", sched, "us per step", s.last_run_device_seconds / 400 * 1e6)
