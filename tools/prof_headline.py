"""The two headline kernels at bench.py's sizes, one launch each, for `ncu --set full`
(`tools/gpu_round2.sh`): predict_var_kernel on one 18944-query panel of the c4 model and the
wide sampler kernel on a c5 sub-ensemble (8192 walkers, 4 steps)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alabi_b200 import workloads
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

which = sys.argv[1] if len(sys.argv) > 1 else "both"
rng = np.random.default_rng(0)
if which in ("both", "c4"):
    c4 = workloads.make_config("c4")
    g = workloads.build_gp(c4)
    g.compute(c4["X"])
    t = torch.from_numpy(rng.uniform(-3, 3, size=(148 * 128, 10))).cuda()
    for _ in range(2):
        mu, var = g.predict(c4["y"], t, return_var=True)
    torch.cuda.synchronize()
    print("c4 predict", float(mu[0]), float(var[0]))
    del g
if which in ("both", "c5"):
    c5 = workloads.make_config("c5")
    g5 = workloads.build_gp(c5)
    g5.compute(c5["X"])
    s = EnsembleSampler(8192, 20, SurrogateLogProb(g5, c5["y"], c5["bounds"]), seed=1)
    s.run_mcmc(rng.uniform(-1, 1, size=(8192, 20)), 4, store=False)
    torch.cuda.synchronize()
    print("c5 sampler acc", s.acceptance_fraction.mean(), "device s", s.last_run_device_seconds)
