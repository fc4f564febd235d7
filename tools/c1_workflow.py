"""BASELINE config c1 end to end on the GPU path (timing of the alabi-shaped workflow):
2-D Rosenbrock, ExpSquared, 50 initial samples + 100 BAPE iterations, emcee with 100 walkers."""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200 import benchmarks as bm
np.random.seed(1)
t0 = time.time()
sm = ab.SurrogateModel(lnlike_fn=bm.rosenbrock["fn"], bounds=bm.rosenbrock["bounds"], savedir=tempfile.mkdtemp(), cache=False, verbose=False)
sm.init_samples(ntrain=50, ntest=20)
t1 = time.time()
sm.init_gp(kernel="ExpSquaredKernel", fit_amp=True, fit_mean=True, white_noise=-12, hyperopt_method=os.environ.get("C1_HYPEROPT", "ml"), gp_nopt=3)
t2 = time.time()
sm.active_train(niter=100, algorithm="bape", gp_opt_freq=20, show_progress=False)
t3 = time.time()
sm.run_emcee(nwalkers=100, nsteps=5000)
t4 = time.time()
tr = sm.training_results
print(f"init_samples {t1-t0:.2f}s  init_gp {t2-t1:.2f}s  active_train(100) {t3-t2:.2f}s  run_emcee(100x5000) {t4-t3:.2f}s")
print(f"per iteration: fit {np.mean(tr['gp_train_time'])*1e3:.2f} ms, acquisition {np.mean(tr['obj_fn_opt_time'])*1e3:.1f} ms; "
      f"final test MSE {tr['test_mse'][-1]:.3e}; N = {sm.ntrain}; emcee samples {sm.emcee_samples.shape}")
