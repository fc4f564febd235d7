#!/usr/bin/env python
"""Headline benchmark of the alabi surrogate hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): 2-D Gaussian shells, Matern-3/2
kernel, N = 1000 training points.  One "step" = the hot path over one batch:
GP predictive mean + variance at M query points per GPU (K3: on-the-fly cross
covariance + DMMA GEMM with L^-1).  ``value`` = points/s over all GPUs with the
queries resident in HBM; ``e2e`` = the same through the public API
(``GP.predict`` on host NumPy buffers: pinned H2D of the queries and D2H of mean
and variance inside the timed region).  The second half of the metric,
surrogate-MCMC walker-steps/s (K5, 1000 walkers per GPU), is reported under
``mcmc``.  The GP is trained once on rank 0 and L / alpha are broadcast (NCCL);
afterwards every rank works on its own queries / sub-ensemble (weak scaling,
no data-path collective).

``kernels`` (N = 1 only) adds the covariance build, the Cholesky, both solves and the
log-likelihood gradient at the training-set sizes of configs c4 (N = 8192, d = 10)
and c5 (N = 16384, d = 20), each against its roof, and ``one_point_calls_c2``: the
end-to-end latency of the reference's M = 1 calls (``GP.predict`` / ``GP.predict_grad`` on one
host point) on the c2 model.

``--impl reference`` times the CPU oracle (NumPy/SciPy restatement of the
george + emcee path, all host BLAS threads) on a bounded sample of the same
workload and prints the same JSON line with ``"impl": "reference"``.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gp_predict_points_per_sec_mean_var"
UNIT = "points/s"
N_TRAIN, NDIM = 1000, 2
M_PER_STEP = 1 << 20            # query points per GPU per step
N_ROTATE = 10                   # distinct query batches (10 x 16 MiB inputs + 160 MiB outputs > L2)
NWALKERS, MCMC_STEPS = 1000, 5000


def shells(x, r=2.0, w=0.1, c=3.5):
    """Gaussian shells log-density (restated from alabi/benchmarks.py:100-113)."""
    const = np.log(1.0 / np.sqrt(2.0 * np.pi * w ** 2))
    s = lambda cx: const - (np.sqrt((x[:, 0] - cx) ** 2 + x[:, 1] ** 2) - r) ** 2 / (2.0 * w ** 2)
    return np.logaddexp(s(-c), s(c))


def workload():
    rng = np.random.default_rng(2)
    X = rng.uniform(-6.0, 6.0, size=(N_TRAIN, NDIM))
    y = shells(X)
    hp = dict(log_M=np.log(np.full(NDIM, 0.5 ** 2)), amp=float(np.var(y)), mean=float(np.median(y)), white_noise=-12.0)
    return X, y, hp, [(-6.0, 6.0)] * NDIM


def config_dict(n_gpus):
    return {"workload": "c2: 2-D Gaussian shells, Matern-3/2, N=1000; predict mean+var over "
                        f"{M_PER_STEP} query points per GPU per step; mcmc {NWALKERS} walkers x {MCMC_STEPS} steps per GPU",
            "n_train": N_TRAIN, "ndim": NDIM, "kernel": "Matern32Kernel", "points_per_step_per_gpu": M_PER_STEP,
            "l2_policy": f"inputs rotate over {N_ROTATE} distinct query batches (> 126 MB L2 with outputs)",
            "parallelism": f"replicated factor, queries/walkers sharded x{n_gpus}"}


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle sampling during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_reference_run(steps, warmup, sample_points=None, mcmc_steps=None):
    """CPU oracle on the host cores: predict mean+var points/s and stretch-move
    walker-steps/s on a bounded sample of the c2 workload."""
    from oracle import gp as ogp, emcee as oem
    try:
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        cores = 1
    X, y, hp, bounds = workload()
    gp = ogp.OracleGP("Matern32Kernel", NDIM, hp["log_M"], log_const=np.log(hp["amp"]), mean=hp["mean"],
                      white_noise=hp["white_noise"])
    gp.compute(X)
    rng = np.random.default_rng(7)
    m = int(sample_points or 100000)
    chunk = 20000
    times = []
    for it in range(warmup + steps):
        xq = rng.uniform(-6, 6, size=(m, NDIM))
        t0 = time.perf_counter()
        for a in range(0, m, chunk):
            gp.predict(y, xq[a:a + chunk], return_var=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    pred = m * len(times) / sum(times)
    b = np.asarray(bounds)

    def lp(q):
        q = np.atleast_2d(q)
        inside = np.all((q > b[:, 0]) & (q < b[:, 1]), axis=1)
        return np.where(inside, gp.predict(y, q), -np.inf)
    ns = int(mcmc_steps or 20)
    s = oem.StretchEnsemble(NWALKERS, NDIM, lp, seed=3, vectorize=True)
    p0 = rng.uniform(-6, 6, size=(NWALKERS, NDIM))
    s.run_mcmc(p0, 2)
    t0 = time.perf_counter()
    s.run_mcmc(s.chain[-1], ns)
    mc = NWALKERS * ns / (time.perf_counter() - t0)
    return dict(pred=pred, mcmc=mc, cores=int(cores), ms_per_step=1e3 * sum(times) / len(times),
                sample=f"{m} query points per step (chunks of {chunk}), {ns} ensemble steps of {NWALKERS} walkers "
                       f"with vectorised log-prob; NumPy/SciPy oracle, {cores} BLAS threads")


def kernel_table(lib, dmma_peak_tflops):
    """K1 / K2 at the training-set sizes of BASELINE configs c4 and c5 (rank 0, N = 1 only;
    a few hundred ms of GPU time): covariance build, Cholesky, both triangular solves and
    the log-likelihood gradient, each as absolute time and as a fraction of its roof
    (HBM copy bandwidth from MEASURED_PEAKS.json, FP64 DMMA peak measured in this run)."""
    import torch
    import alabi_b200 as ab
    from alabi_b200 import _lib
    hbm = 6458.1
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass

    def ev(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best
    out = {"hbm_peak_gb_s": hbm, "fp64_tensor_peak_tflops": dmma_peak_tflops}
    for name, n, d in (("c4", 8192, 10), ("c5", 16384, 20)):
        rng = np.random.default_rng(n)
        X = rng.uniform(-3, 3, size=(n, d))
        y = -0.5 * np.sum((X / 1.5) ** 2, axis=1) + 0.01 * rng.normal(size=n)
        g = ab.GP(kernel=ab.kernels.ExpSquaredKernel(metric=np.full(d, 2.25), ndim=d) * np.var(y), fit_mean=True,
                  mean=np.median(y), white_noise=-6.0, fit_white_noise=True)
        g.compute(X)
        h = g._hd.h
        lib.ab_gp_set_profiling(h, 1)
        t_fac = ev(lambda: lib.ab_gp_factor(h))                      # covariance build + Cholesky
        cms, ccnt = ctypes.c_double(), ctypes.c_longlong()
        lib.ab_gp_profile_read(h, 0, ctypes.byref(cms), ctypes.byref(ccnt))
        fms, fcnt = ctypes.c_double(), ctypes.c_longlong()
        lib.ab_gp_profile_read(h, 1, ctypes.byref(fms), ctypes.byref(fcnt))
        lib.ab_gp_set_profiling(h, 0)
        t_cov = cms.value * 1e-3 / max(ccnt.value, 1)
        t_chol = fms.value * 1e-3 / max(fcnt.value, 1)
        yd = torch.from_numpy(y).cuda()
        ll = ctypes.c_double()
        t_solve = ev(lambda: lib.ab_gp_log_likelihood(h, _lib.ptr(yd), ctypes.byref(ll)))
        gr = (ctypes.c_double * (d + 3))()

        def full_grad():
            lib.ab_gp_factor(h)
            lib.ab_gp_grad_log_likelihood(h, _lib.ptr(yd), gr)
        t_grad = ev(full_grad, reps=2)
        out[name] = {
            "n_train": n, "ndim": d,
            "cov_build": {"ms": t_cov * 1e3, "gb_s": 4.0 * n * n / t_cov * 1e-9, "frac_hbm": 4.0 * n * n / t_cov * 1e-9 / hbm,
                          "bytes": "4 N^2 (lower triangle written once)", "note": "FP64-ALU bound for d >= 4 (DESIGN.md)"},
            "cholesky": {"ms": t_chol * 1e3, "tflops": n ** 3 / 3.0 / t_chol * 1e-12,
                         "frac_fp64_tensor": n ** 3 / 3.0 / t_chol * 1e-12 / dmma_peak_tflops},
            "compute_total_ms": t_fac * 1e3,
            "solves_loglike_ms": t_solve * 1e3,
            "factor_plus_gradient_ms": t_grad * 1e3,
            "factor_plus_gradient_tflops": float(n) ** 3 / t_grad * 1e-12,
        }
        del g
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    line = {"metric": METRIC, "value": r["pred"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": r["pred"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["pred"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mcmc": {"value": r["mcmc"], "unit": "walker-steps/s"}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true",
                    help="skip the K1 / K2 measurements at the c4 / c5 training-set sizes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner)
    # are sent to stderr until the line is written
    sys.stdout.flush()
    _saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import alabi_b200 as ab
    from alabi_b200 import _lib, parallel as par
    from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

    rank, world, local = par.init_distributed()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    warmup = max(args.warmup, 3)

    # ---- train once on rank 0, broadcast L and alpha ------------------------------------
    X, y, hp, bounds = workload()
    kern = ab.kernels.Matern32Kernel(metric=np.exp(hp["log_M"]), ndim=NDIM) * (hp["amp"] * NDIM)
    gp = ab.GP(kernel=kern, fit_mean=True, mean=hp["mean"], white_noise=hp["white_noise"], fit_white_noise=True)
    if rank == 0:
        gp.compute(X)
        gp._set_targets(y)
    if world > 1:
        par.broadcast_gp(gp)
        dist.barrier()
    h = gp._hd.h

    # ---- inputs: resident in HBM, rotating so that no step finds its batch in L2 -----------
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    xq = [(torch.rand((M_PER_STEP, NDIM), generator=g, device=dev, dtype=torch.float64) * 12.0 - 6.0)
          for _ in range(N_ROTATE)]
    mu = [torch.empty(M_PER_STEP, dtype=torch.float64, device=dev) for _ in range(N_ROTATE)]
    var = [torch.empty(M_PER_STEP, dtype=torch.float64, device=dev) for _ in range(N_ROTATE)]

    def step(i):
        k = i % N_ROTATE
        _lib.check(lib.ab_gp_predict(h, _lib.ptr(xq[k]), M_PER_STEP, _lib.ptr(mu[k]), _lib.ptr(var[k])), "ab_gp_predict")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    lib.ab_gp_set_profiling(h, 1)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.ab_launch_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(warmup + i)
    e1.record()
    barrier()
    launches = lib.ab_launch_counter() - launches0
    sampler.stop_flag = True
    sampler.join(2)
    el = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el)
    value = world * M_PER_STEP * args.steps / elapsed
    kms, kcnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h, 3, ctypes.byref(kms), ctypes.byref(kcnt))       # variance GEMM launches
    pms, pcnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h, 2, ctypes.byref(pms), ctypes.byref(pcnt))
    lib.ab_gp_set_profiling(h, 0)

    # ---- e2e: public API with host buffers (pinned), copies inside the timed region -----------
    m_e2e = M_PER_STEP
    host_q = [torch.empty((m_e2e, NDIM), dtype=torch.float64).pin_memory() for _ in range(2)]
    for t_ in host_q:
        t_.copy_(torch.rand((m_e2e, NDIM), dtype=torch.float64) * 12.0 - 6.0)
    host_q_np = [t_.numpy() for t_ in host_q]
    for i in range(2):
        gp.predict(y, host_q_np[i % 2], return_var=True)
    barrier()
    e_steps = max(args.steps // 2, 2)
    t0 = time.perf_counter()
    for i in range(e_steps):
        m_, v_ = gp.predict(y, host_q_np[i % 2], return_var=True)
    barrier()
    e2e_el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_el, op=dist.ReduceOp.MAX)
    e2e_val = world * m_e2e * e_steps / float(e2e_el)

    # ---- surrogate MCMC (K5): independent sub-ensembles, one per GPU ---------------------------------
    lp = SurrogateLogProb(gp, y, bounds)
    es = EnsembleSampler(NWALKERS, NDIM, lp, seed=99)
    p0 = np.random.default_rng(5 + rank).uniform(-6, 6, size=(NWALKERS, NDIM))
    es.run_mcmc(p0, 200, store=False, walker_offset=rank * NWALKERS)
    barrier()
    es.run_mcmc(None, MCMC_STEPS, store=False, walker_offset=rank * NWALKERS)
    mc_dev = torch.tensor([es.last_run_device_seconds], dtype=torch.float64, device=dev)
    barrier()
    t0 = time.perf_counter()
    es2 = EnsembleSampler(NWALKERS, NDIM, lp, seed=100)
    es2.run_mcmc(p0, MCMC_STEPS, walker_offset=rank * NWALKERS)          # chain + log-prob copied to the host
    barrier()
    mc_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mc_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(mc_e2e, op=dist.ReduceOp.MAX)
    mcmc = {"value": world * NWALKERS * MCMC_STEPS / float(mc_dev), "unit": "walker-steps/s",
            "e2e": world * NWALKERS * MCMC_STEPS / float(mc_e2e), "walkers_per_gpu": NWALKERS, "steps": MCMC_STEPS,
            "acceptance": float(es.acceptance_fraction.mean()),
            "d2h_bytes": int(MCMC_STEPS * NWALKERS * (NDIM + 1) * 8)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (variance GEMM): FP64 tensor pipe -------------------------------
    peak = ctypes.c_double()
    lib.ab_fp64_tensor_peak(local, ctypes.byref(peak))
    flops_per_point = float(N_TRAIN) ** 2                   # N^2 per query point (triangular solve)
    q_per_launch = M_PER_STEP * args.steps / max(kcnt.value, 1)
    avg_launch_s = kms.value * 1e-3 / max(kcnt.value, 1)
    achieved = q_per_launch * flops_per_point / avg_launch_s * 1e-12
    traffic = None
    tf = os.path.join(ROOT, "profiles", "predict_var_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "tensor", "kernel": "predict_var_kernel (DMMA.8x8x4 GEMM L^-1 x K*)", "achieved": achieved,
                "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": traffic,
                "peak_source": "FP64 DMMA issue-rate microbench measured in this run (MEASURED_PEAKS.json has no FP64 "
                               "entry; nominal 40 TFLOP/s)",
                "launches": int(kcnt.value), "avg_launch_ms": avg_launch_s * 1e3,
                "share_of_step": kms.value * 1e-3 / elapsed,
                "panel_kernel_share_of_step": pms.value * 1e-3 / elapsed}

    kernels = None if (args.no_kernel_table or world > 1) else kernel_table(lib, peak.value)
    if kernels is not None:
        # the reference's M = 1 calls (acquisition polish, host samplers) on the c2 model, end to end
        # through GP.predict / GP.predict_grad with host buffers (few-query kernels)
        x1 = np.array([[0.7, -1.3]])

        def per_call_us(f, n=300):
            for _ in range(20):
                f()
            torch.cuda.synchronize()
            t0_ = time.perf_counter()
            for _ in range(n):
                f()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0_) / n * 1e6
        kernels["one_point_calls_c2"] = {
            "predict_mean_var_us": per_call_us(lambda: gp.predict(y, x1, return_var=True)),
            "predict_mean_us": per_call_us(lambda: gp.predict(y, x1, return_cov=False)),
            "predict_grad_us": per_call_us(lambda: gp.predict_grad(y, x1))}

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(steps=2, warmup=1, sample_points=100000, mcmc_steps=10)
        cpu = {"value": r["pred"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "mcmc_walker_steps_per_s": r["mcmc"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(m_e2e * NDIM * 8),
                    "d2h_bytes_per_step": int(m_e2e * 16)},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "mcmc": mcmc, "kernels": kernels}
    sys.stdout.flush()
    os.dup2(_saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
