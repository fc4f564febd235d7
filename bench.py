#!/usr/bin/env python
"""Headline benchmark of the alabi surrogate hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = the two largest configurations of BASELINE.json (inputs from ``alabi_b200.workloads``):

* **c4** (configs[3]: 10-D correlated Gaussian, ExpSquared, N = 8192, "batched predict over 1e7
  query points and 1e6-candidate utility argmax sharded over 8 GPUs") carries the headline
  ``value``: GP predictive mean + variance, points/s.  One *step* is one batch of ``M_TOTAL``
  = 606 208 candidate points (a slice of the config's query / candidate sets) for the WHOLE job:
  every rank predicts mean + variance of its contiguous share (K3), evaluates the BAPE utility
  and its argmin (K4), and one ``all_gather`` of a (value, global index) pair per rank picks the
  global winner — the collective is inside the timed step.  Total work per step is fixed, so
  ``scaling`` is "strong".  The GP is trained once on rank 0 (K1 + K2) and L, the diagonal-block
  inverses and alpha are broadcast over NCCL; the broadcast is timed and reported under ``train``.
* **c5** (configs[4]: 20-D, N = 16384, "emcee 65536 walkers sharded across 8 B200") carries
  ``mcmc``: surrogate-MCMC walker-steps/s (K5), 65 536 walkers in total split into one
  sub-ensemble per rank, ``MCMC_STEPS`` stored steps, and the ``all_gather`` of the chain blocks
  over NVLink inside the timed region — fused into the sampler kernel: every rank's kernel stores
  its rows into the gathered buffer of every rank (peer memory, ``parallel.PeerChainBuffers``), and
  the timed region ends with the barrier after which every rank holds the whole chain
  (``mcmc.value_with_nccl_allgather_after_the_run`` is the unfused route, timed beside it).  At
  N = 1 the single GPU advances all 65 536 walkers.

``value`` is measured with the inputs resident in HBM (CUDA events, max over ranks).  ``e2e`` is the
same metric through the public API on HOST buffers (``GP.predict`` on pageable NumPy arrays:
host-to-device copy of the queries and device-to-host copy of mean and variance inside the timed
region); ``mcmc.e2e`` goes through ``parallel.sharded_ensemble`` / ``EnsembleSampler.run_mcmc`` with
the chain delivered to host memory.  ``roofline`` describes the dominant kernel of the headline
(the variance GEMM, FP64 tensor pipe) and ``mcmc.roofline`` the sampler kernel (FP64 FMA pipe),
both against peaks measured in this run and against the nominal 40 TFLOP/s.  ``extra.c2`` keeps
round 1's c2 numbers (N = 1000, 2-D shells), ``kernels`` the covariance build / Cholesky / solves /
gradient at the c4 and c5 sizes (N = 1 only), ``cpu_baseline`` the NumPy/SciPy oracle on the host
cores for a bounded sample of the same workload (N = 1 only).

``--impl reference`` times that CPU oracle (the restatement of the george + emcee path; george and
emcee themselves are not installable here) on the same configurations with all
host BLAS threads, each step a bounded sample, and prints the same JSON line with
``"impl": "reference"``.
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
M_TOTAL = 8 * 148 * 128 * 4     # candidate points per step over all GPUs (4 variance waves per GPU at 8 GPUs)
N_ROTATE = 3                    # distinct query batches per rank
WALKERS_TOTAL = 65536           # c5
MCMC_STEPS = 64
CPU_SAMPLE_POINTS = 2048        # CPU arm: query points per step (bounded sample of the c4 batch)
CPU_SAMPLE_WALKERS, CPU_SAMPLE_MCMC_STEPS = 256, 2
NOMINAL_FP64_TFLOPS = 40.0      # north_star's figure for both the FP64 tensor and the FP64 FMA pipe


def config_dict(n_gpus):
    return {"workload": "c4 (BASELINE configs[3]): 10-D correlated Gaussian, ExpSquaredKernel, N=8192, white_noise=-12; "
                        f"one step = predict mean+var + BAPE utility argmin over {M_TOTAL} candidate points in total "
                        "(a slice of the config's 1e7 queries / 1e6 candidates), sharded over the GPUs, one all_gather "
                        "of (value, index) per step inside the timed region.  mcmc: c5 (configs[4]): 20-D, N=16384, "
                        f"{WALKERS_TOTAL} walkers in total x {MCMC_STEPS} stored steps, one sub-ensemble per GPU, "
                        "chain-block all_gather inside the timed region",
            "n_train": 8192, "ndim": 10, "kernel": "ExpSquaredKernel", "points_per_step_total": M_TOTAL,
            "mcmc_n_train": 16384, "mcmc_ndim": 20, "mcmc_walkers_total": WALKERS_TOTAL, "mcmc_steps": MCMC_STEPS,
            "l2_policy": f"inputs rotate over {N_ROTATE} distinct query batches; every 18944-query panel writes and "
                         "re-reads a 1.24 GB cross-covariance panel (> 126 MB L2) between two uses of any input",
            "parallelism": f"factor trained on rank 0 and broadcast (NCCL); queries / candidates / walkers sharded x{n_gpus}"}


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


def sampler_flops_per_walker_step(n, d):
    """K3 mean-only + 6 d = N (3 d + c_k + 2) + 6 d with c_k = 25 (exp)."""
    return float(n) * (3 * d + 25 + 2) + 6 * d


# ------------------------------------------------------------------------------------------------
# CPU arm: the NumPy/SciPy oracle on the host cores (bounded samples of the same workload)
# ------------------------------------------------------------------------------------------------
def _blas_threads():
    try:
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
        return int(max([p.get("num_threads", 1) for p in threadpool_info()] + [1]))
    except Exception:  # noqa: BLE001
        return 1


def cpu_reference_run(steps, warmup, sample_points=CPU_SAMPLE_POINTS, mcmc=True, alpha5=None):
    """Oracle GP on c4: ``steps`` timed passes of predict mean+var over ``sample_points`` query points;
    oracle stretch move on the c5 surrogate: a few ensemble steps of a bounded sub-ensemble.
    ``alpha5``: K^-1 (y - mean) of the c5 model when the caller already holds it (the b200 arm's
    in-run baseline passes the device's, which spares a minute of CPU Cholesky that is training,
    not sampling; the ``--impl reference`` arm factorises on the CPU)."""
    from oracle import gp as ogp, emcee as oem
    from alabi_b200 import workloads
    cores = _blas_threads()
    c4 = workloads.make_config("c4")
    hp = c4["hp"]
    t0 = time.perf_counter()
    gp = ogp.make_gp(c4["kind"], c4["X"], c4["y"], hp["log_M"], amp=hp["amp"], mean=hp["mean"],
                     white_noise=hp["white_noise"])
    t_factor = time.perf_counter() - t0
    rng = np.random.default_rng(7)
    m = int(sample_points)
    times = []
    for it in range(warmup + steps):
        xq = rng.uniform(-3, 3, size=(m, c4["ndim"]))
        t0 = time.perf_counter()
        gp.predict(c4["y"], xq, return_var=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    pred = m * len(times) / sum(times)
    out = dict(pred=pred, cores=cores, ms_per_step=1e3 * sum(times) / len(times), factor_s=t_factor,
               sample=f"c4 model (N=8192, CPU Cholesky {t_factor:.1f} s, not counted); {m} query points per step, "
                      f"{len(times)} timed steps; NumPy/SciPy oracle (LAPACK cho_solve), {cores} BLAS threads")
    if mcmc:
        c5 = workloads.make_config("c5")
        hp5 = c5["hp"]
        b = c5["bounds"]
        if alpha5 is None:
            g5 = ogp.make_gp(c5["kind"], c5["X"], c5["y"], hp5["log_M"], amp=hp5["amp"], mean=hp5["mean"],
                             white_noise=hp5["white_noise"])
            alpha5 = g5._compute_alpha(c5["y"])
            how = "alpha from the CPU Cholesky"
        else:
            how = "alpha taken from the device factorisation (training is not part of a sampler step)"
        log_const = float(np.log(hp5["amp"] / c5["ndim"]))

        def lp(q):        # oracle predict, mean only: k(x*, X) alpha + mean  (oracle/gp.py OracleGP.predict)
            q = np.atleast_2d(q)
            inside = np.all((q > b[:, 0]) & (q < b[:, 1]), axis=1)
            mu = ogp.kernel_value(c5["kind"], q, c5["X"], hp5["log_M"], log_const) @ alpha5 + hp5["mean"]
            return np.where(inside, mu, -np.inf)
        nw, ns = CPU_SAMPLE_WALKERS, CPU_SAMPLE_MCMC_STEPS
        s = oem.StretchEnsemble(nw, c5["ndim"], lp, seed=3, vectorize=True)
        p0 = rng.uniform(-1, 1, size=(nw, c5["ndim"]))
        s.run_mcmc(p0, 1)
        t0 = time.perf_counter()
        s.run_mcmc(s.chain[-1], ns)
        out["mcmc"] = nw * ns / (time.perf_counter() - t0)
        out["sample"] += (f"; mcmc: c5 surrogate (N=16384, d=20), {ns} ensemble steps of {nw} walkers, vectorised "
                          f"log-prob, {how}")
    return out


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    line = {"metric": METRIC, "value": r["pred"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": r["pred"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["pred"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mcmc": {"value": r.get("mcmc"), "unit": "walker-steps/s"}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# device helpers
# ------------------------------------------------------------------------------------------------
def kernel_table(lib, dmma_peak_tflops):
    """K1 / K2 at the training-set sizes of BASELINE configs c4 and c5 (rank 0, N = 1 only;
    a few hundred ms of GPU time): covariance build, Cholesky, both triangular solves and
    the log-likelihood gradient, each as absolute time and as a fraction of its roof
    (HBM copy bandwidth from MEASURED_PEAKS.json, FP64 DMMA peak measured in this run)."""
    import torch
    from alabi_b200 import _lib, workloads
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
    for name in ("c4", "c5"):
        cfg = workloads.make_config(name)
        n, d, y = len(cfg["X"]), cfg["ndim"], cfg["y"]
        g = workloads.build_gp(cfg)
        g.compute(cfg["X"])
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
                         "frac_fp64_tensor": n ** 3 / 3.0 / t_chol * 1e-12 / dmma_peak_tflops,
                         "frac_nominal_40": n ** 3 / 3.0 / t_chol * 1e-12 / NOMINAL_FP64_TFLOPS},
            "compute_total_ms": t_fac * 1e3,
            "solves_loglike_ms": t_solve * 1e3,
            "factor_plus_gradient_ms": t_grad * 1e3,
            "factor_plus_gradient_tflops": float(n) ** 3 / t_grad * 1e-12,
        }
        del g
        torch.cuda.empty_cache()
    return out


def c2_extra(lib, steps=5, fma_peak_tflops=None):
    """Round 1's headline configuration (c2: 2-D shells, Matern-3/2, N = 1000), one GPU: predict
    mean+var points/s (device-resident and through GP.predict on host buffers), the sampler
    (1000 walkers x 5000 steps) device-timed and end to end, and one-point call latencies."""
    import torch
    from alabi_b200 import _lib, workloads
    from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb
    cfg = workloads.make_config("c2")
    gp = workloads.build_gp(cfg)
    gp.compute(cfg["X"])
    y = cfg["y"]
    gp._set_targets(y)
    h = gp._hd.h
    m = 1 << 20
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(77)
    xq = [torch.rand((m, 2), generator=g, device=dev, dtype=torch.float64) * 12.0 - 6.0 for _ in range(4)]
    mu = torch.empty(m, dtype=torch.float64, device=dev)
    var = torch.empty(m, dtype=torch.float64, device=dev)
    for i in range(3):
        lib.ab_gp_predict(h, _lib.ptr(xq[i % 4]), m, _lib.ptr(mu), _lib.ptr(var))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        lib.ab_gp_predict(h, _lib.ptr(xq[i % 4]), m, _lib.ptr(mu), _lib.ptr(var))
    e1.record()
    e1.synchronize()
    pred = m * steps / (e0.elapsed_time(e1) * 1e-3)
    hq = [np.random.default_rng(i).uniform(-6, 6, size=(m, 2)) for i in range(2)]
    gp.predict(y, hq[0], return_var=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(max(steps // 2, 2)):
        gp.predict(y, hq[i % 2], return_var=True)
    torch.cuda.synchronize()
    pred_e2e = m * max(steps // 2, 2) / (time.perf_counter() - t0)
    nw, ns = 1000, 5000
    lp = SurrogateLogProb(gp, y, cfg["bounds"])
    p0 = np.random.default_rng(5).uniform(-6, 6, size=(nw, 2))
    es = EnsembleSampler(nw, 2, lp, seed=99)
    es.run_mcmc(p0, 200, store=False)
    es.run_mcmc(None, ns, store=False)
    mc_dev = nw * ns / es.last_run_device_seconds
    EnsembleSampler(nw, 2, lp, seed=98).run_mcmc(p0, ns)          # warms the page-locked chain buffers
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    es2 = EnsembleSampler(nw, 2, lp, seed=100)
    es2.run_mcmc(p0, ns)
    mc_e2e = nw * ns / (time.perf_counter() - t0)
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
    one = {"predict_mean_var_us": per_call_us(lambda: gp.predict(y, x1, return_var=True)),
           "predict_mean_us": per_call_us(lambda: gp.predict(y, x1, return_cov=False)),
           "predict_grad_us": per_call_us(lambda: gp.predict_grad(y, x1))}
    return {"workload": "c2: 2-D Gaussian shells, Matern-3/2, N=1000, white_noise=-12; one GPU",
            "predict_points_per_s": pred, "predict_e2e_points_per_s": pred_e2e,
            "mcmc_walker_steps_per_s": mc_dev, "mcmc_e2e_walker_steps_per_s": mc_e2e,
            "mcmc_shape": f"{nw} walkers x {ns} steps, chain and log-prob delivered to host ({ns * nw * 3 * 8} B)",
            "sampler_flops_per_walker_step": 1000.0 * (3 * 2 + 35 + 2) + 12,
            "mcmc_roofline": {"bound": "latency (dataflow schedule: per half-step one publish -> poll round trip through the L2, "
                                       "the kernel evaluations of two units per SM and the accept; DESIGN.md section 3); "
                                       "against the FP64 FMA roof for reference",
                              "achieved": mc_dev * (1000.0 * (3 * 2 + 35 + 2) + 12) * 1e-12, "peak": fma_peak_tflops,
                              "unit": "TFLOP/s",
                              "frac": (mc_dev * (1000.0 * (3 * 2 + 35 + 2) + 12) * 1e-12 / fma_peak_tflops) if fma_peak_tflops else None},
            "one_point_calls": one}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true",
                    help="skip the K1 / K2 measurements at the c4 / c5 training-set sizes and the c2 extras")
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
    from alabi_b200 import _lib, parallel as par, workloads
    from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

    rank, world, local = par.init_distributed()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def sum_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t)

    def train_and_broadcast(name):
        """rank 0: K1 + K2 + alpha; then the NCCL broadcast of L, D^-1 and alpha (timed)."""
        cfg = workloads.make_config(name)
        gp = workloads.build_gp(cfg)
        info = {"n_train": len(cfg["X"]), "ndim": cfg["ndim"]}
        if rank == 0:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gp.compute(cfg["X"])
            gp._set_targets(cfg["y"])
            torch.cuda.synchronize()
            info["train_ms"] = (time.perf_counter() - t0) * 1e3
        if world > 1:
            barrier()                      # sets the NCCL communicator up
            warm = torch.zeros(1 << 24, dtype=torch.float64, device=dev)
            dist.broadcast(warm, src=0)    # and the large-message broadcast channels (128 MB), before the timed broadcast
            del warm
            barrier()
            st = {}
            par.broadcast_gp(gp, stats=st)
            barrier()
            sec = max_over_ranks(st.get("seconds", 0.0))
            info["broadcast"] = {"bytes": st.get("bytes"), "ms": sec * 1e3,
                                 "gb_s": (st.get("bytes", 0) / sec * 1e-9) if sec > 0 else None,
                                 "nvlink_peak_gb_s_per_direction": 900.0,
                                 "what": "L (npad x npad) + diagonal-block inverses + alpha, rank 0 -> all"}
        return cfg, gp, info

    # =============================================================================================
    # c4: predict mean + var + utility argmin, candidates sharded, all_gather of the argmin
    # =============================================================================================
    cfg4, gp4, train4 = train_and_broadcast("c4")
    y4, d4, n4 = cfg4["y"], cfg4["ndim"], len(cfg4["X"])
    h4 = gp4._hd.h
    lo, hi = par.shard_range(M_TOTAL, rank, world)
    m_loc = hi - lo
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    # candidates slightly beyond the box, so the strict-bounds test of the utilities has work to do
    xq = [(torch.rand((m_loc, d4), generator=g, device=dev, dtype=torch.float64) * 6.2 - 3.1) for _ in range(N_ROTATE)]
    mu = [torch.empty(m_loc, dtype=torch.float64, device=dev) for _ in range(N_ROTATE)]
    var = [torch.empty(m_loc, dtype=torch.float64, device=dev) for _ in range(N_ROTATE)]
    bnds = np.ascontiguousarray(cfg4["bounds"].reshape(-1))
    bptr = bnds.ctypes.data_as(_lib.c_double_p)
    y_best = float(np.max(y4))
    winners = []

    def step(i):
        k = i % N_ROTATE
        _lib.check(lib.ab_gp_predict(h4, _lib.ptr(xq[k]), m_loc, _lib.ptr(mu[k]), _lib.ptr(var[k])), "ab_gp_predict")
        idx, val = ctypes.c_int64(), ctypes.c_double()
        _lib.check(lib.ab_utility_eval(h4, 0, _lib.ptr(xq[k]), _lib.ptr(mu[k]), _lib.ptr(var[k]), m_loc, bptr, y_best, 0.01,
                                       None, ctypes.byref(idx), ctypes.byref(val)), "ab_utility_eval")
        winners.append(par.argmin_allgather(val.value, idx.value + lo if idx.value >= 0 else -1))

    for i in range(warmup):
        step(i)
    barrier()
    lib.ab_gp_set_profiling(h4, 1)
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
    launches = sum_over_ranks(lib.ab_launch_counter() - launches0)
    sampler.stop_flag = True
    sampler.join(2)
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    value = M_TOTAL * args.steps / elapsed
    kms, kcnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h4, 3, ctypes.byref(kms), ctypes.byref(kcnt))       # variance GEMM launches
    pms, pcnt = ctypes.c_double(), ctypes.c_longlong()
    lib.ab_gp_profile_read(h4, 2, ctypes.byref(pms), ctypes.byref(pcnt))       # cross-covariance panel + mean
    lib.ab_gp_set_profiling(h4, 0)
    own_elapsed = e0.elapsed_time(e1) * 1e-3

    # ---- e2e: GP.predict on pageable host arrays (H2D of the queries, D2H of mean and variance) -------
    hrng = np.random.default_rng(4321 + rank)
    host_q = [hrng.uniform(-3.1, 3.1, size=(m_loc, d4)) for _ in range(2)]
    gp4.predict(y4, host_q[0], return_var=True)
    barrier()
    e_steps = max(args.steps // 2, 2)
    t0 = time.perf_counter()
    for i in range(e_steps):
        m_, v_ = gp4.predict(y4, host_q[i % 2], return_var=True)
    barrier()
    e2e_val = M_TOTAL * e_steps / max_over_ranks(time.perf_counter() - t0)
    del host_q, xq, mu, var
    gp4 = None                                  # c4 factor + inverse (1.6 GB) released before c5
    torch.cuda.empty_cache()

    # =============================================================================================
    # c5: surrogate MCMC, 65536 walkers in total, chain-block all_gather inside the timed region
    # =============================================================================================
    cfg5, gp5, train5 = train_and_broadcast("c5")
    y5, d5, n5 = cfg5["y"], cfg5["ndim"], len(cfg5["X"])
    lp5 = SurrogateLogProb(gp5, y5, cfg5["bounds"])
    wlo, whi = par.shard_range(WALKERS_TOTAL, rank, world)
    p0 = np.random.default_rng(55).uniform(-1.0, 1.0, size=(WALKERS_TOTAL, d5))
    es = EnsembleSampler(whi - wlo, d5, lp5, seed=99)
    # warm-up: the same run and gather once (kernel module, NCCL channels of these sizes, the chain
    # blocks in torch's caching allocator), results discarded
    es.run_mcmc(p0[wlo:whi], MCMC_STEPS, store="device", walker_offset=wlo)
    _w1 = par.allgather_walkers(es.device_chain, WALKERS_TOTAL)
    _w2 = par.allgather_walkers(es.device_log_prob, WALKERS_TOTAL)
    del _w1, _w2
    es.device_chain = es.device_log_prob = None
    es._device_rows_pending = False
    # N > 1: the chain-block all_gather is FUSED into the sampler kernel (every rank's kernel stores its
    # rows into the gathered buffer of every rank over NVLink, parallel.PeerChainBuffers); what is left of
    # the collective inside the timed region is the barrier that tells a rank all columns are in place
    fused = world > 1 and WALKERS_TOTAL % world == 0
    bufs = par.PeerChainBuffers.get(lib, local, MCMC_STEPS, WALKERS_TOTAL, d5) if fused else None
    fused = bufs is not None                                   # no peer access: the NCCL route
    if fused:
        es.run_mcmc(None, MCMC_STEPS, store="device", walker_offset=wlo, gather=bufs.spec(wlo))     # warm-up of this route
        torch.cuda.synchronize()
        dist.barrier()
        es.device_chain = es.device_log_prob = None
        es._device_rows_pending = False          # (pending device rows would be copied to the host by the next run)
    barrier()
    # the unfused route (one NCCL all_gather_into_tensor per array after the run), timed for comparison
    mc_nccl_s = None
    if world > 1:
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        es.run_mcmc(None, MCMC_STEPS, store="device", walker_offset=wlo)
        _w1 = par.allgather_walkers(es.device_chain, WALKERS_TOTAL)
        _w2 = par.allgather_walkers(es.device_log_prob, WALKERS_TOTAL)
        u1.record()
        barrier()
        mc_nccl_s = max_over_ranks(u0.elapsed_time(u1) * 1e-3)
        del _w1, _w2
        es.device_chain = es.device_log_prob = None
        es._device_rows_pending = False
        barrier()
    mc_launch0 = lib.ab_launch_counter()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    if fused:
        es.run_mcmc(None, MCMC_STEPS, store="device", walker_offset=wlo, gather=bufs.spec(wlo))
        torch.cuda.synchronize()
        dist.barrier()
        full_chain, full_logp = bufs.chain, bufs.log_prob
    else:
        es.run_mcmc(None, MCMC_STEPS, store="device", walker_offset=wlo)
        full_chain = par.allgather_walkers(es.device_chain, WALKERS_TOTAL)
        full_logp = par.allgather_walkers(es.device_log_prob, WALKERS_TOTAL)
    m1.record()
    barrier()
    mc_dev_s = max_over_ranks(m0.elapsed_time(m1) * 1e-3)
    mc_kernel_s = max_over_ranks(es.last_run_device_seconds)
    mc_launches = sum_over_ranks(lib.ab_launch_counter() - mc_launch0)
    gathered_ok = tuple(full_chain.shape) == (MCMC_STEPS, WALKERS_TOTAL, d5)
    if fused:
        # every rank's columns of the last stored row arrived in THIS rank's buffer: they equal what the owner holds
        last_local = es.device_chain[-1].contiguous()
        rows = [torch.empty_like(last_local) for _ in range(world)]
        dist.all_gather(rows, last_local)
        gathered_ok = gathered_ok and bool(torch.equal(torch.cat(rows, dim=0), full_chain[-1]))
    acc = float(es.acceptance_fraction.mean())
    alpha5_host = np.array(gp5._alpha) if (rank == 0 and world == 1) else None
    del full_chain, full_logp
    es.device_chain = es.device_log_prob = None
    torch.cuda.empty_cache()
    # end to end through the public API: chain in host memory (rank 0 holds the gathered chain)
    factory = lambda k: EnsembleSampler(k, d5, lp5, seed=100)
    par.sharded_ensemble(factory, p0, MCMC_STEPS, gather=True, to_host=(rank == 0))   # same sizes: warms the page-locked buffers
    barrier()
    t0 = time.perf_counter()
    s_e2e, chain_host = par.sharded_ensemble(factory, p0, MCMC_STEPS, gather=True, to_host=(rank == 0))
    barrier()
    mc_e2e_s = max_over_ranks(time.perf_counter() - t0)
    chain_bytes = int(MCMC_STEPS * WALKERS_TOTAL * d5 * 8)
    del chain_host, s_e2e

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rooflines ---------------------------------------------------------------------------------------
    peak = ctypes.c_double()
    lib.ab_fp64_tensor_peak(local, ctypes.byref(peak))
    fma_peak = ctypes.c_double()
    lib.ab_fp64_fma_peak(local, ctypes.byref(fma_peak))
    flops_per_point = float(n4) ** 2                        # N^2 per query point (triangular solve)
    q_per_launch = m_loc * args.steps / max(kcnt.value, 1)
    avg_launch_s = kms.value * 1e-3 / max(kcnt.value, 1)
    achieved = q_per_launch * flops_per_point / avg_launch_s * 1e-12
    traffic, traffic_src = None, None
    tf = os.path.join(ROOT, "profiles", "predict_var_traffic.json")
    if os.path.exists(tf):
        try:
            tj = json.load(open(tf))
            traffic = tj.get("c4", {}).get("dram_bytes_per_launch")
            traffic_src = "static: " + tj.get("c4", {}).get("source", "profiles/predict_var_traffic.json") \
                if traffic is not None else None
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "tensor", "kernel": "predict_var_pair_kernel (DMMA.8x8x4 GEMM L^-1 x K*, row-block pairs of a query tile on adjacent CTAs), c4: N=8192",
                "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value,
                "frac_of_nominal_40": achieved / NOMINAL_FP64_TFLOPS, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": q_per_launch * n4 * 8.0 + 4.0 * n4 * n4,
                "peak_source": "FP64 DMMA issue-rate microbench measured in this run (MEASURED_PEAKS.json has no FP64 "
                               "entry; nominal 40 TFLOP/s)",
                "launches": int(kcnt.value), "avg_launch_ms": avg_launch_s * 1e3,
                "queries_per_launch": q_per_launch,
                "share_of_step": kms.value * 1e-3 / own_elapsed,
                "panel_kernel_share_of_step": pms.value * 1e-3 / own_elapsed}
    fl_ws = sampler_flops_per_walker_step(n5, d5)
    mc_ach = fl_ws * WALKERS_TOTAL * MCMC_STEPS / world / mc_kernel_s * 1e-12      # per GPU
    mcmc = {"value": WALKERS_TOTAL * MCMC_STEPS / mc_dev_s, "unit": "walker-steps/s",
            "e2e": WALKERS_TOTAL * MCMC_STEPS / mc_e2e_s, "walkers_total": WALKERS_TOTAL, "steps": MCMC_STEPS,
            "kernel_only": WALKERS_TOTAL * MCMC_STEPS / mc_kernel_s, "acceptance": acc,
            "gathered_chain_ok": bool(gathered_ok), "gpu_launches": int(mc_launches),
            "allgather_bytes_per_rank": int(chain_bytes / world + MCMC_STEPS * WALKERS_TOTAL * 8 / world),
            "allgather": ("fused into the sampler kernel: every rank's kernel stores its chain rows into the gathered buffer of "
                          "every rank (peer memory over NVLink, CUDA IPC); the timed region ends with the barrier that says all "
                          "columns are in place") if fused else ("none (one GPU)" if world == 1 else "NCCL all_gather_into_tensor after the run"),
            "value_with_nccl_allgather_after_the_run": (WALKERS_TOTAL * MCMC_STEPS / mc_nccl_s) if mc_nccl_s else None,
            "d2h_bytes": chain_bytes + (int(MCMC_STEPS * WALKERS_TOTAL * 8) if world == 1 else 0),
            "roofline": {"bound": "fp64_fma", "kernel": "ensemble_kernel (wide unit), c5: N=16384, d=20",
                         "achieved": mc_ach, "peak": fma_peak.value, "unit": "TFLOP/s (per GPU)",
                         "frac": mc_ach / fma_peak.value, "frac_of_nominal_40": mc_ach / NOMINAL_FP64_TFLOPS,
                         "flops_per_walker_step": fl_ws, "traffic": None,
                         "peak_source": "FP64 FMA issue-rate microbench measured in this run"}}

    extra, kernels = None, None
    if not (args.no_kernel_table or world > 1):
        gp5 = lp5 = es = None
        torch.cuda.empty_cache()
        kernels = kernel_table(lib, peak.value)
        extra = {"c2": c2_extra(lib, fma_peak_tflops=fma_peak.value)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(steps=2, warmup=1, alpha5=alpha5_host)
        cpu = {"value": r["pred"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "mcmc_walker_steps_per_s": r.get("mcmc")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(M_TOTAL * d4 * 8),
                    "d2h_bytes_per_step": int(M_TOTAL * 16), "host_buffers": "pageable NumPy arrays (what a reference user passes)",
                    "api": "GP.predict(y, X*, return_var=True)"},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "mcmc": mcmc, "train": {"c4": train4, "c5": train5},
            "argmin_last_step": {"value": winners[-1][0], "global_index": winners[-1][1]},
            "extra": extra, "kernels": kernels}
    sys.stdout.flush()
    os.dup2(_saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
