"""Extended-precision reference values for the BENCHMARKED configurations at the reference's
default white noise (-12), where the covariance matrix is too ill-conditioned for two FP64
implementations to agree to 1e-9 (kappa * eps ~ 1e-4 .. 1e-2).

    python tests/golden/make_extended.py c1 c2 c3 c4      # minutes to ~half an hour (c4), offline

For each configuration of ``alabi_b200.workloads`` this script builds the covariance matrix
K = (amp / ndim) k(X, X) + exp(white_noise) I in x87 extended precision (``numpy.longdouble``,
64-bit mantissa, eps = 1.1e-19; kernel values through expl / sqrtl), factorises it with a blocked
left-looking Cholesky written here, and evaluates at 256 query points

    alpha = K^-1 (y - mean),  logL,  mu = k*^T alpha + mean,  sigma^2 = amp' - |L^-1 k*|^2

It is the arbiter, not the oracle: its error is kappa * 1e-19, three orders below what any FP64
path (LAPACK or the device) can reach, so |device - truth| and |oracle - truth| can be compared
(tests/test_gpu_benchmark_configs.py).  Query points: 64 training points, 64 training points
shifted by 1e-3 of the box, 128 uniform points in the box.  The condition number is computed
in FP64 from the extreme eigenvalues.  Stored per config in ``extended_<name>.npz`` together
with a SHA-256 of the inputs so a test can prove it rebuilt the same X, y.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from alabi_b200 import workloads  # noqa: E402

LD = np.longdouble
HERE = os.path.dirname(os.path.abspath(__file__))


def radial_ld(kind, r2):
    if kind == "ExpSquaredKernel":
        return np.exp(-LD(0.5) * r2)
    if kind == "Matern32Kernel":
        r = np.sqrt(LD(3.0) * r2)
        return (LD(1.0) + r) * np.exp(-r)
    r = np.sqrt(LD(5.0) * r2)
    return (LD(1.0) + r + r * r / LD(3.0)) * np.exp(-r)


def kernel_ld(kind, x1, x2, log_M, amp):
    inv_M = np.exp(-log_M.astype(LD))
    r2 = np.zeros((len(x1), len(x2)), dtype=LD)
    for k in range(x1.shape[1]):
        dk = x1[:, k, None].astype(LD) - x2[None, :, k].astype(LD)
        r2 += dk * dk * inv_M[k]
    return amp * radial_ld(kind, r2)


def cholesky_ld(A, nb=96):
    """Lower Cholesky factor in extended precision (left-looking, blocked)."""
    n = len(A)
    L = A.copy()
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0:
            L[j0:, j0:j1] -= L[j0:, :j0] @ L[j0:j1, :j0].T
        for j in range(j0, j1):
            if j > j0:
                L[j:, j] -= L[j:, j0:j] @ L[j, j0:j]
            if not L[j, j] > 0:
                raise np.linalg.LinAlgError(f"pivot {j}")
            L[j, j] = np.sqrt(L[j, j])
            L[j + 1:, j] /= L[j, j]
    return np.tril(L)


def solve_lower_ld(L, B, nb=96):
    """Z = L^-1 B (B: n x m)."""
    n = len(L)
    Z = B.astype(LD).copy()
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0:
            Z[j0:j1] -= L[j0:j1, :j0] @ Z[:j0]
        for j in range(j0, j1):
            if j > j0:
                Z[j] -= L[j, j0:j] @ Z[j0:j]
            Z[j] /= L[j, j]
    return Z


def solve_upper_ld(L, B, nb=96):
    """Z = L^-T B."""
    n = len(L)
    Z = B.astype(LD).copy()
    for j1 in range(n, 0, -nb):
        j0 = max(j1 - nb, 0)
        if j1 < n:
            Z[j0:j1] -= L[j1:, j0:j1].T @ Z[j1:]
        for j in range(j1 - 1, j0 - 1, -1):
            if j < j1 - 1:
                Z[j] -= L[j + 1:j1, j] @ Z[j + 1:j1]
            Z[j] /= L[j, j]
    return Z


def queries(cfg, nq=256):
    rng = np.random.default_rng(1000 + int(cfg["name"][1:]))
    X, b = cfg["X"], cfg["bounds"]
    idx = rng.choice(len(X), size=nq // 2, replace=False)
    near = X[idx[nq // 4:]] + 1e-3 * (b[:, 1] - b[:, 0]) * rng.uniform(-1, 1, size=(nq // 2 - nq // 4, X.shape[1]))
    uni = rng.uniform(b[:, 0], b[:, 1], size=(nq - nq // 2, X.shape[1]))
    return np.ascontiguousarray(np.vstack([X[idx[:nq // 4]], near, uni]))


def input_hash(cfg, xq):
    h = hashlib.sha256()
    for a in (cfg["X"], cfg["y"], xq, cfg["hp"]["log_M"]):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    h.update(np.array([cfg["hp"]["amp"], cfg["hp"]["mean"], cfg["hp"]["white_noise"]]).tobytes())
    return h.hexdigest()


def run(name):
    t0 = time.time()
    cfg = workloads.make_config(name)
    X, y, hp, d = cfg["X"], cfg["y"], cfg["hp"], cfg["ndim"]
    n = len(X)
    # george: kernel * c -> ConstantKernel(log(c / ndim)); the FP64 paths use amp = exp(log(c / ndim)),
    # so exactly that double is the amplitude here
    amp = LD(float(np.exp(np.log(hp["amp"] / d))))
    wn = LD(float(np.exp(hp["white_noise"])))
    xq = queries(cfg)
    K = kernel_ld(cfg["kind"], X, X, hp["log_M"], amp)
    K[np.diag_indices(n)] += wn
    # condition number in FP64 (extreme eigenvalues)
    from scipy.linalg import eigvalsh
    Kd = K.astype(np.float64)
    lam_max = float(eigvalsh(Kd, subset_by_index=[n - 1, n - 1])[0])
    lam_min = float(eigvalsh(Kd, subset_by_index=[0, 0])[0])
    del Kd
    print(f"{name}: kernel matrix built ({time.time() - t0:.0f} s), lambda {lam_min:.3e} .. {lam_max:.3e}", flush=True)
    L = cholesky_ld(K)
    del K
    print(f"{name}: factorised ({time.time() - t0:.0f} s)", flush=True)
    r = (y.astype(LD) - LD(hp["mean"]))[:, None]
    z = solve_lower_ld(L, r)
    alpha = solve_upper_ld(L, z)[:, 0]
    logdet = LD(2.0) * np.sum(np.log(np.diag(L)))
    logl = -LD(0.5) * np.sum(z * z) - LD(0.5) * logdet - LD(0.5) * LD(n) * np.log(LD(2.0) * LD(np.pi))
    Ks = kernel_ld(cfg["kind"], X, xq, hp["log_M"], amp)          # n x nq
    mu = Ks.T @ alpha + LD(hp["mean"])
    V = solve_lower_ld(L, Ks)
    var = amp - np.sum(V * V, axis=0)
    out = os.path.join(HERE, f"extended_{name}.npz")
    np.savez(out, xq=xq, mu=mu.astype(np.float64), var=var.astype(np.float64), logl=np.float64(logl),
             logdet=np.float64(logdet), alpha_norm=np.float64(np.sqrt(np.sum(alpha * alpha))),
             alpha_head=alpha[:64].astype(np.float64), lam_min=lam_min, lam_max=lam_max,
             kappa=max(lam_max, 0.0) / max(lam_min, 1e-300), amp=np.float64(amp), n=n,
             sha256=input_hash(cfg, xq))
    print(f"{name}: done in {time.time() - t0:.0f} s, kappa = {lam_max / lam_min:.3e}, logL = {float(logl):.12g} -> {out}",
          flush=True)


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["c1", "c2"]:
        run(nm)
