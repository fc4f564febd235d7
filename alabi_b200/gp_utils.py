"""GP configuration and hyper-parameter search with the semantics of
alabi/gp_utils.py, driving the GPU GP (``alabi_b200.GP``) instead of george.

* ``configure_gp``                       alabi/gp_utils.py:170-248
* ``regularization_term / _gradient``    alabi/gp_utils.py:30-108
* ``optimize_gp`` (ML restarts)          alabi/gp_utils.py:251-447 (+ ``_nll`` / ``_grad_nll`` :111-167)
* ``weighted_mse_by_probability``        alabi/gp_utils.py:450-508
* ``optimize_gp_kfold_cv`` (3 stages)    alabi/gp_utils.py:511-1231
* stage-2 / stage-3 candidate clouds     alabi/gp_utils.py:1234-1367
"""
import copy

import numpy as np

from .gp import GP

__all__ = ["configure_gp", "optimize_gp", "optimize_gp_kfold_cv", "regularization_term", "regularization_gradient",
           "weighted_mse_by_probability"]


def regularization_term(hparams, lengthscale_indices, amp_0=1.0, mu_0=1.0, sigma_0=2.0):
    """Negative log of a LogNormal(mu_0 + log sqrt(len(hparams)), sigma_0) prior
    summed over the length-scale entries (the reference scales with the length
    of the hyper-vector, not the problem dimension; kept)."""
    hparams = np.asarray(hparams, dtype=np.float64)
    ll = hparams[lengthscale_indices]
    mu = mu_0 + 0.5 * np.log(len(hparams))
    return amp_0 * np.sum(ll + 0.5 * np.log(2 * np.pi * sigma_0 ** 2) + (ll - mu) ** 2 / (2 * sigma_0 ** 2))


def regularization_gradient(hparams, lengthscale_indices, amp_0=1.0, mu_0=1.0, sigma_0=2.0):
    """(1 + (l - mu) / sigma_0^2) / exp(l) on the length-scale slots, zero elsewhere
    (the reference differentiates w.r.t. the linear scale; kept)."""
    hparams = np.asarray(hparams, dtype=np.float64)
    g = np.zeros_like(hparams)
    ll = hparams[lengthscale_indices]
    mu = mu_0 + 0.5 * np.log(len(hparams))
    g[lengthscale_indices] = (1.0 + (ll - mu) / sigma_0 ** 2) / np.exp(ll)
    return amp_0 * g


def configure_gp(theta, y, kernel, fit_amp=True, fit_mean=True, fit_white_noise=False, white_noise=-12,
                 hyperparameters=None, device=None):
    """Build the GP (amplitude = var(y), mean = median(y)), optionally set a
    hyper-vector, and factorise.  Returns None when the covariance matrix is
    not positive definite."""
    theta = np.asarray(theta, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if np.any(~np.isfinite(theta)):
        raise ValueError("All theta values must be finite!")
    if np.any(~np.isfinite(y)):
        raise ValueError("All y values must be finite!")
    if fit_amp:
        kernel = kernel * np.var(y)
    else:
        kernel = copy.deepcopy(kernel)
    gp = GP(kernel=kernel, fit_mean=fit_mean, mean=np.median(y), white_noise=white_noise,
            fit_white_noise=fit_white_noise, device=device)
    if hyperparameters is not None:
        if np.any(~np.isfinite(hyperparameters)):
            raise ValueError("All hyperparameter values must be finite!")
        gp.set_parameter_vector(hyperparameters)
    try:
        gp.compute(theta)
    except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all
        print(f"configure_gp error: {e}")
        return None
    return gp


def _nll(p, gp, y, gp_hyper_prior):
    """Negative log marginal likelihood at hyper-vector ``p``; inf outside the hyper-prior
    or when K(p) does not factorise (alabi/gp_utils.py:111-140)."""
    if not np.isfinite(gp_hyper_prior(p)):
        return np.inf
    try:
        gp.set_parameter_vector(p)
    except np.linalg.LinAlgError:
        return np.inf
    ll = gp.log_likelihood(y, quiet=True)
    return -ll if np.isfinite(ll) else np.inf


def _grad_nll(p, gp, y):
    """Gradient of ``_nll`` (alabi/gp_utils.py:143-167); one factorisation + K^-1 on the device."""
    gp.set_parameter_vector(p)
    return -gp.grad_log_likelihood(y, quiet=True)


def optimize_gp(gp, _theta, _y, gp_hyper_prior, p0, bounds=None, method="l-bfgs-b", optimizer_kwargs=None,
                regularize=True, amp_0=1.0, mu_0=1.0, sigma_0=2.0, lengthscale_indices=None):
    """Maximum-marginal-likelihood hyper-parameters from one start (``p0`` 1-D) or several
    (``p0`` 2-D: the restart with the largest log-likelihood wins), alabi/gp_utils.py:251-447.
    Every objective / gradient evaluation is a device factorisation (K1 + K2); scipy only
    drives.  Failed restarts and a failed single run leave the initial vector in place."""
    from scipy.optimize import minimize
    import warnings
    init_hp = np.array(gp.get_parameter_vector(), dtype=np.float64)
    _y = np.asarray(_y, dtype=np.float64).reshape(-1)
    p0 = np.asarray(p0, dtype=np.float64)
    if regularize and lengthscale_indices is None:
        # as in the reference, positions are taken in the KERNEL's name list (they are then
        # applied to the full GP vector; pass lengthscale_indices to address it exactly)
        lengthscale_indices = [i for i, nm in enumerate(gp.kernel.get_parameter_names()) if "metric:log_m" in nm.lower()]
        if len(lengthscale_indices) == 0:
            warnings.warn("Could not infer lengthscale indices from kernel; regularisation is not applied.")
            regularize = False
    if method not in ("newton-cg", "bfgs", "l-bfgs-b", "powell", "nelder-mead"):
        print(f"Warning: {method} not a valid method. Using 'l-bfgs-b' optimizer instead.")
        method = "l-bfgs-b"
    if method == "bfgs":
        bounds = None
    reg = (lambda p: regularization_term(p, lengthscale_indices, amp_0=amp_0, mu_0=mu_0, sigma_0=sigma_0)) \
        if regularize else (lambda p: 0.0)
    reg_grad = (lambda p: regularization_gradient(p, lengthscale_indices, amp_0=amp_0, mu_0=mu_0, sigma_0=sigma_0)) \
        if regularize else (lambda p: 0.0)
    obj = lambda p: _nll(p, gp, _y, gp_hyper_prior) + reg(p)
    jac = (lambda p: _grad_nll(p, gp, _y) + reg_grad(p)) if method in ("newton-cg", "l-bfgs-b") else None

    def restore():
        gp.set_parameter_vector(init_hp)
        gp.recompute()

    if p0.ndim > 1 and p0.shape[0] > 1:
        found, mll = [], []
        for i, x0 in enumerate(p0):
            try:
                r = minimize(obj, x0, method=method, jac=jac, bounds=bounds, options=optimizer_kwargs)
                if np.isfinite(gp_hyper_prior(r.x)):
                    gp.set_parameter_vector(r.x)
                    found.append(np.array(r.x))
                    mll.append(gp.log_likelihood(_y, quiet=True))
                    continue
                print(f"\nWarning: GP hyperparameter optimization restart {i} failed. Solution failed prior bounds.\n")
            except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all
                print(f"\nWarning: GP hyperparameter optimization restart {i} failed with error: {e}\n")
            found.append(init_hp)
            mll.append(-np.inf)
        if len(mll) > 0 and max(mll) > -np.inf:
            gp.set_parameter_vector(found[int(np.argmax(mll))])
            gp.recompute()
        else:
            print("\nWarning: All hyperparameter optimizations failed. Using initial values.\n")
            restore()
        return gp
    try:
        r = minimize(obj, p0.reshape(-1), method=method, jac=jac, bounds=bounds, options=optimizer_kwargs)
        if r.success and np.isfinite(gp_hyper_prior(r.x)):
            gp.set_parameter_vector(r.x)
            gp.recompute()
        else:
            print("\nWarning: GP hyperparameter optimization failed. Using initial values.\n")
            restore()
    except Exception as e:  # noqa: BLE001
        print(f"\nWarning: GP hyperparameter optimization failed with error: {e}. Using initial values.\n")
        restore()
    return gp


def weighted_mse_by_probability(y_true, y_pred, weight_method="exponential", temperature=1.0):
    """MSE weighted towards high-likelihood points (alabi/gp_utils.py:450-508)."""
    return _weighted_mse(np.asarray(y_true), np.asarray(y_pred), weight_method, temperature)


def _weighted_mse(y_true, y_pred, method="exponential", temperature=1.0):
    if method == "exponential":
        w = np.exp(y_true / temperature)
    elif method == "linear":
        w = y_true - np.min(y_true) + 1e-6
    elif method == "softmax":
        w = np.exp(y_true / temperature)
        w = w / np.sum(w) * len(w)
    elif method == "rank":
        w = np.argsort(np.argsort(y_true)) + 1
    else:
        raise ValueError(f"Unknown weight_method: {method}")
    w = w / np.mean(w)
    return np.average((y_true - y_pred) ** 2, weights=w)


def _fold_score(scoring, y_val, y_pred, wmethod, wfactor):
    if scoring == "mse":
        return float(np.mean((y_val - y_pred) ** 2))
    if scoring == "mae":
        return float(np.mean(np.abs(y_val - y_pred)))
    if scoring == "r2":
        ss_res, ss_tot = np.sum((y_val - y_pred) ** 2), np.sum((y_val - np.mean(y_val)) ** 2)
        return float(-(1.0 - ss_res / ss_tot))
    if scoring == "weighted_mse":
        return float(_weighted_mse(y_val, y_pred, wmethod, wfactor))
    raise ValueError(f"Unsupported scoring method: {scoring}")


def _evaluate_candidates(gp, theta, y, y_scaler, cands, k_folds, scoring, wmethod, wfactor, batched=True,
                         random_state=None):
    """scores[cand, fold]; inf for failed folds.  Every (candidate, fold) is one factorise +
    log-likelihood + mean-predict job; a fresh shuffled ``KFold`` split is drawn per candidate like
    the reference's worker does (alabi/gp_utils.py:532; ``random_state`` makes the splits
    reproducible: candidate c uses ``random_state + c``).  ``batched=True`` (default) hands ALL jobs
    of the stage to one batched device call (``GP.cv_batch``); ``False`` runs them one after
    another on one handle (the pre-batching path, kept for cross-checks)."""
    from sklearn.model_selection import KFold
    scores = np.full((len(cands), k_folds), np.inf)
    ok_c = [ci for ci, hp in enumerate(cands) if np.all(np.isfinite(hp))]
    folds = {}
    for ci in ok_c:
        kf = KFold(n_splits=k_folds, shuffle=True, random_state=None if random_state is None else int(random_state) + ci)
        folds[ci] = list(kf.split(theta))

    def score(ci, fi, ll, pred, va):
        if not np.isfinite(ll) or len(pred) == 0 or not np.all(np.isfinite(pred)):
            return np.inf
        y_val = y_scaler.inverse_transform(y[va].reshape(-1, 1)).flatten()
        y_pred = y_scaler.inverse_transform(np.asarray(pred).reshape(-1, 1)).flatten()
        try:
            return _fold_score(scoring, y_val, y_pred, wmethod, wfactor)
        except Exception:  # noqa: BLE001 - a failed fold scores inf, like the reference
            return np.inf

    if batched and hasattr(gp, "cv_batch") and ok_c:
        jobs = [(ci, fi) for ci in ok_c for fi in range(k_folds)]
        try:
            preds, lls, status = gp.cv_batch(theta, y, np.asarray(cands)[ok_c], [folds[ci][fi] for ci, fi in jobs])
        except Exception as e:  # noqa: BLE001 - e.g. out of device memory: fall through to the serial path
            print(f"CV: batched evaluation failed ({e}); evaluating the candidates one by one")
        else:
            # the y scaler is applied ONCE to all targets and ONCE to all predictions (875 jobs would
            # otherwise make 1750 sklearn transformer calls)
            y_true = y_scaler.inverse_transform(y.reshape(-1, 1)).flatten()
            sizes = [len(p) for p in preds]
            flat = np.concatenate(preds) if sum(sizes) else np.empty(0)
            good = np.isfinite(flat)
            flat_t = np.full(len(flat), np.nan)
            if good.any():
                flat_t[good] = y_scaler.inverse_transform(flat[good].reshape(-1, 1)).flatten()
            off = np.concatenate([[0], np.cumsum(sizes)])
            for b, (ci, fi) in enumerate(jobs):
                pb = flat_t[off[b]:off[b + 1]]
                if status[b] != 0 or not np.isfinite(lls[b]) or len(pb) == 0 or not np.all(np.isfinite(pb)):
                    continue
                try:
                    scores[ci, fi] = _fold_score(scoring, y_true[folds[ci][fi][1]], pb, wmethod, wfactor)
                except Exception:  # noqa: BLE001 - a failed fold scores inf, like the reference
                    scores[ci, fi] = np.inf
            return scores
    work = copy.copy(gp)
    for ci in ok_c:
        for fi, (tr, va) in enumerate(folds[ci]):
            try:
                work.set_parameter_vector(cands[ci])
                work.compute(theta[tr])
                ll = work.log_likelihood(y[tr])
                if not np.isfinite(ll):
                    raise ValueError("invalid log-likelihood")
                pred = work.predict(y[tr], theta[va], return_var=False, return_cov=False)
                scores[ci, fi] = score(ci, fi, ll, pred, va)
            except Exception:  # noqa: BLE001 - a failed fold scores inf, like the reference
                scores[ci, fi] = np.inf
    return scores


def _evaluate_candidate_worker(args):
    """One candidate in the reference's worker format (alabi/gp_utils.py:511-637):
    ``(cand_idx, hyperparams, gp, _theta, _y, y_scaler, k_folds, scoring, weighted_mse_method, weighted_mse_factor)``
    -> ``(cand_idx, fold_scores, "success" | message)``; the batched path above is what the search itself uses."""
    cand_idx, hp, gp, _theta, _y, y_scaler, k_folds, scoring, wmethod, wfactor = args
    try:
        hp = np.asarray(hp, dtype=np.float64)
        if not np.all(np.isfinite(hp)):
            return (cand_idx, None, "Invalid hyperparameters (NaN/Inf)")
        scores = list(_evaluate_candidates(gp, _theta, _y, y_scaler, [hp], k_folds, scoring, wmethod, wfactor, batched=False)[0])
        if all(np.isinf(scores)):
            return (cand_idx, scores, "All folds failed.")
        return (cand_idx, scores, "success")
    except Exception as e:  # noqa: BLE001 - the reference returns the message
        return (cand_idx, None, str(e))


def _mean_scores(scores):
    out = np.full(len(scores), np.inf)
    for i, row in enumerate(scores):
        ok = row[np.isfinite(row)]
        if len(ok):
            out[i] = np.mean(ok)
    return out


def _generate_stage_candidates(best_params, n_candidates, width_factor, gp=None):
    """Gaussian cloud of width ``width_factor`` around ``best_params`` (first
    candidate = the centre).  Length scales move together only if the entries
    the reference inspects are all equal (it indexes the full vector with
    kernel-local positions; kept)."""
    best = np.asarray(best_params, dtype=np.float64)
    npar = len(best)
    idx = None
    if gp is not None:
        idx = [i for i, nm in enumerate(gp.kernel.get_parameter_names()) if "metric:log_m" in nm.lower()]
    if idx is not None and len(idx) > 1:
        idx = [i for i in idx if i < npar]
        uniform = np.allclose(best[idx], best[idx[0]])
    elif npar > 2:
        idx = list(range(2, npar))
        uniform = np.allclose(best[2:], best[2])
    else:
        idx, uniform = [], False
    cands = [best.copy()]
    for _ in range(int(n_candidates) - 1):
        if uniform and len(idx) > 0:
            c = best.copy()
            for j in range(npar):
                if j not in idx:
                    c[j] += np.random.normal(0, width_factor)
            c[idx] += np.random.normal(0, width_factor)
        else:
            c = best + np.random.normal(0, width_factor, npar)
        cands.append(c)
    return np.array(cands)


_generate_stage2_candidates = _generate_stage_candidates
_generate_stage3_candidates = _generate_stage_candidates


def optimize_gp_kfold_cv(gp, _theta, _y, hyperparameter_candidates, y_scaler, k_folds=5, scoring="mse", pool=None,
                         stage2_candidates=None, stage2_width=0.5, stage3_candidates=None, stage3_width=0.2,
                         weighted_mse_method="exponential", weighted_mse_factor=1.0, verbose=True, batched=True,
                         random_state=None, shard_candidates=False):
    """Pick the hyper-vector with the best mean k-fold validation score out of
    the given candidates, then refine around the winner with up to two
    Gaussian candidate clouds.  Returns the GP set to the winner and computed
    on all data (None if every candidate fails).  ``pool`` is accepted for
    signature compatibility: every stage (candidates x folds jobs) is ONE batched device call
    (``GP.cv_batch`` -> ``ab_gp_cv_batch``); ``batched=False`` evaluates the jobs one after another,
    ``random_state`` makes the shuffled splits reproducible."""
    theta = np.asarray(_theta, dtype=np.float64)
    y = np.asarray(_y, dtype=np.float64)
    if theta.ndim == 1:
        theta = theta.reshape(-1, 1)
    if y.ndim != 1:
        y = y.squeeze()
    cands = np.asarray(hyperparameter_candidates, dtype=np.float64)
    if cands.ndim == 1:
        cands = cands.reshape(1, -1)
    n = len(theta)
    if len(y) != n:
        raise ValueError(f"_theta and _y must have same length, got {len(theta)} and {len(y)}")
    if n < k_folds:
        raise ValueError(f"Number of samples ({n}) must be >= k_folds ({k_folds})")
    if k_folds < 2:
        raise ValueError(f"k_folds must be >= 2, got {k_folds}")
    two_stage = stage2_candidates is not None
    three_stage = stage3_candidates is not None

    stage = [0]

    def run(c):
        # one batched device call per stage (all candidates x folds); ``random_state`` fixes the splits
        rs = None if random_state is None else int(random_state) + 100003 * stage[0]
        stage[0] += 1
        from . import parallel as par
        if shard_candidates and par.world_size() > 1:
            # candidates sharded over the ranks, one all_gather of the per-fold scores;
            # the splits are seeded by the GLOBAL candidate index, so the scores equal a one-rank run
            c = np.asarray(par.broadcast_object(np.asarray(c)))
            seed0 = par.broadcast_object(int(np.random.randint(2 ** 31)) if rs is None else rs)

            def part(idx):
                sc = np.full((len(idx), k_folds), np.inf)
                for row, ci in enumerate(idx):
                    sc[row] = _evaluate_candidates(gp, theta, y, y_scaler, c[ci:ci + 1], k_folds, scoring, weighted_mse_method,
                                                   weighted_mse_factor, batched=batched, random_state=seed0 + int(ci))[0]
                return sc
            return _mean_scores(par.sharded_rows(part, len(c)))
        return _mean_scores(_evaluate_candidates(gp, theta, y, y_scaler, c, k_folds, scoring, weighted_mse_method,
                                                 weighted_mse_factor, batched=batched, random_state=rs))

    s1 = run(cands)
    if np.all(np.isinf(s1)):
        if verbose:
            print("CV: every stage-1 candidate failed")
        return None
    best_hp, best_score = cands[int(np.argmin(s1))], float(np.min(s1))
    if verbose:
        print(f"CV stage 1: best {scoring} = {best_score:.6g}")
    if two_stage:
        c2 = _generate_stage2_candidates(best_hp, stage2_candidates, stage2_width, gp=gp)
        s2 = run(c2)
        if not np.all(np.isinf(s2)) and np.min(s2) < best_score:
            best_hp, best_score = c2[int(np.argmin(s2))], float(np.min(s2))
        if verbose:
            print(f"CV stage 2: best {scoring} = {best_score:.6g}")
        if three_stage:
            c3 = _generate_stage3_candidates(best_hp, stage3_candidates, stage3_width, gp=gp)
            s3 = run(c3)
            if not np.all(np.isinf(s3)) and np.min(s3) < best_score:
                best_hp, best_score = c3[int(np.argmin(s3))], float(np.min(s3))
            if verbose:
                print(f"CV stage 3: best {scoring} = {best_score:.6g}")
    try:
        gp.set_parameter_vector(best_hp)
        gp.compute(theta)
    except Exception as e:  # noqa: BLE001
        if verbose:
            print(f"CV: could not set the best hyper-parameters: {e}")
    return gp
