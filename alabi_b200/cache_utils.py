"""Model cache and human-readable run reports (alabi/cache_utils.py:18-183).

``SurrogateModel.save`` pickles the model and writes ``<model_name>.txt`` next to it: a GP
section (configuration, final hyper-parameters, last test error) and, once the samplers have
run, an emcee and a dynesty section with the posterior summary statistics.  The section titles
and the ``label: value`` lines follow the reference's report so that scripts which read those
files keep working.  No MPI here: one process per GPU, every rank may load the cache itself.
"""
import os
import pickle

import numpy as np

__all__ = ["load_pickle", "load_model_cache", "write_report_gp", "write_report_emcee", "write_report_dynesty"]

_RULE = "=" * 66


def load_pickle(savedir, fname="surrogate_model.pkl"):
    with open(os.path.join(savedir, fname), "rb") as f:
        return pickle.load(f)


def load_model_cache(savedir):
    """Reload a cached ``SurrogateModel`` (alabi/cache_utils.py:27-68).  The GP inside is
    re-factorised on this process's GPU on first use."""
    try:
        return load_pickle(savedir)
    except Exception as e:  # noqa: BLE001
        print(f"Failed to load model cache: {e}")
        raise


def _section(title):
    return f"{_RULE}\n{title} \n{_RULE}\n\n"


def _block(heading, pairs):
    out = f"{heading}: \n{'-' * (len(heading) + 1)} \n"
    for label, value in pairs:
        out += f"{label}: {value} \n"
    return out + "\n"


_GP_FIELDS = (("Kernel", "kernel_name"), ("Function bounds", "bounds"), ("fit mean", "fit_mean"),
              ("fit amplitude", "fit_amp"), ("fit white_noise", "fit_white_noise"), ("GP white noise", "white_noise"),
              ("Hyperparameter bounds", "hp_bounds"), ("Active learning algorithm", "algorithm"),
              ("Number of total training samples", "ntrain"), ("Number of initial training samples", "ninit_train"),
              ("Number of active training samples", "nactive"), ("Number of test samples", "ntest"))


def write_report_gp(sm, file):
    """GP section; (re)creates ``file + ".txt"``."""
    text = _section("GP summary")
    text += _block("Configuration", [(lab, getattr(sm, att)) for lab, att in _GP_FIELDS if hasattr(sm, att)])
    text += "Results: \n-------- \nGP final hyperparameters: \n"
    for name, value in zip(sm.gp.get_parameter_names(), sm.gp.get_parameter_vector()):
        text += f"   [{name}] \t{value} \n"
    text += "\n"
    if hasattr(sm, "train_runtime"):
        text += f"Active learning train runtime (s): {np.round(sm.train_runtime)} \n\n"
    tr = getattr(sm, "training_results", None)
    if tr and len(tr.get("test_mse", [])):
        text += f"Final test error (MSE): {tr['test_mse'][-1]} \n\n"
    with open(file + ".txt", "w") as f:
        f.write(text)


def _posterior_summary(sm, samples):
    mean, std = np.mean(samples, axis=0), np.std(samples, axis=0)
    labels = getattr(sm, "labels", None) or [f"theta_{i}" for i in range(sm.ndim)]
    return "Summary statistics: \n" + "".join(f"{labels[i]} = {mean[i]} +/- {std[i]} \n" for i in range(sm.ndim)) + "\n"


def write_report_emcee(sm, file):
    """emcee section, appended to ``file + ".txt"``."""
    text = _section("emcee summary")
    text += _block("Configuration", [("Number of walkers", sm.nwalkers), ("Number of steps per walker", sm.nsteps)])
    text += "Results: \n-------- \n"
    text += "Mean acceptance fraction: {0:.3f} \n".format(sm.acc_frac)
    text += "Mean autocorrelation time: {0:.3f} steps \n".format(sm.autcorr_time)
    text += f"Burn: {sm.iburn} \nThin: {sm.ithin} \n"
    text += f"Total burned, thinned, flattened samples: {sm.emcee_samples.shape[0]} \n\n"
    text += f"emcee runtime (s): {np.round(sm.emcee_runtime)} \n\n"
    text += _posterior_summary(sm, sm.emcee_samples)
    with open(file + ".txt", "a") as f:
        f.write(text)


def write_report_dynesty(sm, file):
    """dynesty section, appended to ``file + ".txt"``."""
    text = _section("dynesty summary")
    text += "Configuration: \n-------------- \nResults: \n-------- \n"
    text += f"Total weighted samples: {sm.dynesty_samples.shape[0]} \n\n"
    text += f"Dynesty runtime (s): {np.round(sm.dynesty_runtime)} \n\n"
    text += _posterior_summary(sm, sm.dynesty_samples)
    with open(file + ".txt", "a") as f:
        f.write(text)
