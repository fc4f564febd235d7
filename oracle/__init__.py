"""CPU oracle for the alabi surrogate hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy/SciPy FP64 restatement of the algorithms the
reference (jbirky/alabi, mounted at /root/reference) runs on its hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``alabi_b200/`` imports it and the product path fails loudly when the CUDA
library is missing.

Parity status
-------------
* ``oracle.utility`` (BAPE / AGP / Jones, ``logsubexp``, ``lnprior_uniform``,
  ``prior_transform_uniform``, ``regularization_term/gradient``,
  ``estimate_burnin``) and ``oracle.benchmarks`` are PINNED: they are checked
  in ``tests/test_oracle_golden.py`` against vectors produced by importing the
  reference's own pure-NumPy functions (``tests/golden/make_golden.py``).
* ``oracle.gp`` (george 0.4.x semantics) and ``oracle.emcee`` (emcee 3.x
  stretch move) restate third-party packages that are NOT vendored in
  /root/reference and are not installed in the build image (no network):
  **parity unpinned** for these two modules.  They are anchored on the
  reference's call sites (cited per function) and on self-consistency checks
  (finite differences, closed forms, detailed balance); see DESIGN.md.
  ``tests/test_reference_packages.py`` compares them with george / emcee /
  dynesty themselves wherever those packages can be imported (it skips here).
* ``oracle.nested`` replays the device's constrained random walks (dynesty's
  ``rwalk`` replacement step; dynesty unpinned and not installed): **parity
  unpinned** against dynesty, draw layout restated from include/alabi_b200.h.
* The reference's OWN code around the GP is pinned through this oracle: ``tests/golden/
  make_hostlogic_golden.py`` runs the reference's ``SurrogateModel.init_gp`` / ``_opt_gp`` / ``_fit_gp`` /
  ``surrogate_log_likelihood`` / ``lnprob`` / cached likelihood / ``eval_gp_at_iteration``, its k-fold CV
  worker and candidate clouds and ``ut.minimize_objective`` with ``george`` replaced by a shim on
  ``oracle.gp``; ``tests/test_host_logic.py`` makes ``alabi_b200.core`` reproduce those numbers on the same
  oracle GP.  That pins alabi's logic GIVEN the george semantics restated here; the semantics themselves
  stay unpinned as said above.
* The MATHEMATICS of the benchmarked configurations c1-c4 (kernel matrix,
  Cholesky, alpha, log-likelihood, mean, variance) is additionally pinned by
  extended-precision references (``tests/golden/make_extended.py`` ->
  ``tests/golden/extended_c*.npz``), against which both this oracle and the
  device path are measured.
"""
