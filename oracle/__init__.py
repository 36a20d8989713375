"""CPU oracle for the BayesGP hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy FP64 restatement of the reference's algorithm for
the path named in BASELINE.json (latent-Gaussian log-posterior -> inner
Newton/Laplace at every AGHQ node -> posterior sampling -> predict).  It exists
to *check* the CUDA product in ``bayesgp_b200``; it is never the thing shipped
or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Where the arithmetic lives in the reference
-------------------------------------------
* ``/root/reference/src/BayesGP.cpp:30-253``   the TMB objective template
* ``/root/reference/R/02_model_fit.R:1-306``   data marshalling (tmbdat layout)
* ``/root/reference/R/01_utility.R:67-440``    basis / penalty constructors
* ``/root/reference/R/03_post_fit.R:53-296``   predict / sample -> function
* third-party, un-vendored (absent from /root/reference): TMB (Laplace, inner
  newton), aghq >= 0.4.1 (optimize_theta, normalize_logpost, sample_marginal),
  numDeriv (Richardson jacobian), mvQuad (Gauss-Hermite grid), stats::optim
  (vmmin BFGS), stats::quantile (type 7).  Their published algorithms are
  restated in ``oracle/aghq.py`` / ``oracle/laplace.py`` following SURVEY.md
  Appendix A; every function cites the call site in the reference it serves.

Parity pin
----------
The reference's own tests pin nothing on this path (parser-only testthat file).
The only external known-answer values are the rendered README fit
(``/root/reference/README.md:71-96``); ``tests/test_oracle_readme.py`` checks
the oracle against them (latent dimension, log normalising constant to printed
precision, theta posterior mean/sd given the README's grid centre/scale, fixed
effect moments within Monte-Carlo error).  R/TMB/aghq are not installed in the
build container, so no reference-generated fixtures beyond that exist:
**parity is pinned by the README printout only** (see DESIGN.md section 3).
"""
