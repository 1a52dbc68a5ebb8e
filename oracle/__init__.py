"""CPU oracle for the tinygp hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a NumPy/SciPy (and, for the quasiseparable recursions, plain C)
restatement of the reference algorithm (dfm/tinygp @ 5302d5a).  It exists only
so that ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` have something to check the CUDA path
against.  Nothing under ``tinygp_b200/`` imports it.

PARITY PINNING.  The reference is pure Python on JAX; ``jax``/``equinox`` are not installed
here and there is no network, and the reference ships **no golden vectors** (its tests are all
relations at rtol=atol=5e-7, src/tinygp/test_utils.py:9-26).  The oracle is pinned three ways:

(i)   **reference-generated goldens**: ``tests/golden/reference_vectors.json`` holds the outputs
      of the UNMODIFIED reference sources (``/root/reference/src/tinygp``) executed over
      ``tests/golden/jaxshim`` -- NumPy stand-ins for the jax/equinox *array library only*
      (vmap = loop + stack, lax.scan = loop, jit = identity, linalg = LAPACK); every line of
      tinygp's own algorithm runs as written.  The stand-ins are validated by the reference's
      own test-suite (163 tests pass over them, ``tests/golden/run_reference_tests.sh``).
      34 cases: BASELINE configs 1-5 at oracle sizes, every stationary leaf / distance default /
      sum / product / transform, every state-space kernel, ties, both scan modes, non-PD -> -inf,
      log_probability / normalization / solves / condition / predict.  The oracle agrees with
      them to <= 1e-12 (``tests/test_reference_golden.py``, asserted at 1e-10).  Caveat, stated
      plainly: the third-party arithmetic underneath (XLA:CPU's libm / LAPACK / scan order) is
      NumPy/SciPy's here, not jaxlib's -- those differ at the last-ulp level only.
(ii)  the reference's own test *relations*, restated in ``tests/test_oracle_*.py`` (dense ==
      quasisep == kalman log-probability, test_solver.py:27-103, test_kalman.py:47-69; QSM
      Cholesky == dense Cholesky, test_core.py:308-323; generators == dense kernel and
      transition == expm(F^T dt), test_quasisep.py:53-72; Celerite closed form,
      test_quasisep.py:83-97) and the multivariate-normal definition via scipy.stats;
(iii) the known-answer vectors of tests/golden/known_answers.json (independent scalar-formula +
      scipy.stats formulation) and the full-size LAPACK values of tests/golden/full_size.json.
"""
