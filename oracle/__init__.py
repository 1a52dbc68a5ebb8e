"""CPU oracle for the tinygp hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a NumPy/SciPy (and, for the quasiseparable recursions, plain C)
restatement of the reference algorithm (dfm/tinygp @ 5302d5a).  It exists only
so that ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` have something to check the CUDA path
against.  Nothing under ``tinygp_b200/`` imports it.

PARITY PINNING.  The reference is pure Python on JAX; ``jax``/``equinox`` are
not installed here and there is no network, so the reference itself cannot be
imported, and the reference ships **no golden vectors** (its tests are all
relations at rtol=atol=5e-7, src/tinygp/test_utils.py:9-26).  The oracle is
therefore pinned by (i) the reference's own test *relations*, restated in
``tests/test_oracle_*.py`` (dense == quasisep == kalman log-probability,
test_solver.py:27-103, test_kalman.py:47-69; QSM Cholesky == dense Cholesky,
test_core.py:308-323; generators == dense kernel and transition == expm(F^T dt),
test_quasisep.py:53-72; Celerite closed form, test_quasisep.py:83-97) and
(ii) the multivariate-normal definition via scipy.stats, and (iii) the known-answer
vectors of tests/golden/known_answers.json, produced by tests/golden/make_golden.py
from an independent scalar-formula + scipy.stats formulation (NOT by the reference).
Absolute values are "parity unpinned" in the sense of the task statement: no
reference-produced number exists to compare against.
"""
