"""The restated reference tests that were added after the round's GPU minutes were spent (user-defined state-space kernels,
models with more than 8 states, the Conditioned kernel through the device GEMM), run with the real backend.  On the CPU they
pass over the mock C-ABI and the host build of the device QSM source (tests/test_host_layer_reference_tests.py); this file sorts
last so that under `pytest -x` nothing else is hidden if one of them meets a first-run surprise on the GPU."""

import pytest

import test_host_layer_reference_tests as _cpu
from test_host_layer_reference_tests import gp_data, kdata, random  # noqa: F401  (fixtures)
from test_zx_reference_tests_gpu import LATE, _clone

# Not gating: these have never met a GPU (see above), so a first-run surprise is reported as XFAIL and a pass as XPASS instead
# of turning the whole `-m gpu` run red for features outside the measured path.  Remove the marker after the first GPU run.
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="first GPU run of code checked on the CPU only (host build of the device source)",
                                strict=False)]

for _name in sorted(LATE):
    globals()[_name] = _clone(getattr(_cpu, _name))
