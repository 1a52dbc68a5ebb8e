"""The CUDA path against the goldens produced by the reference itself (tests/golden/reference_vectors.json; see
tests/test_reference_golden.py for the CPU half and how the goldens are made).  Sorted late (`zy`) because these cases
were written after round 1's GPU minutes were spent: under the driver's `pytest -x` a first-run surprise here must not
hide the results of the files before it."""

import numpy as np
import pytest

from test_reference_golden import CASES, GOLD, IDS, compare, product_namespace, refcases  # noqa: F401

DENSE = [c for c in CASES if c["kind"] == "dense"]      # the quasiseparable cases run in test_zzy_* (they need first-run code)


@pytest.mark.gpu
@pytest.mark.parametrize("case", DENSE, ids=[c["name"] for c in DENSE])
def test_cuda_path_matches_reference(case):
    got = refcases.run_case(product_namespace(), case)
    compare(got, GOLD["cases"][case["name"]], case["name"])
