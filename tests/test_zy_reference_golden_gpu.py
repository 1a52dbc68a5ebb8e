"""The CUDA path against the goldens produced by the reference itself (tests/golden/reference_vectors.json; see
tests/test_reference_golden.py for the CPU half and how the goldens are made).  Sorted late (`zy`) because these cases
were written after round 1's GPU minutes were spent: under the driver's `pytest -x` a first-run surprise here must not
hide the results of the files before it."""

import numpy as np
import pytest

from test_reference_golden import CASES, GOLD, IDS, compare, product_namespace, refcases  # noqa: F401

@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_path_matches_reference(case):
    got = refcases.run_case(product_namespace(), case)
    compare(got, GOLD["cases"][case["name"]], case["name"])


@pytest.mark.gpu
def test_cuda_path_unsorted_raises_like_the_reference():
    import tinygp_b200 as tg
    from tinygp_b200.kernels import quasisep
    with pytest.raises(ValueError) as e:
        tg.GaussianProcess(quasisep.Matern32(1.5), np.array([0.0, 2.0, 1.0]), diag=0.1)
    assert str(e.value) == GOLD["unsorted_raises"]
