"""Quasiseparable reference goldens and the reference's quasisep == direct consistency test on the CUDA path.  They
exercise `b200gp_qs_condition` and the GramBack scan, whose fine-grained first-run tests are in test_zz_first_run_gpu.py
(which therefore sorts before this file)."""

import numpy as np
import pytest

import test_host_layer_reference_tests as _cpu
from test_reference_golden import CASES, GOLD, compare, product_namespace, refcases

QS = [c for c in CASES if c["kind"] == "quasisep"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", QS, ids=[c["name"] for c in QS])
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


@pytest.mark.gpu
@pytest.mark.parametrize("parallel", [False, True], ids=["sequential", "parallel"])
@pytest.mark.parametrize("pair", range(len(_cpu.KERNEL_PAIRS)))
def test_consistent_with_direct(pair, parallel):
    """tests/test_solvers/test_quasisep/test_solver.py:62-103 of the reference, real backend"""
    _cpu.test_consistent_with_direct(pair, parallel)      # called directly: the CPU module's mock fixture does not apply
