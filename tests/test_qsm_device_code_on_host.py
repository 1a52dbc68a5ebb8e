"""The device source of the QSM algebra (tinygp_b200/csrc/qsm.cu + qsm_core.cuh) compiled for the CPU and driven
through the real Python classes (tinygp_b200.solvers.quasisep.core / ops) against the reference's golden outputs and the
oracle -- with several chunk lengths, so the chunk-composite / boundary-state / replay decomposition of every scan is
exercised.  What this does not cover is launch geometry and warp-level sharing: the `-m gpu` twin (test_qsm_gpu.py)."""

import pytest

import qsmchecks
from qsmhost import HostBackend
from qsmutil import qsmcases
from tinygp_b200.solvers.quasisep import core


@pytest.fixture(scope="module")
def backend():
    return HostBackend()


@pytest.fixture(params=[0, 1, 3, 7])
def ops(backend, monkeypatch, request):
    monkeypatch.setattr(core, "_backend", lambda: backend)
    backend.set_option("qsm_chunk", request.param)
    yield qsmchecks.operands()
    backend.set_option("qsm_chunk", 0)


def test_dense_forms_parts_and_scaling(ops):
    qsmchecks.check_dense_and_parts(ops)


def test_qsm_mul_every_type_pair(ops):
    qsmchecks.check_products(ops)


def test_elementwise_sum_and_product(ops):
    qsmchecks.check_sums(ops)


def test_inverses_gram_cholesky_solves(ops):
    qsmchecks.check_inverses_and_factor(ops)


@pytest.mark.parametrize("chunk", [0, 2, 5])
@pytest.mark.parametrize("case", qsmcases.CONDITION, ids=lambda c: c["name"])
def test_conditioned_covariance_generators(backend, monkeypatch, case, chunk):
    monkeypatch.setattr(core, "_backend", lambda: backend)
    backend.set_option("qsm_chunk", chunk)
    try:
        qsmchecks.check_condition_algebra(case)
    finally:
        backend.set_option("qsm_chunk", 0)


@pytest.mark.parametrize("n,m1,m2,chunk", [(300, 3, 2, 0), (300, 2, 5, 16), (1000, 4, 4, 0)])
def test_many_chunks_against_the_oracle(backend, monkeypatch, n, m1, m2, chunk):
    monkeypatch.setattr(core, "_backend", lambda: backend)
    backend.set_option("qsm_chunk", chunk)
    try:
        qsmchecks.check_against_oracle_large(n, m1, m2, seed=n + m1)
    finally:
        backend.set_option("qsm_chunk", 0)
