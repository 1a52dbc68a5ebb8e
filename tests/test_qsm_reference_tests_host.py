"""tests/qsm_reference_tests.py (the reference's test_core.py restated) on the CPU: the real Python classes over the host
build of the device source (tests/qsmhost.py), two chunk lengths."""

import pytest

import qsm_reference_tests as R
from qsmhost import HostBackend
from tinygp_b200.solvers.quasisep import core


@pytest.fixture(scope="module")
def backend():
    return HostBackend()


@pytest.fixture(autouse=True, params=[0, 5], ids=["auto-chunk", "chunk5"])
def _use_host_build(backend, monkeypatch, request):
    monkeypatch.setattr(core, "_backend", lambda: backend)
    backend.set_option("qsm_chunk", request.param)
    yield
    backend.set_option("qsm_chunk", 0)


NAMES = ["random", "celerite"]
PAR = [False, True]


def test_quasisep_def():
    R.check_quasisep_def()


@pytest.mark.parametrize("parallel", PAR)
@pytest.mark.parametrize("name", NAMES)
def test_strict_tri_matmul(name, parallel):
    R.check_strict_tri_matmul(name, parallel)


@pytest.mark.parametrize("parallel", PAR)
@pytest.mark.parametrize("name", NAMES)
def test_tri_matmul(name, parallel):
    R.check_tri_matmul(name, parallel)


@pytest.mark.parametrize("parallel", PAR)
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("symm", [True, False])
def test_square_matmul(symm, name, parallel):
    R.check_square_matmul(symm, name, parallel)


def test_tri_inv():
    R.check_tri_inv()


@pytest.mark.parametrize("parallel", PAR)
def test_tri_solve(parallel):
    R.check_tri_solve(parallel)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("symm", [True, False])
def test_square_inv(symm, name):
    R.check_square_inv(symm, name)


@pytest.mark.parametrize("name", NAMES)
def test_gram(name):
    R.check_gram(name)


@pytest.mark.parametrize("parallel", PAR)
def test_cholesky(parallel):
    R.check_cholesky(parallel)


def test_tri_qsmul():
    R.check_tri_qsmul()


def test_square_qsmul():
    R.check_square_qsmul()


def test_ops():
    R.check_ops()
