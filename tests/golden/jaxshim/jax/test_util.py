import pytest


def check_grads(*_a, **_k):
    pytest.skip("no autodiff in the NumPy stand-in for jax")
