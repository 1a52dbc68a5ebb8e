"""Make the UNMODIFIED reference importable in this container: put the NumPy stand-ins for jax/equinox
(tests/golden/jaxshim) and /root/reference/src on sys.path.  Test infrastructure; never imported by the product."""

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("TINYGP_REFERENCE_SRC", "/root/reference/src")


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "tinygp"))


def install():
    if not available():
        raise RuntimeError(f"the reference sources are not present at {REFERENCE_SRC}")
    shim = os.path.join(HERE, "jaxshim")
    for p in (REFERENCE_SRC, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "tinygp.tinygp_version" not in sys.modules:  # written by setuptools_scm at build time; absent from a checkout
        m = types.ModuleType("tinygp.tinygp_version")
        m.__version__ = "0+reference.checkout"
        m.version = m.__version__
        sys.modules["tinygp.tinygp_version"] = m
    import jax

    assert getattr(jax, "__shim__", False), "a real jax is installed: use it instead of the shim"
    import tinygp

    return tinygp
