#!/bin/bash
# Validate the jax/equinox stand-ins by running the reference's OWN test-suite over them (read-only checkout).
# The jax.grad tests and the george/celerite comparisons cannot run (no autodiff in the shim, packages absent).
cd "$(dirname "$0")"
PYTHONPATH="$PWD:$PYTHONPATH" PYTHONDONTWRITEBYTECODE=1 python -m pytest -p refplugin -p no:cacheprovider \
    --rootdir /tmp -q "${@:-/root/reference/tests}"
