"""Solver plugins (same export names as src/tinygp/solvers/__init__.py:33-36)."""

from tinygp_b200.solvers.direct import DirectSolver as DirectSolver
from tinygp_b200.solvers.quasisep import QuasisepSolver as QuasisepSolver
from tinygp_b200.solvers.solver import Solver as Solver

B200DirectSolver = DirectSolver
B200QuasisepSolver = QuasisepSolver
