"""Quasiseparable solver and matrix algebra (same layout as src/tinygp/solvers/quasisep/: solver, core, ops, general, block)."""

from tinygp_b200.solvers.quasisep import block as block
from tinygp_b200.solvers.quasisep import core as core
from tinygp_b200.solvers.quasisep import general as general
from tinygp_b200.solvers.quasisep import ops as ops
from tinygp_b200.solvers.quasisep.solver import QuasisepSolver as QuasisepSolver
