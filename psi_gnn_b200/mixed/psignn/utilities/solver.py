"""Drop-in for the reference's ``mixed/psignn/utilities/solver.py`` (same names and signatures)."""
from ....solver import broyden, anderson, forward_iteration, newton, LayerOperator, VjpOperator  # noqa: F401
