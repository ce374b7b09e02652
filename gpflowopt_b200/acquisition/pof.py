"""ProbabilityOfFeasibility lives in leaves.py; this module keeps the import path of the reference (gpflowopt/acquisition/pof.py)."""
from .leaves import ProbabilityOfFeasibility  # noqa: F401
