"""ProbabilityOfImprovement lives in leaves.py; this module keeps the import path of the reference (gpflowopt/acquisition/poi.py)."""
from .leaves import ProbabilityOfImprovement  # noqa: F401
