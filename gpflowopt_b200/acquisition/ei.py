"""ExpectedImprovement lives in leaves.py; this module keeps the import path of the reference (gpflowopt/acquisition/ei.py)."""
from .leaves import ExpectedImprovement  # noqa: F401
