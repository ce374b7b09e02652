"""LowerConfidenceBound lives in leaves.py; this module keeps the import path of the reference (gpflowopt/acquisition/lcb.py)."""
from .leaves import LowerConfidenceBound  # noqa: F401
