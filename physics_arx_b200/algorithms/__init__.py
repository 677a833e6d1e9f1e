"""``tigre.algorithms`` surface used by the reference: ``ossart`` (pCT_CBCT_Reconstruction.py:103)."""
from .ossart import OS_SART, ossart  # noqa: F401

__all__ = ["ossart", "OS_SART"]
