"""B200-native sCBCT physics path (Ax, scatter/noise, Atb, OS-SART) for Physics-ArX.

Drop-in for the TIGRE call surface the reference uses
(Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py:67-103).
Importing the package is cheap; the CUDA library (csrc/libpax.so) is loaded on first
use and its absence is a hard error -- there is no CPU fallback.
"""
from .geometry import Geometry, geometry_default, order_subsets  # noqa: F401

__all__ = ["Geometry", "geometry_default", "Ax", "Atb", "algorithms", "utilities"]


def __getattr__(name):
    if name in ("Ax", "Atb", "Plan"):
        from . import projector
        return getattr(projector, name)
    if name in ("algorithms", "utilities", "phantom", "compat", "dist"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
