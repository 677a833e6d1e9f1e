"""``import physics_arx_b200.compat as tigre`` -- the TIGRE names the reference script uses.

``install()`` additionally registers the shim under ``sys.modules['tigre']`` (and its
``algorithms`` / ``utilities`` submodules) so that
Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py:10-15 imports
(``import tigre``, ``import tigre.algorithms as algs``, ``from tigre.utilities import
CTnoise``, ``import tigre.utilities.gpu as gpu``) resolve to this package unmodified.
"""
from __future__ import annotations

import sys

from . import algorithms, utilities  # noqa: F401
from .geometry import Geometry, geometry_default  # noqa: F401
from .projector import Atb, Ax  # noqa: F401
from .utilities import CTnoise, gpu  # noqa: F401

geometry = sys.modules[Geometry.__module__]


def install():
    me = sys.modules[__name__]
    sys.modules.setdefault("tigre", me)
    sys.modules.setdefault("tigre.algorithms", algorithms)
    sys.modules.setdefault("tigre.utilities", utilities)
    sys.modules.setdefault("tigre.utilities.CTnoise", CTnoise)
    sys.modules.setdefault("tigre.utilities.gpu", gpu)
    return me
