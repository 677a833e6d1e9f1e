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


class _SitkImage:
    """What the reference script touches of a SimpleITK image (:60-63, :105-110), over io_nrrd.NrrdImage."""

    def __init__(self, nrrd_image):
        self._img = nrrd_image

    def GetSize(self):
        return self._img.GetSize()

    def GetSpacing(self):
        return self._img.GetSpacing()

    def GetOrigin(self):
        return self._img.GetOrigin()

    def GetDirection(self):
        return self._img.GetDirection()

    def SetSpacing(self, spacing):
        self._img.spacing = tuple(float(v) for v in spacing)

    def SetOrigin(self, origin):
        self._img.origin = tuple(float(v) for v in origin)

    def SetDirection(self, direction):
        import numpy as np
        self._img.direction = np.asarray(direction, dtype=np.float64).reshape(3, 3)


def _sitk_module():
    """A module object with the five SimpleITK calls of the reference script, backed by io_nrrd."""
    import types

    import numpy as np

    from . import io_nrrd
    m = types.ModuleType("SimpleITK")
    m.__doc__ = "physics_arx_b200 shim: ReadImage / GetArrayFromImage / GetImageFromArray / WriteImage over io_nrrd"
    m.Image = _SitkImage
    m.ReadImage = lambda path, *a, **k: _SitkImage(io_nrrd.read_nrrd(path))
    m.GetArrayFromImage = lambda img: np.array(img._img.array)
    m.GetImageFromArray = lambda arr, *a, **k: _SitkImage(io_nrrd.NrrdImage(np.asarray(arr)))
    m.WriteImage = lambda img, path, *a, **k: io_nrrd.write_nrrd(
        path, img._img.array, img._img.spacing, img._img.origin, img._img.direction, img._img.space)
    return m


def install(simpleitk="auto", tigre_build="2021"):
    """Register this package as ``tigre`` (+ submodules).  ``simpleitk``: 'auto' registers the io_nrrd-backed
    ``SimpleITK`` shim when the real package is not importable (it is not in this image), True always, False
    never -- so that the reference script (which does ``import SimpleITK as sitk``, :15) runs unmodified.
    ``tigre_build``: whose W / V weights ``algorithms.ossart`` applies when the caller names none (the script
    does not, :101-103) -- '2021': the Python package of mid-2021 the script was written against (W from the
    DSD-DSO cube with threshold min(dVoxel)/4, analytic V), 'upstream': current TIGRE's (MATLAB
    ``computeW``, V = Atb(ones): what a half-fan scan needs), None: leave the package's setting alone."""
    if tigre_build is not None:
        if tigre_build not in ("2021", "upstream"):
            raise ValueError(f"tigre_build must be '2021', 'upstream' or None, got {tigre_build!r}")
        from .algorithms.ossart import set_default_variants
        set_default_variants(*(("python2021", "B") if tigre_build == "2021" else ("matlab", "A")))
    me = sys.modules[__name__]
    sys.modules.setdefault("tigre", me)
    sys.modules.setdefault("tigre.algorithms", algorithms)
    sys.modules.setdefault("tigre.utilities", utilities)
    sys.modules.setdefault("tigre.utilities.CTnoise", CTnoise)
    sys.modules.setdefault("tigre.utilities.gpu", gpu)
    if simpleitk == "auto":
        import importlib.util
        simpleitk = "SimpleITK" not in sys.modules and importlib.util.find_spec("SimpleITK") is None
    if simpleitk:
        sys.modules["SimpleITK"] = _sitk_module()
    return me
