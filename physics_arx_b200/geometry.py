"""Cone-beam geometry bag with the names and ordering the reference uses.

The reference builds its geometry with ``tigre.geometry_default`` and then
overrides attributes by hand
(Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py:67-83):
``DSD, nDetector, sDetector, dDetector, nVoxel, sVoxel, dVoxel`` and, commented
out, ``offDetector`` (:83).  This module keeps that attribute surface
(SURVEY.md A.1): volumes are (Z,Y,X), detectors are (V,U), units mm / radians.

Only what the path needs is supported: circular cone-beam trajectories about
z with per-angle DSD / DSO / offOrigin / offDetector.  Detector rotation, COR
and the second/third Euler angles must be zero (the reference leaves them at
their zero defaults); anything else raises ``NotImplementedError``.
"""
from __future__ import annotations

import copy

import numpy as np

# layout of the flat blocks handed to the C-ABI (include/pax.h) and the oracle
ANGLE_STRIDE = 8  # [theta, DSD, DSO, offOrigZ, offOrigY, offOrigX, offDetV, offDetU]


class Geometry:
    """Attribute bag, TIGRE-style.  All array attributes are numpy arrays."""

    def __init__(self):
        self.mode = "cone"
        self.DSD = 1536.0
        self.DSO = 1000.0
        self.nDetector = np.array((128, 128))
        self.dDetector = np.array((3.2, 3.2))
        self.sDetector = self.nDetector * self.dDetector
        self.nVoxel = np.array((64, 64, 64))
        self.sVoxel = np.array((256.0, 256.0, 256.0))
        self.dVoxel = self.sVoxel / self.nVoxel
        self.offOrigin = np.array((0.0, 0.0, 0.0))
        self.offDetector = np.array((0.0, 0.0))
        self.rotDetector = np.array((0.0, 0.0, 0.0))
        self.COR = 0.0
        self.accuracy = 0.5

    def __repr__(self):
        keys = ("mode", "DSD", "DSO", "nDetector", "dDetector", "sDetector", "nVoxel",
                "sVoxel", "dVoxel", "offOrigin", "offDetector", "rotDetector", "COR", "accuracy")
        return "Geometry(\n" + "\n".join(f"  {k} = {getattr(self, k)!r}" for k in keys) + "\n)"

    def copy(self) -> "Geometry":
        return copy.deepcopy(self)

    # ------------------------------------------------------------------ checks
    def check_geo(self, angles) -> np.ndarray:
        """Validate and return angles as an (N,3) float64 array [theta,0,0].

        Mirrors the role of TIGRE's ``geo.check_geo(angles)``:
        shapes are checked, scalars are broadcast per angle by ``pack``.
        """
        if self.mode != "cone":
            raise NotImplementedError("only mode='cone' is on this path")
        ang = np.asarray(angles, dtype=np.float64)
        if ang.ndim == 1:
            ang = np.stack([ang, np.zeros_like(ang), np.zeros_like(ang)], axis=1)
        if ang.ndim != 2 or ang.shape[1] != 3:
            raise ValueError(f"angles must be (N,) or (N,3), got {ang.shape}")
        if np.any(ang[:, 1:] != 0):
            raise NotImplementedError("only rotation about z (angles[:,1:]==0) is on this path")
        for name, n in (("nVoxel", 3), ("sVoxel", 3), ("dVoxel", 3), ("nDetector", 2),
                        ("sDetector", 2), ("dDetector", 2)):
            arr = np.asarray(getattr(self, name))
            if arr.shape != (n,):
                raise ValueError(f"geo.{name} must have shape ({n},), got {arr.shape}")
        nvox = np.asarray(self.nVoxel)
        ndet = np.asarray(self.nDetector)
        if np.any(nvox < 1) or np.any(ndet < 1):
            raise ValueError("geo.nVoxel and geo.nDetector must be positive")
        if not np.allclose(np.asarray(self.dVoxel, float) * nvox, np.asarray(self.sVoxel, float),
                           rtol=1e-5, atol=1e-6):
            raise ValueError("geo.sVoxel != geo.nVoxel * geo.dVoxel")
        if not np.allclose(np.asarray(self.dDetector, float) * ndet,
                           np.asarray(self.sDetector, float), rtol=1e-5, atol=1e-6):
            raise ValueError("geo.sDetector != geo.nDetector * geo.dDetector")
        if np.any(np.asarray(self.rotDetector, float) != 0):
            raise NotImplementedError("rotDetector != 0 is not on this path")
        if np.any(np.asarray(self.COR, float) != 0):
            raise NotImplementedError("COR != 0 is not on this path")
        if not (float(self.accuracy) > 0):
            raise ValueError("geo.accuracy must be > 0")
        return ang

    # ------------------------------------------------------------- flat blocks
    def _per_angle(self, value, n, width, name):
        arr = np.asarray(value, dtype=np.float64)
        if width == 1:
            if arr.ndim == 0:
                return np.full((n,), float(arr))
            if arr.shape == (n,):
                return arr.copy()
        else:
            if arr.shape == (width,):
                return np.broadcast_to(arr, (n, width)).copy()
            if arr.shape == (n, width):
                return arr.copy()
        raise ValueError(f"geo.{name} has shape {arr.shape}; expected scalar/({width},) or per-angle")

    def pack(self, angles):
        """Return (geo_block[11] f64, angle_rows[N,8] f64) for the C-ABI / oracle."""
        ang = self.check_geo(angles)
        n = ang.shape[0]
        nv = np.asarray(self.nVoxel)
        nd = np.asarray(self.nDetector)
        dv = np.asarray(self.dVoxel, dtype=np.float64)
        dd = np.asarray(self.dDetector, dtype=np.float64)
        g = np.array([nv[0], nv[1], nv[2], nd[0], nd[1], dv[0], dv[1], dv[2], dd[0], dd[1],
                      float(self.accuracy)], dtype=np.float64)
        rows = np.empty((n, ANGLE_STRIDE), dtype=np.float64)
        rows[:, 0] = ang[:, 0]
        rows[:, 1] = self._per_angle(self.DSD, n, 1, "DSD")
        rows[:, 2] = self._per_angle(self.DSO, n, 1, "DSO")
        rows[:, 3:6] = self._per_angle(self.offOrigin, n, 3, "offOrigin")
        rows[:, 6:8] = self._per_angle(self.offDetector, n, 2, "offDetector")
        if np.any(rows[:, 2] <= 0) or np.any(rows[:, 1] <= rows[:, 2]):
            raise ValueError("need DSD > DSO > 0")
        return g, rows

    # ------------------------------------------------------ derived quantities
    def w_box(self, variant="matlab"):
        """Half-sizes (z,y,x) in mm and threshold of the box whose chord length is 1/W.

        SURVEY.md A.5.  'matlab' (default; MATLAB computeW and the appendix's main
        statement): in-plane sVoxel*1.1, z = max(sDetector_V, sVoxel_z), threshold
        min(dVoxel)/2.  'python2021': every sVoxel entry := DSD-DSO, then the
        array's last entry := max(sDetector_U, that), threshold min(dVoxel)/4.
        """
        sv = np.asarray(self.sVoxel, dtype=np.float64)
        sd = np.asarray(self.sDetector, dtype=np.float64)
        dmin = float(np.min(np.asarray(self.dVoxel, dtype=np.float64)))
        if variant == "matlab":
            s = np.array([max(sd[0], sv[0]), sv[1] * 1.1, sv[2] * 1.1])
            thr = dmin / 2.0
        elif variant == "python2021":
            r = float(np.max(np.asarray(self.DSD, float)) - np.max(np.asarray(self.DSO, float)))
            s = np.array([r, r, max(sd[1], r)])
            thr = dmin / 4.0
        else:
            raise ValueError(f"unknown W variant {variant!r}")
        return s / 2.0, thr

    def v_scale(self):
        """Shrink factor of set_v's geometry (A.5 variant A): 0.9*max(sY,sX)/||(sY,sX)||."""
        sv = np.asarray(self.sVoxel, dtype=np.float64)
        return float(0.9 * np.max(sv[1:]) / np.linalg.norm(sv[1:]))


def geometry_default(high_resolution=False, nVoxel=None):
    """``tigre.geometry_default`` (reference call site pCT_CBCT_Reconstruction.py:67).

    SURVEY.md A.1: DSD 1536, DSO 1000, accuracy 0.5, mode 'cone', offsets 0;
    low resolution: 128x128 detector @3.2 mm, 64^3 voxels over 256 mm;
    high resolution (not used by the reference; recalled, unverified):
    512x512 @0.8 mm, 512^3 voxels.
    """
    geo = Geometry()
    if nVoxel is not None:
        geo.nDetector = np.array((int(nVoxel[0]), int(nVoxel[1])))
        geo.dDetector = np.array((0.8, 0.8)) * 512.0 / geo.nDetector
        geo.nVoxel = np.array(nVoxel, dtype=np.int64)
    elif high_resolution:
        geo.nDetector = np.array((512, 512))
        geo.dDetector = np.array((0.8, 0.8))
        geo.nVoxel = np.array((512, 512, 512))
    geo.sDetector = geo.nDetector * geo.dDetector
    geo.sVoxel = np.array((256.0, 256.0, 256.0))
    geo.dVoxel = geo.sVoxel / geo.nVoxel
    return geo


def order_subsets(n_angles, blocksize, strategy="ordered", seed=0):
    """Index blocks for OS-SART (A.6: ``OrderStrategy=None`` -> contiguous ordered blocks)."""
    if blocksize is None or blocksize < 1:
        raise ValueError("blocksize must be >= 1")
    idx = np.arange(n_angles)
    blocks = [idx[i:i + blocksize] for i in range(0, n_angles, blocksize)]
    if strategy in (None, "ordered"):
        return blocks
    if strategy == "random":
        rng = np.random.default_rng(seed)
        order = rng.permutation(len(blocks))
        return [blocks[i] for i in order]
    raise ValueError(f"unknown OrderStrategy {strategy!r}")
