"""Seeded synthetic planning-CT phantoms and the benchmark geometries (host side, NumPy).

SURVEY.md 8(d): lung-like phantom in [0,1] (body ellipsoid 0.25, two lungs 0.05,
spine cylinder 0.45, heart 0.27, 10 random spheres) and an "artifact" volume
(low-pass zero-mean Gaussian field x 0.05) standing in for the PL-AHE output the
reference adds to the planning CT (CBCT_pCT_Artifact_adding.sh:92-111).  The
reference's own sample patient is absent (.MISSING_LARGE_BLOBS:18-57), hence
synthetic data.
"""
from __future__ import annotations

import numpy as np

from .geometry import Geometry


def _grid(nvox, dvox):
    nz, ny, nx = (int(v) for v in nvox)
    z = (np.arange(nz) - nz / 2 + 0.5) * dvox[0]
    y = (np.arange(ny) - ny / 2 + 0.5) * dvox[1]
    x = (np.arange(nx) - nx / 2 + 0.5) * dvox[2]
    return z[:, None, None], y[None, :, None], x[None, None, :]


def lung_phantom(nvox, dvox, seed=0, body_mm=None):
    """(Z,Y,X) float32 phantom in [0,1]; anatomy scaled to the in-plane extent."""
    nvox = np.asarray(nvox)
    dvox = np.asarray(dvox, dtype=np.float64)
    z, y, x = _grid(nvox, dvox)
    sy, sx = nvox[1] * dvox[1], nvox[2] * dvox[2]
    sz = nvox[0] * dvox[0]
    if body_mm is None:
        body_mm = (0.46 * sz, 0.33 * sy, 0.44 * sx)   # semi-axes z,y,x
    bz, by, bx = body_mm
    vol = np.zeros(tuple(int(v) for v in nvox), dtype=np.float32)

    def ell(cz, cy, cx, rz, ry, rx):
        return ((z - cz) / rz) ** 2 + ((y - cy) / ry) ** 2 + ((x - cx) / rx) ** 2 <= 1.0

    vol[ell(0, 0, 0, bz, by, bx)] = 0.25
    vol[ell(0, -0.05 * by, -0.45 * bx, 0.75 * bz, 0.6 * by, 0.33 * bx)] = 0.05
    vol[ell(0, -0.05 * by, 0.45 * bx, 0.75 * bz, 0.6 * by, 0.33 * bx)] = 0.05
    vol[ell(-0.1 * bz, 0.1 * by, 0.05 * bx, 0.3 * bz, 0.3 * by, 0.2 * bx)] = 0.27
    spine = (((y - 0.62 * by) / (0.16 * by)) ** 2 + (x / (0.1 * bx)) ** 2 <= 1.0) & (np.abs(z) <= bz)
    vol[np.broadcast_to(spine, vol.shape)] = 0.45
    rng = np.random.default_rng(seed)
    for _ in range(10):
        c = rng.uniform(-0.6, 0.6, size=3) * np.array([bz, by, bx])
        r = rng.uniform(0.03, 0.08) * min(by, bx)
        val = rng.uniform(0.1, 0.4)
        m = ell(c[0], c[1], c[2], r, r, r) & ell(0, 0, 0, bz, by, bx)
        vol[m] = np.float32(val)
    return vol


def artifact_volume(nvox, seed=1, amplitude=0.05, smooth=3):
    """Low-pass zero-mean Gaussian field (stand-in for the PL-AHE artifact volume)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(tuple(int(v) for v in nvox))
    for ax in range(3):
        a = _box(a, ax, smooth)
    a -= a.mean()
    a *= amplitude / (np.abs(a).max() + 1e-12)
    return a.astype(np.float32)


def _box(a, axis, r):
    c = np.cumsum(np.concatenate([np.zeros_like(np.take(a, [0], axis=axis)), a], axis=axis), axis=axis)
    n = a.shape[axis]
    idx_hi = np.clip(np.arange(n) + r + 1, 0, n)
    idx_lo = np.clip(np.arange(n) - r, 0, n)
    return (np.take(c, idx_hi, axis=axis) - np.take(c, idx_lo, axis=axis)) / (2 * r + 1)


def make_geometry(nvox, dvox, ndet, ddet, DSD=1500.0, DSO=1000.0, off_detector=(0.0, 0.0)):
    geo = Geometry()
    geo.DSD, geo.DSO = float(DSD), float(DSO)
    geo.nVoxel = np.array(nvox, dtype=np.int64)
    geo.dVoxel = np.array(dvox, dtype=np.float64)
    geo.sVoxel = geo.nVoxel * geo.dVoxel
    geo.nDetector = np.array(ndet, dtype=np.int64)
    geo.dDetector = np.array(ddet, dtype=np.float64)
    geo.sDetector = geo.nDetector * geo.dDetector
    geo.offDetector = np.array(off_detector, dtype=np.float64)
    return geo


# ---- the named benchmark configurations ----
def config_c1():
    """C1: 128^3 @2 mm, 256 angles, 256x256 detector @2 mm, 5 iterations."""
    geo = make_geometry((128, 128, 128), (2.0, 2.0, 2.0), (256, 256), (2.0, 2.0))
    return dict(name="C1", geo=geo, angles=np.linspace(0, 2 * np.pi, 256), niter=5, blocksize=20,
                I0=1e8, sigma=4.0)


def config_c2():
    """C2: 512x512x128 planning CT, 656 half-fan projections of 1024x768, 20 iterations."""
    d_inplane = 500.0 / 512.0
    geo = make_geometry((128, 512, 512), (2.5, d_inplane, d_inplane), (768, 1024), (0.388, 0.388),
                        off_detector=(0.0, -160.0))
    return dict(name="C2", geo=geo, angles=np.linspace(0, 2 * np.pi, 656), niter=20, blocksize=20,
                I0=1e8, sigma=4.0)


def config_tiny(n=24, nang=12, ndet=32):
    """A seconds-scale case for oracle parity tests."""
    geo = make_geometry((n, n, n), (256.0 / n,) * 3, (ndet, ndet), (512.0 / ndet,) * 2)
    return dict(name="tiny", geo=geo, angles=np.linspace(0, 2 * np.pi, nang, endpoint=False),
                niter=2, blocksize=5, I0=1e8, sigma=4.0)
