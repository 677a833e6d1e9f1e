"""Python face of the float64 CPU oracle (oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(physics_arx_b200) never imports this module.

PARITY UNPINNED (see oracle.c header): the reference holds no golden vectors for
this path and its arithmetic lives in un-vendored TIGRE.  The call sequence and
constants follow
Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py:67-103;
the projector / backprojector / weight / noise maths follow SURVEY.md Appendix A.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_ax_interp.argtypes = [_dp, _dp, ctypes.c_long, _dp, _dp, _dp]
        lib.oracle_atb_fdk.argtypes = [_dp, _dp, ctypes.c_long, _dp, _dp, ctypes.c_double]
        lib.oracle_raylen_w.argtypes = [_dp, _dp, ctypes.c_long, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_double, _dp]
        lib.oracle_philox.argtypes = [ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.POINTER(ctypes.c_uint32)]
        lib.oracle_philox_raw.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        lib.oracle_philox_raw.restype = None
        lib.oracle_ctnoise.argtypes = [_dp, ctypes.c_long, ctypes.c_uint64, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_uint64, _dp]
        lib.oracle_sart_update.argtypes = [_dp, _dp, _dp, ctypes.c_long, ctypes.c_long,
                                           ctypes.c_double, ctypes.c_int]
        lib.oracle_resample_linear.argtypes = [_dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, _dp, _dp,
                                               ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                               ctypes.c_double, ctypes.c_int]
        lib.oracle_rescale_intensity.argtypes = [_dp, ctypes.c_long, ctypes.c_double, ctypes.c_double, _dp]
        lib.oracle_plahe.argtypes = [_dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                                     ctypes.c_double, ctypes.c_double, _dp]
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        lib.oracle_set_threads.restype = ctypes.c_int
        for f in (lib.oracle_ax_interp, lib.oracle_atb_fdk, lib.oracle_raylen_w,
                  lib.oracle_philox, lib.oracle_ctnoise, lib.oracle_sart_update,
                  lib.oracle_resample_linear, lib.oracle_rescale_intensity, lib.oracle_plahe):
            f.restype = None
        _lib = lib
    return _lib


def set_threads(n=None):
    """Run the oracle's OpenMP loops on n threads (default: every host core); returns the count in effect."""
    return int(_load().oracle_set_threads(int(n if n else (os.cpu_count() or 1))))


def _p(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ax(vol, geo, angles, return_counts=False):
    """Ax 'interpolated' -> (N,V,U) float64.  SURVEY.md A.3; reference :86."""
    g, rows = geo.pack(angles)
    vol = _c64(vol)
    if vol.shape != tuple(int(v) for v in geo.nVoxel):
        raise ValueError("volume shape != geo.nVoxel")
    n = rows.shape[0]
    out = np.empty((n, int(g[3]), int(g[4])), dtype=np.float64)
    cnt = np.zeros((n,), dtype=np.float64)
    _load().oracle_ax_interp(_p(g), _p(rows), n, _p(vol), _p(out), _p(cnt))
    return (out, cnt) if return_counts else out


def atb(proj, geo, angles, vscale=1.0):
    """Atb 'FDK' -> (Z,Y,X) float64.  SURVEY.md A.4.  proj=None -> all-ones projections."""
    g, rows = geo.pack(angles)
    n = rows.shape[0]
    out = np.empty(tuple(int(v) for v in geo.nVoxel), dtype=np.float64)
    if proj is None:
        pp = ctypes.cast(None, _dp)
    else:
        proj = _c64(proj)
        if proj.shape != (n, int(g[3]), int(g[4])):
            raise ValueError("projection shape != (N, *geo.nDetector)")
        pp = _p(proj)
    _load().oracle_atb_fdk(_p(g), _p(rows), n, pp, _p(out), float(vscale))
    return out


def compute_w(geo, angles, variant="matlab"):
    """W = 1 / chord length through the enlarged box, 0 where the ray misses (A.5)."""
    g, rows = geo.pack(angles)
    n = rows.shape[0]
    (hz, hy, hx), thr = geo.w_box(variant)
    out = np.empty((n, int(g[3]), int(g[4])), dtype=np.float64)
    _load().oracle_raylen_w(_p(g), _p(rows), n, hz, hy, hx, thr, _p(out))
    return out


def compute_v(geo, angles, blocks, variant="A"):
    """V per subset -> (nSubsets, Y, X) (A.5)."""
    g, rows = geo.pack(angles)
    ny, nx = int(g[1]), int(g[2])
    out = np.empty((len(blocks), ny, nx), dtype=np.float64)
    if variant == "A":
        scale = geo.v_scale()
        for s, idx in enumerate(blocks):
            sub = np.asarray(angles)[idx]
            gs = _subset_geo(geo, idx, len(np.asarray(angles)))
            out[s] = atb(None, gs, sub, vscale=scale).mean(axis=0)
    elif variant == "B":
        # 2021 Python: V(y,x,theta) = (DSO/(DSO + y sin(-A) - x cos(-A)))^2, A = theta + pi/2
        dv = np.asarray(geo.dVoxel, float)
        sv = np.asarray(geo.sVoxel, float)
        x = -sv[2] / 2 + dv[2] / 2 + dv[2] * np.arange(nx) + rows[0, 5]
        y = -sv[1] / 2 + dv[1] / 2 + dv[1] * np.arange(ny) + rows[0, 4]
        yy, xx = np.meshgrid(y, x, indexing="ij")
        for s, idx in enumerate(blocks):
            acc = np.zeros((ny, nx))
            for i in idx:
                A = rows[i, 0] + np.pi / 2
                dso = rows[i, 2]
                acc += (dso / (dso + yy * np.sin(-A) - xx * np.cos(-A))) ** 2
            out[s] = acc
    else:
        raise ValueError(variant)
    return out


def _subset_geo(geo, idx, n):
    """Geometry restricted to a subset of angles (per-angle attributes sliced)."""
    gs = geo.copy()
    for name, width in (("DSD", 1), ("DSO", 1), ("offOrigin", 3), ("offDetector", 2)):
        arr = np.asarray(getattr(geo, name), dtype=np.float64)
        if (width == 1 and arr.shape == (n,)) or (width > 1 and arr.shape == (n, width)):
            setattr(gs, name, arr[idx])
    return gs


def philox(seed, index):
    out = (ctypes.c_uint32 * 4)()
    _load().oracle_philox(int(seed), int(index), out)
    return np.array(list(out), dtype=np.uint32)


def philox_raw(ctr, key):
    c = (ctypes.c_uint32 * 4)(*[int(v) for v in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(v) for v in key])
    out = (ctypes.c_uint32 * 4)()
    _load().oracle_philox_raw(c, k, out)
    return [int(v) for v in out]


def ctnoise_add(proj, Poisson=1e5, Gaussian=np.array([0.0, 0.5]), seed=0, index0=0, maxval=None):
    """CTnoise.add (A.7; reference :88) with this build's Philox stream."""
    if not np.isscalar(Poisson) and np.ndim(Poisson) != 0:
        raise ValueError("Poisson must be a scalar")
    Gaussian = np.asarray(Gaussian, dtype=np.float64)
    if Gaussian.shape != (2,):
        raise ValueError("Gaussian must have shape (2,)")
    p = _c64(proj)
    m = float(p.max()) if maxval is None else float(maxval)
    out = np.empty_like(p)
    _load().oracle_ctnoise(_p(p), p.size, int(index0), float(Poisson), float(Gaussian[0]),
                           float(Gaussian[1]), m, int(seed), _p(out))
    return out


def ossart(proj, geo, angles, niter, blocksize=20, lmbda=1.0, lmbda_red=0.99, noneg=True,
           w_variant="matlab", v_variant="A", order="ordered", init=None, return_l2=False):
    """OS-SART (A.6; reference call site pCT_CBCT_Reconstruction.py:101-103).

    x <- max(0, x + lambda * V_s^-1 * Atb_FDK(W_s * (b_s - Ax(x)))), lambda *= lmbda_red per
    outer iteration, x0 = 0, contiguous ordered subsets.  Everything float64.
    """
    from physics_arx_b200.geometry import order_subsets  # pure-python host logic, no CUDA

    angles = np.asarray(angles, dtype=np.float64)
    if angles.ndim == 2:
        angles = angles[:, 0]
    n = angles.shape[0]
    proj = _c64(proj)
    blocks = order_subsets(n, blocksize, order)
    W = compute_w(geo, angles, w_variant)
    V = compute_v(geo, angles, blocks, v_variant)
    with np.errstate(divide="ignore"):
        invV = np.where(V > 1e-12, 1.0 / V, 0.0)
    nz, ny, nx = (int(v) for v in geo.nVoxel)
    x = np.zeros((nz, ny, nx)) if init is None else _c64(init).copy()
    lam = float(lmbda)
    l2 = []
    lib = _load()
    for _ in range(int(niter)):
        err = 0.0
        for s, idx in enumerate(blocks):
            gs = _subset_geo(geo, idx, n)
            r = proj[idx] - ax(x, gs, angles[idx])
            err += float((r * r).sum())
            r *= W[idx]
            upd = atb(r, gs, angles[idx])
            lib.oracle_sart_update(_p(x), _p(upd), _p(np.ascontiguousarray(invV[s])), nz, ny * nx,
                                   lam, int(bool(noneg)))
        l2.append(np.sqrt(err))
        lam *= lmbda_red
    return (x, np.array(l2)) if return_l2 else x


# ------------------------------------------------------------------ Artifact-Induction tools
def index_map(dst_origin, dst_spacing, dst_direction, src_origin, src_spacing, src_direction):
    """12 doubles: continuous SOURCE index (x,y,z) of DESTINATION index (i,j,k), identity transform
    (what itk::ResampleImageFilter evaluates; origins/spacings in (x,y,z) order, directions 3x3)."""
    dd = np.asarray(dst_direction, dtype=np.float64).reshape(3, 3)
    sd = np.asarray(src_direction, dtype=np.float64).reshape(3, 3)
    a = np.linalg.inv(sd @ np.diag(np.asarray(src_spacing, dtype=np.float64)))
    lin = a @ dd @ np.diag(np.asarray(dst_spacing, dtype=np.float64))
    off = a @ (np.asarray(dst_origin, dtype=np.float64) - np.asarray(src_origin, dtype=np.float64))
    return np.ascontiguousarray(np.concatenate([lin, off[:, None]], axis=1).reshape(-1))


def resample_linear(src, m, dst_shape, default=0.0, cast="float"):
    """ResampleImages / ResampleImagesFloat / AddImages' resampling (ITK linear, identity transform)."""
    src = _c64(src)
    m = _c64(m)
    out = np.empty(tuple(int(v) for v in dst_shape), dtype=np.float64)
    _load().oracle_resample_linear(_p(src), *src.shape, _p(m), _p(out), *out.shape, float(default),
                                   1 if cast == "short" else 0)
    return out


def rescale_intensity(img, out_min=0.0, out_max=1.0):
    img = _c64(img)
    out = np.empty_like(img)
    _load().oracle_rescale_intensity(_p(img), img.size, float(out_min), float(out_max), _p(out))
    return out


def plahe(img, radius=5, alpha=0.5, beta=0.5):
    """itk::AdaptiveHistogramEqualizationImageFilter (AdaptiveHistogramMatch.cxx:30-42)."""
    img = _c64(img)
    out = np.empty_like(img)
    _load().oracle_plahe(_p(img), *img.shape, int(radius), float(alpha), float(beta), _p(out))
    return out
