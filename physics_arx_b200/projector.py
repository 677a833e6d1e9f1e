"""``Ax`` / ``Atb`` with the reference's call surface, and the resident-volume ``Plan``.

The reference calls ``tigre.Ax(img, geo, angles, 'interpolated', gpuids=gpuids)``
(pCT_CBCT_Reconstruction.py:86) and reaches ``tigre.Atb(..., 'FDK')`` through
``algs.ossart`` (:103).  Same names, argument order and error behaviour here
(SURVEY.md 8(b)): float32 only (``TypeError``), shapes must match ``geo``
(``ValueError``); NumPy in -> NumPy out, torch CUDA tensor in -> torch CUDA tensor out
(zero-copy).  PyTorch is used for device memory and streams only; all arithmetic is
in libpax.so (csrc/), called through the C ABI of include/pax.h.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from .geometry import Geometry


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _device_from_gpuids(gpuids):
    """First device a ``gpuids`` argument names (calls that list several go through multidev.MultiDevice)."""
    if gpuids is None:
        return torch.cuda.current_device()
    devices = getattr(gpuids, "devices", gpuids)
    if isinstance(devices, (list, tuple, range)):
        if len(devices) == 0:
            raise ValueError("gpuids is empty")
        return int(devices[0])
    return int(devices)


def _multi(gpuids):
    """The device list if ``gpuids`` names more than one device (TIGRE splits such a call over all of them,
    reference :31,86,103), else None."""
    if gpuids is None:
        return None
    from .multidev import devices_of
    devs = devices_of(gpuids)
    return devs if len(devs) > 1 else None


class Plan:
    """Per-(geometry, angle list) device constants plus helpers on the resident layout.

    The resident layout of a volume is libpax's z-fastest zero-rimmed array
    (include/pax.h: pax_vol_pack); OS-SART keeps its iterate in that layout for the
    whole run so that Ax and Atb never convert.
    """

    def __init__(self, geo: Geometry, angles, device=None, projector=None):
        if not torch.cuda.is_available():
            raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        g, rows = geo.pack(angles)
        self.geo = geo
        self.w_geo = geo               # geometry the W box is derived from (the FULL detector when
                                       # this plan describes a row band of it)
        self.n_angles = rows.shape[0]
        self.nz, self.ny, self.nx = int(g[0]), int(g[1]), int(g[2])
        self.nv, self.nu = int(g[3]), int(g[4])
        cg = _lib.PaxGeo(self.nz, self.ny, self.nx, self.nv, self.nu, g[5], g[6], g[7], g[8], g[9], g[10])
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_plan_create(ctypes.byref(cg), rows.ctypes.data_as(_lib.c_double_p),
                                                self.n_angles, ctypes.byref(handle)), "pax_plan_create")
        self._h = handle
        self.vol_elems = int(self.lib.pax_vol_elems(self._h))
        self.nzp = ((self.nz + 5) + 3) // 4 * 4          # resident layout (include/pax.h): (Y+2, X+2, Zp), z fastest
        self.nxp_nzp_bytes = (self.nx + 2) * self.nzp * 4  # one y-row of it
        assert (self.ny + 2) * (self.nx + 2) * self.nzp == self.vol_elems
        self._work = None              # Ax scratch (quad-packed volume), grown on demand
        # PAX_AX_KERNEL / projector=: auto (default) | fan | ray (= quad) | plain  (include/pax.h)
        kind = (projector or os.environ.get("PAX_AX_KERNEL", "auto")).lower()
        if kind not in ("auto", "fan", "ray", "quad", "plain"):
            raise ValueError(f"unknown projector {kind!r}")
        self.use_workspace = kind != "plain"
        self.set_projector({"auto": "auto", "fan": "fan"}.get(kind, "ray"))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self.lib.pax_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_projector(self, kind):
        """'auto' | 'ray' | 'fan': which forward-projector kernel pax_ax runs (include/pax.h)."""
        code = {"auto": _lib.PAX_PROJECTOR_AUTO, "ray": _lib.PAX_PROJECTOR_RAY,
                "fan": _lib.PAX_PROJECTOR_FAN}[kind]
        _lib.check(self.lib.pax_plan_set_projector(self._h, code), "pax_plan_set_projector")

    @property
    def projector(self):
        return {_lib.PAX_PROJECTOR_RAY: "ray", _lib.PAX_PROJECTOR_FAN: "fan"}[
            int(self.lib.pax_plan_projector(self._h))]

    def set_column_share(self, part, parts):
        """Multi-GPU: the fan-plane kernel of this plan walks share `part` of `parts` of the detector's column
        tiles only (include/pax.h: pax_plan_set_column_share)."""
        _lib.check(self.lib.pax_plan_set_column_share(self._h, int(part), int(parts)), "pax_plan_set_column_share")

    def fan_info(self):
        rows, warps, run = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        _lib.check(self.lib.pax_plan_fan_info(self._h, ctypes.byref(rows), ctypes.byref(warps),
                                              ctypes.byref(run)), "pax_plan_fan_info")
        return {"rows_per_tile": rows.value, "warps": warps.value, "mean_run": run.value}

    def fault(self):
        """1 if the fan kernel ever met a z window wider than its lanes on this plan (self-check)."""
        return int(self.lib.pax_plan_fault(self._h))

    def check_fault(self):
        """Raise if the fan-plane projector flagged a launch whose z range did not fit its lanes (the results
        of such a launch are wrong; pax_fan_configure is meant to rule it out).  Synchronises the device."""
        if self.projector == "fan" and self.fault():
            raise _lib.PaxError("the fan-plane projector met a detector row whose z range does not fit its lanes "
                                "on this geometry; use projector='ray' (PAX_AX_KERNEL=ray)")

    def ax_launches(self, channels=1):
        """kernels one pax_ax call launches (bench.py's gpu_launches claim)."""
        if self.projector == "fan":
            return channels            # (+ 1 finish kernel per channel when a small launch is split)
        return 1 + (channels if self._workspace(channels) is not None else 0)

    # ------------------------------------------------------------------ volumes
    def new_volume(self):
        return torch.zeros(self.vol_elems, dtype=torch.float32, device=self.device)

    def _check_vol(self, vol):
        if vol.dtype != torch.float32:
            raise TypeError("volume must be float32")
        if tuple(vol.shape) != (self.nz, self.ny, self.nx):
            raise ValueError(f"volume shape {tuple(vol.shape)} != geo.nVoxel {(self.nz, self.ny, self.nx)}")

    def pack(self, vol, out=None):
        """(Z,Y,X) float32 CUDA tensor -> resident layout."""
        self._check_vol(vol)
        vol = vol.contiguous()
        out = torch.empty(self.vol_elems, dtype=torch.float32, device=self.device) if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_vol_pack(self._h, _ptr(vol), _ptr(out), _stream()), "pax_vol_pack")
        return out

    def unpack(self, res, out=None):
        out = torch.empty((self.nz, self.ny, self.nx), dtype=torch.float32, device=self.device) \
            if out is None else out
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_vol_unpack(self._h, _ptr(res), _ptr(out), _stream()), "pax_vol_unpack")
        return out

    def _workspace(self, channels, count=None):
        """Scratch pax_ax wants for this call: the ray kernel's re-packed volume(s), or -- fan kernel, small
        launches only -- the partial sums of a launch split along the chords (include/pax.h)."""
        if not self.use_workspace:
            return None
        need = int(self.lib.pax_ax_workspace_bytes(self._h, channels))
        if need == 0 and count is not None:
            need = int(self.lib.pax_ax_fan_scratch_bytes(self._h, int(count)))
        if need == 0:                  # geometry outside the workspace kernel's range / nothing wanted
            return None
        if self._work is None or self._work.numel() * 4 < need:
            self._work = torch.empty((need + 3) // 4, dtype=torch.float32, device=self.device)
        return self._work

    # --------------------------------------------------------------- projectors
    def _w_box(self, w_variant):
        (hz, hy, hx), thr = self.w_geo.w_box(w_variant)
        return float(hz), float(hy), float(hx), float(thr)

    def axpy(self, y_res, x_res, a):
        """y_res += a * x_res (resident layout): image-domain scatter transfer, AddImages.cxx:57-62."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_vol_axpy(self._h, _ptr(y_res), _ptr(x_res), float(a), _stream()),
                       "pax_vol_axpy")
        return y_res

    def ax(self, res0, first=0, count=None, out=None, res1=None, a0=1.0, a1=0.0, c=0.0,
           w_variant="matlab", nsamples=None, maxout=None, out1=None):
        """out[(count,V,U)] = a0*Ax(res0) + a1*Ax(res1) + c*chord.  (PAX_AX_PLAIN)

        With ``out1`` (and ``res1``) the second channel's line integrals go to ``out1`` instead of
        being mixed in: both projections from one pass."""
        count = self.n_angles - first if count is None else count
        if out is None:
            out = torch.empty((count, self.nv, self.nu), dtype=torch.float32, device=self.device)
        hz, hy, hx, thr = self._w_box(w_variant)
        epi = _lib.PaxAxEpilogue(_lib.PAX_AX_PLAIN, a0, a1, c, None, hz, hy, hx, thr, None,
                                 nsamples.data_ptr() if nsamples is not None else None,
                                 out1.data_ptr() if out1 is not None else None,
                                 maxout.data_ptr() if maxout is not None else None)
        if maxout is not None:
            maxout.zero_()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_ax(self._h, _ptr(res0), _ptr(res1), first, count, _ptr(out),
                                       ctypes.byref(epi), _ptr(self._workspace(1 if res1 is None else 2, count)),
                                       _stream()), "pax_ax")
            if maxout is not None:    # maxout: 1-element float32 tensor, holds max(out) afterwards
                _lib.check(self.lib.pax_max_decode(_ptr(maxout), _stream()), "pax_max_decode")
        return out

    def ax_residual(self, res, b, first, count, out, w_variant="matlab", sumsq=None):
        """out = W * (b - Ax(res)) for angles [first, first+count).  (PAX_AX_RESIDUAL)"""
        hz, hy, hx, thr = self._w_box(w_variant)
        epi = _lib.PaxAxEpilogue(_lib.PAX_AX_RESIDUAL, 1.0, 0.0, 0.0, b.data_ptr(), hz, hy, hx, thr,
                                 sumsq.data_ptr() if sumsq is not None else None, None, None, None)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_ax(self._h, _ptr(res), None, first, count, _ptr(out),
                                       ctypes.byref(epi), _ptr(self._workspace(1, count)), _stream()), "pax_ax")
        return out

    def atb(self, proj, first, count, out_res, mode=_lib.PAX_ATB_STORE, inv_v=None, lam=1.0, noneg=True,
            rows=None):
        """rows=(part, parts): only the volume rows y = part (mod parts) (multi-GPU row shares)."""
        epi = _lib.PaxAtbEpilogue(mode, int(bool(noneg)), float(lam),
                                  inv_v.data_ptr() if inv_v is not None else None,
                                  int(rows[0]) if rows else 0, int(rows[1]) if rows else 0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_atb(self._h, _ptr(proj), first, count, _ptr(out_res),
                                        ctypes.byref(epi), _stream()), "pax_atb")
        return out_res

    def sart_update(self, x_res, upd_res, inv_v, lam, noneg=True):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_sart_update(self._h, _ptr(x_res), _ptr(upd_res), _ptr(inv_v),
                                                float(lam), int(bool(noneg)), _stream()), "pax_sart_update")
        return x_res

    def compute_w(self, first=0, count=None, w_variant="matlab", out=None):
        count = self.n_angles - first if count is None else count
        if out is None:
            out = torch.empty((count, self.nv, self.nu), dtype=torch.float32, device=self.device)
        hz, hy, hx, thr = self._w_box(w_variant)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_compute_w(self._h, first, count, hz, hy, hx, thr, _ptr(out), _stream()),
                       "pax_compute_w")
        return out

    def compute_v(self, first, count, out=None):
        if out is None:
            out = torch.empty((self.ny, self.nx), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pax_compute_v(self._h, first, count, self.geo.v_scale(), _ptr(out),
                                              _stream()), "pax_compute_v")
        return out


def lincomb(out, x, y, a, b):
    """out = a*x + b*y on float32 CUDA tensors of equal size (libpax k_lincomb)."""
    lib = _lib.load()
    assert out.numel() == x.numel() == y.numel() and out.is_contiguous() and x.is_contiguous() and y.is_contiguous()
    with torch.cuda.device(out.device):
        _lib.check(lib.pax_lincomb(_ptr(out), _ptr(x), _ptr(y), float(a), float(b), out.numel(), _stream()),
                   "pax_lincomb")
    return out


def reduce_max(x):
    """max over a float32 CUDA tensor -> 1-element CUDA tensor (stays on the device)."""
    lib = _lib.load()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.pax_reduce_max(_ptr(x), x.numel(), _ptr(out), _stream()), "pax_reduce_max")
    return out


# ---------------------------------------------------------------- reference call surface
def _to_cuda(arr, device, what):
    """NumPy / torch input -> (float32 contiguous CUDA tensor, was_numpy)."""
    if isinstance(arr, np.ndarray):
        if arr.dtype != np.float32:
            raise TypeError(f"{what} must be float32, got {arr.dtype}")
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=False), True
    if isinstance(arr, torch.Tensor):
        if arr.dtype != torch.float32:
            raise TypeError(f"{what} must be float32, got {arr.dtype}")
        if not arr.is_cuda:
            return arr.contiguous().to(device), False
        return arr.contiguous(), False
    raise TypeError(f"{what} must be a numpy array or torch tensor, got {type(arr)}")


def Ax(img, geo, angles, projection_type="interpolated", **kwargs):
    """Forward projection, ``tigre.Ax`` surface (reference call site :86).

    Extra keyword ``scatter=(art, scale)`` projects a second volume in the same pass and
    returns ``Ax(img) + scale*Ax(art)`` -- the reference's image-domain scatter transfer
    (AddImages.cxx:57-62) moved into the projection epilogue by linearity.
    """
    if projection_type not in ("interpolated", None):
        raise ValueError(f"projection_type {projection_type!r} is not on this path "
                         "(only 'interpolated', the projector the reference calls)")
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    dev = torch.device("cuda", _device_from_gpuids(kwargs.get("gpuids")))
    vol, was_np = _to_cuda(img, dev, "img")
    devs = _multi(kwargs.get("gpuids"))
    if devs is not None and kwargs.get("plan") is None:
        from .multidev import MultiDevice
        md = MultiDevice(geo, angles, devs)
        md.plans[0]._check_vol(vol)
        art_t, a1 = None, 0.0
        if kwargs.get("scatter") is not None:
            art, a1 = kwargs["scatter"]
            art_t, _ = _to_cuda(art, vol.device, "scatter volume")
        return md.gather_projections(md.ax(vol, art_t, a1), was_np)
    plan = kwargs.get("plan") or Plan(geo, angles, device=vol.device.index)
    res0 = plan.pack(vol)
    res1, a1 = None, 0.0
    if kwargs.get("scatter") is not None:
        art, a1 = kwargs["scatter"]
        art_t, _ = _to_cuda(art, vol.device, "scatter volume")
        res1 = plan.pack(art_t)
    out = plan.ax(res0, res1=res1, a0=1.0, a1=float(a1))
    return out.cpu().numpy() if was_np else out


def Atb(projections, geo, angles, backprojection_type="FDK", **kwargs):
    """Backprojection, ``tigre.Atb`` surface ('FDK' weights)."""
    if backprojection_type not in ("FDK", None):
        raise ValueError(f"backprojection_type {backprojection_type!r} is not on this path "
                         "(only 'FDK', the weights ossart uses)")
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    devs = _multi(kwargs.get("gpuids"))
    if devs is not None and kwargs.get("plan") is None:
        if isinstance(projections, np.ndarray) and projections.dtype != np.float32:
            raise TypeError(f"projections must be float32, got {projections.dtype}")
        if isinstance(projections, torch.Tensor) and projections.dtype != torch.float32:
            raise TypeError(f"projections must be float32, got {projections.dtype}")
        from .multidev import MultiDevice
        md = MultiDevice(geo, angles, devs)
        want = (md.n,) + tuple(int(v) for v in geo.nDetector)
        if tuple(projections.shape) != want:
            raise ValueError(f"projections shape {tuple(projections.shape)} != (N, *geo.nDetector) {want}")
        out = md.atb(projections)
        return out.cpu().numpy() if isinstance(projections, np.ndarray) else out
    dev = torch.device("cuda", _device_from_gpuids(kwargs.get("gpuids")))
    proj, was_np = _to_cuda(projections, dev, "projections")
    plan = kwargs.get("plan") or Plan(geo, angles, device=proj.device.index)
    if tuple(proj.shape) != (plan.n_angles, plan.nv, plan.nu):
        raise ValueError(f"projections shape {tuple(proj.shape)} != (N, *geo.nDetector) "
                         f"{(plan.n_angles, plan.nv, plan.nu)}")
    res = plan.new_volume()
    plan.atb(proj, 0, plan.n_angles, res, _lib.PAX_ATB_STORE)
    out = plan.unpack(res)
    return out.cpu().numpy() if was_np else out
