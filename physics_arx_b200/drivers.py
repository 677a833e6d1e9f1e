"""Batch-variation (config C4) and motion-artifact (config C5) drivers over the same kernels.

BASELINE.json configs 4 and 5 / SURVEY.md 8(f)-3.  The reference has no code for either: the
variation list it does have is the seven (alpha, beta) artifact volumes it loops over
(pCT_CBCT_Reconstruction.py:44-45), each pushed through the same Ax -> noise -> ossart sequence
(:86-103); the motion case is only named in README.md:10.

C4  one planning CT, one artifact volume, many (scatter scale, I0, angle stride) variations.  By
    linearity of Ax the two line-integral sets P_ct, P_art are computed ONCE (one two-channel pass,
    PaxAxEpilogue.out1) and every variation is P_ct + s*P_art (pax_lincomb) -> noise -> OS-SART.
    Across GPUs variations are independent: rank r takes variations r::world, no collective.
C5  each projection angle is forward-projected from the volume of its respiratory phase, then the
    motion-corrupted set is reconstructed as usual.
"""
from __future__ import annotations

import ctypes
import itertools

import numpy as np
import torch

from . import _lib
from .algorithms.ossart import OSSARTState
from .projector import Plan, _ptr, _stream, _to_cuda, lincomb, reduce_max


# ------------------------------------------------------------------------------- C4
class VariationBatch:
    """Shared-projection fan-out of sCBCT variations for one planning CT."""

    def __init__(self, geo, angles, niter=5, blocksize=20, sigma=4.0, lmbda=1.0, lmbda_red=0.99,
                 device=None):
        if not torch.cuda.is_available():
            raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.geo = geo
        self.angles = np.asarray(angles, dtype=np.float64)
        self.niter, self.blocksize, self.sigma = int(niter), int(blocksize), float(sigma)
        self.lmbda, self.lmbda_red = float(lmbda), float(lmbda_red)
        with torch.cuda.device(self.device):
            self.plan = Plan(geo, self.angles, device=self.device.index)
        self._states = {}
        self.p_ct = self.p_art = None

    def project(self, ct, art):
        """One two-channel pass -> P_ct, P_art (N,V,U), kept on the device."""
        p = self.plan
        with torch.cuda.device(self.device):
            ct_t, _ = _to_cuda(ct, self.device, "ct")
            art_t, _ = _to_cuda(art, self.device, "artifact")
            r0, r1 = p.pack(ct_t), p.pack(art_t)
            self.p_ct = torch.empty((p.n_angles, p.nv, p.nu), dtype=torch.float32, device=self.device)
            self.p_art = torch.empty_like(self.p_ct)
            p.ax(r0, res1=r1, out=self.p_ct, out1=self.p_art)
        return self.p_ct, self.p_art

    def _state(self, stride):
        if stride not in self._states:
            with torch.cuda.device(self.device):
                self._states[stride] = OSSARTState(None, self.geo, self.angles[::stride], self.blocksize,
                                                   device=self.device.index, distributed=False)
        return self._states[stride]

    def variation(self, scatter_scale, I0, stride=1, seed=0):
        """One variation -> reconstructed (Z,Y,X) CUDA tensor."""
        if self.p_ct is None:
            raise RuntimeError("call project(ct, art) first")
        st = self._state(int(stride))
        p = st.plan
        with torch.cuda.device(self.device):
            a = self.p_ct[::stride].contiguous() if stride != 1 else self.p_ct
            b = self.p_art[::stride].contiguous() if stride != 1 else self.p_art
            lincomb(st.b, a, b, 1.0, float(scatter_scale))
            m = reduce_max(st.b)
            # global pixel index of the strided angle set: angle k*stride of the full scan
            npix = p.nv * p.nu
            _lib.check(p.lib.pax_ctnoise(_ptr(st.b), st.b.numel(), ctypes.c_uint64(0), npix,
                                         ctypes.c_uint64(npix * int(stride)), float(I0), 0.0, self.sigma,
                                         _ptr(m), ctypes.c_uint64(int(seed)), _stream()), "pax_ctnoise")
            x = p.new_volume()
            lam = self.lmbda
            for _ in range(self.niter):
                st.iteration(x, lam, True, False)
                lam *= self.lmbda_red
            return p.unpack(x)

    def run(self, ct, art, scatter_scales, I0s, strides=(1,), seed=0, rank=0, world=1, to_numpy=True):
        """All (scale, I0, stride) combinations; rank r of `world` takes every world-th one."""
        self.project(ct, art)
        combos = list(itertools.product(scatter_scales, I0s, strides))
        out = []
        for i, (s, i0, stride) in enumerate(combos):
            if i % world != rank:
                continue
            vol = self.variation(s, i0, stride, seed=seed + i)
            out.append(dict(index=i, scatter_scale=float(s), I0=float(i0), stride=int(stride),
                            volume=vol.cpu().numpy() if to_numpy else vol))
        return out


def config_c4():
    """BASELINE config 4 / SURVEY.md 8(d): 4 scatter scales x 4 I0 levels x 2 angle strides = 32."""
    return dict(scatter_scales=np.linspace(0.0, 1.5, 4), I0s=(1e5, 1e6, 1e7, 1e8), strides=(1, 2))


# ------------------------------------------------------------------------------- C5
def respiratory_phases(n_angles, n_phases=10, scan_s=60.0, period_s=4.0):
    """phase(theta_k) = floor(n_phases * frac(k * T_proj / T_breath))  (SURVEY.md 8(d), C5)."""
    t = np.arange(n_angles) * (scan_s / n_angles)
    return np.floor(n_phases * np.mod(t / period_s, 1.0)).astype(np.int64) % n_phases


def phase_volumes(vol, dz_mm, n_phases=10, amplitude_mm=15.0):
    """Seeded-free synthetic 4D CT: sinusoidal superior-inferior shift of `vol` (Z,Y,X), linear in z."""
    vol = np.asarray(vol, dtype=np.float32)
    nz = vol.shape[0]
    out = np.empty((n_phases,) + vol.shape, dtype=np.float32)
    k = np.arange(nz, dtype=np.float64)
    for ph in range(n_phases):
        shift = 0.5 * amplitude_mm * (1.0 - np.cos(2 * np.pi * ph / n_phases)) / dz_mm     # voxels
        src = k - shift
        k0 = np.floor(src).astype(np.int64)
        f = (src - k0).astype(np.float32)[:, None, None]
        a = vol[np.clip(k0, 0, nz - 1)]
        b = vol[np.clip(k0 + 1, 0, nz - 1)]
        out[ph] = (1 - f) * a + f * b
    return out


def motion_projections(phase_vols, phase_of_angle, geo, angles, out=None):
    """Ax where angle k sees volume phase_of_angle[k].  phase_vols: (P,Z,Y,X) NumPy or CUDA tensor.

    Angles are grouped by phase so each phase volume is packed and swept once; the result is
    returned in the scan's own angle order, (N,V,U) on the device.
    """
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    angles = np.asarray(angles, dtype=np.float64)
    phase_of_angle = np.asarray(phase_of_angle, dtype=np.int64)
    if phase_of_angle.shape != angles.shape[:1]:
        raise ValueError("phase_of_angle must have one entry per angle")
    n_ph = phase_vols.shape[0]
    if phase_of_angle.min() < 0 or phase_of_angle.max() >= n_ph:
        raise ValueError("phase index out of range")
    order = np.argsort(phase_of_angle, kind="stable")
    plan = Plan(geo, angles[order])
    sorted_out = torch.empty((len(angles), plan.nv, plan.nu), dtype=torch.float32, device=dev)
    res = plan.new_volume()
    pos = 0
    for ph in range(n_ph):
        cnt = int((phase_of_angle == ph).sum())
        if cnt == 0:
            continue
        v, _ = _to_cuda(phase_vols[ph], dev, "phase volume")
        plan.pack(v, out=res)
        plan.ax(res, pos, cnt, out=sorted_out[pos:pos + cnt])
        pos += cnt
    if out is None:
        out = torch.empty_like(sorted_out)
    out[torch.as_tensor(order, device=dev)] = sorted_out
    return out
