"""The whole sCBCT generation loop of the reference, device-resident:

    forward projection of (planning CT + scale * artifact volume)     [:86, AddImages.cxx:57-62]
    -> Poisson + electronic noise                                      [:88]
    -> OS-SART reconstruction                                          [:101-103]

(file:line = Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py).
One ``SCBCTPipeline`` holds the plan and every buffer for a (geometry, angle list); a run
allocates nothing.  With a torch.distributed group the angles of each OS-SART subset are
interleaved over the ranks: the simulation Ax and the noise need no
collective beyond one scalar max, OS-SART all-reduces one partial volume update per subset.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .algorithms.ossart import OSSARTState
from .projector import _ptr, _stream


class SCBCTPipeline:
    def __init__(self, geo, angles, niter=20, blocksize=20, I0=1e8, sigma=4.0, scatter_scale=1.0,
                 lmbda=1.0, lmbda_red=0.99, seed=0, group=None, device=None, w_variant="matlab",
                 v_variant="A", shard="auto", fused_update=None):
        if not torch.cuda.is_available():
            raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.geo = geo
        self.angles = np.asarray(angles, dtype=np.float64)
        self.niter, self.blocksize = int(niter), int(blocksize)
        self.I0, self.sigma, self.scatter_scale = float(I0), float(sigma), float(scatter_scale)
        self.lmbda, self.lmbda_red, self.seed = float(lmbda), float(lmbda_red), int(seed)
        self.group = group
        self.v_variant = v_variant
        # OS-SART state with an empty rank-local projection buffer; simulate() fills st.b
        with torch.cuda.device(self.device):
            self.st = OSSARTState(None, geo, self.angles, blocksize, None, w_variant=w_variant,
                                  v_variant=v_variant, group=group, distributed=group is not None,
                                  device=self.device.index, shard=shard, fused_update=fused_update)
        st = self.st
        self.plan = st.plan
        self.n_local = int(st.b.shape[0])
        self.local_idx = st.local_idx
        p = self.plan
        self.ct_res = p.new_volume()
        self.art_res = p.new_volume()
        self.x_res = st.new_iterate()      # the symmetric buffer when the fused multi-GPU update is in use
        self.out = torch.empty((p.nz, p.ny, p.nx), dtype=torch.float32, device=self.device)
        self.launches = 0
        self._max = torch.zeros(1, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ stages
    def simulate(self, ct_dev, art_dev=None):
        """Noisy projections of this rank's angles into st.b (device resident)."""
        import torch.distributed as dist
        p, st = self.plan, self.st
        p.pack(ct_dev, out=self.ct_res)
        self.launches += 1
        if art_dev is not None:
            # one variation: add the artifact volume in the image domain (what the reference does,
            # AddImages.cxx:57-62) and project one channel.  The two-channel epilogue
            # (Plan.ax(res1=..., a1=...)) is for fanning several scatter scales out of one pass.
            p.pack(art_dev, out=self.art_res)
            p.axpy(self.ct_res, self.art_res, self.scatter_scale)
            self.launches += 2
        m = self._max
        if self.n_local:
            p.ax(self.ct_res, 0, self.n_local, out=st.b, maxout=m)     # max(p) fused in the epilogue
            self.launches += p.ax_launches(1) + 1                       # (quad pack,) projector, decode
        else:
            m.fill_(-3e38)
        if self.group is not None:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        # noise is keyed by the GLOBAL pixel index, so the result does not depend on the sharding
        nu_full = int(np.asarray(self.geo.nDetector)[1])
        nv_full = int(np.asarray(self.geo.nDetector)[0])
        npix = nv_full * nu_full
        lib = p.lib
        rows, cols = st.v1 - st.v0, st.u1 - st.u0
        # the rank's buffer as runs of `inner` consecutive global pixels, `stride` apart: whole projections, a
        # band of rows ([angles][rows*nu] out of [angles][npix]) or a band of columns ([angles*rows][cols] out
        # of rows of nu)
        if cols != nu_full:
            inner, stride, first = cols, nu_full, st.u0
        elif rows != nv_full:
            inner, stride, first = rows * nu_full, npix, st.v0 * nu_full
        else:
            inner, stride, first = 0, 0, 0
        with torch.cuda.device(self.device):
            pos = 0
            if self.n_local:
                # one launch per run of consecutive global angles: one in all on a single GPU, one per subset in the
                # default multi-GPU scheme (a rank owns a contiguous share of every subset's angles), one per angle
                # with the interleaved shares of PAX_MGPU=reduce (656 / N launches of ~10 us, once per volume)
                for g0, cnt in _contiguous_runs(self.local_idx):
                    view = st.b[pos:pos + cnt]
                    _lib.check(lib.pax_ctnoise(_ptr(view), view.numel(),
                                               ctypes.c_uint64(int(g0) * npix + first), inner,
                                               ctypes.c_uint64(stride), self.I0, 0.0,
                                               self.sigma, _ptr(m), ctypes.c_uint64(self.seed), _stream()),
                               "pax_ctnoise")
                    self.launches += 1
                    pos += cnt
        return st.b

    def reconstruct(self):
        st = self.st
        p = self.plan
        if self.v_variant == "A":      # set_v runs inside every ossart() call of the reference
            st.inv_v = st.compute_inv_v()
            self.launches += len(st.blocks)
        self.x_res.zero_()
        lam = self.lmbda
        n0 = st.launches
        for _ in range(self.niter):
            st.iteration(self.x_res, lam, True, False)
            lam *= self.lmbda_red
        self.launches += st.launches - n0
        p.unpack(self.x_res, out=self.out)
        self.launches += 1
        p.check_fault()                # covers simulate() as well: same plan, sticky flag
        return self.out

    # ------------------------------------------------------------- entry points
    def run_device(self, ct_dev, art_dev=None):
        """Inputs already in HBM -> reconstructed (Z,Y,X) volume in HBM."""
        with torch.cuda.device(self.device):
            self.simulate(ct_dev, art_dev)
            return self.reconstruct()

    def run_host(self, ct_host, art_host=None, out_host=None):
        """Host buffers in -> host volume out (the call a user of the reference script makes)."""
        with torch.cuda.device(self.device):
            ct = _as_tensor(ct_host).to(self.device, non_blocking=True)
            art = _as_tensor(art_host).to(self.device, non_blocking=True) if art_host is not None else None
            out = self.run_device(ct, art)
            if out_host is None:
                out_host = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out_host


def _contiguous_runs(idx):
    """[(global_first_angle, count)] covering a sorted-by-subset local index list."""
    runs = []
    for g in np.asarray(idx):
        g = int(g)
        if runs and runs[-1][0] + runs[-1][1] == g:
            runs[-1][1] += 1
        else:
            runs.append([g, 1])
    return [(a, b) for a, b in runs]


def _as_tensor(a):
    if isinstance(a, np.ndarray):
        if a.dtype != np.float32:
            raise TypeError("host volumes must be float32")
        return torch.from_numpy(np.ascontiguousarray(a))
    return a


def generate_scbct(ct, geo, angles, artifact=None, scatter_scale=1.0, I0=1e8, sigma=4.0, niter=30,
                   blocksize=20, seed=0, **kw):
    """One-call form: NumPy (Z,Y,X) planning CT [+ artifact volume] -> NumPy sCBCT volume."""
    pipe = SCBCTPipeline(geo, angles, niter=niter, blocksize=blocksize, I0=I0, sigma=sigma,
                         scatter_scale=scatter_scale, seed=seed, **kw)
    return pipe.run_host(ct, artifact).numpy().copy()
