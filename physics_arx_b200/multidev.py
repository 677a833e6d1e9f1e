"""One process driving several GPUs: TIGRE's ``gpuids`` semantics.

The reference builds ``gpuids`` from EVERY device that carries the first GPU's name and hands it to each call
(pCT_CBCT_Reconstruction.py:24-32, :86, :103); TIGRE then splits the call over those devices inside one
process.  Same here: ``Ax`` / ``Atb`` / ``ossart`` given a ``gpuids`` with more than one device route through
:class:`MultiDevice` -- a plan, buffers and a stream per device, the angle split of ``dist.shard_blocks``
(the angles of every OS-SART subset interleaved over the devices), the volume replicated, and the partial
volume updates summed with single-process NCCL (``torch.cuda.nccl``) once per subset.  The one-process-per-GPU
form (``torchrun`` + ``group=``, algorithms/ossart.py) remains the faster one: no Python loop over devices and
the fused reduce+update kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .dist import shard_blocks, subset_geometry
from .geometry import order_subsets
from .projector import Plan


def devices_of(gpuids):
    """Device indices a ``gpuids`` argument names (None -> the current device)."""
    if gpuids is None:
        return [torch.cuda.current_device()]
    devs = getattr(gpuids, "devices", gpuids)
    if isinstance(devs, (list, tuple, range)):
        devs = [int(d) for d in devs]
        if not devs:
            raise ValueError("gpuids is empty")
        return devs
    return [int(devs)]


def _all_reduce(tensors):
    """Sum a list of equally sized tensors, one per device, in place on every device."""
    if len(tensors) == 1:
        return
    try:
        import torch.cuda.nccl as nccl
        nccl.all_reduce(tensors)
        return
    except Exception:                      # noqa: BLE001 -- no single-process NCCL: reduce through device 0
        pass
    acc = tensors[0]
    for t in tensors[1:]:
        acc += t.to(acc.device)
    for t in tensors[1:]:
        t.copy_(acc)


class MultiDevice:
    def __init__(self, geo, angles, devices, blocksize=None, OrderStrategy=None, seed=0):
        if not torch.cuda.is_available():
            raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
        angles = np.asarray(angles, dtype=np.float64)
        if angles.ndim == 2:
            geo.check_geo(angles)
            angles = angles[:, 0]
        self.geo, self.angles, self.n = geo, angles, angles.shape[0]
        self.devices = [torch.device("cuda", int(d)) for d in devices]
        G = len(self.devices)
        self.blocks = (order_subsets(self.n, blocksize, OrderStrategy, seed) if blocksize
                       else [np.arange(self.n, dtype=np.int64)])
        self.local, self.ranges, self.plans = [], [], []
        for i, dev in enumerate(self.devices):
            local, ranges = shard_blocks(self.blocks, i, G)
            idx = local if len(local) else np.array([0])      # a device may own no angle: keep a plan
            with torch.cuda.device(dev):
                plan = Plan(subset_geometry(geo, idx, self.n), angles[idx], device=dev.index)
            plan.w_geo = geo
            self.local.append(local)
            self.ranges.append(ranges if len(local) else [(0, 0)] * len(self.blocks))
            self.plans.append(plan)

    # ------------------------------------------------------------------ helpers
    def _replicate(self, res0):
        """The resident volume of device 0 on every device (peer copies)."""
        out = [res0]
        for dev in self.devices[1:]:
            out.append(res0.to(dev, non_blocking=True))
        return out

    def _sync(self):
        for dev in self.devices:
            torch.cuda.synchronize(dev)

    # --------------------------------------------------------------- projectors
    def ax(self, vol, res1=None, a1=0.0):
        """vol: (Z,Y,X) float32 tensor on device 0 (or host) -> list of per-device projection tensors and the
        angle indices they hold."""
        d0 = self.devices[0]
        with torch.cuda.device(d0):
            res = self._replicate(self.plans[0].pack(vol.to(d0)))
            art = self._replicate(self.plans[0].pack(res1.to(d0))) if res1 is not None else [None] * len(res)
        outs = []
        for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
            k = len(self.local[i])
            with torch.cuda.device(dev):
                outs.append(plan.ax(res[i], 0, k, res1=art[i], a0=1.0, a1=float(a1)) if k else
                            torch.empty((0, plan.nv, plan.nu), dtype=torch.float32, device=dev))
        return outs

    def gather_projections(self, outs, to_numpy):
        nv, nu = (int(v) for v in self.geo.nDetector)
        if to_numpy:
            full = np.empty((self.n, nv, nu), dtype=np.float32)
            for loc, o in zip(self.local, outs):
                if len(loc):
                    full[loc] = o.cpu().numpy()
            return full
        d0 = self.devices[0]
        full = torch.empty((self.n, nv, nu), dtype=torch.float32, device=d0)
        for loc, o in zip(self.local, outs):
            if len(loc):
                full[torch.as_tensor(loc, device=d0)] = o.to(d0)
        return full

    def scatter_projections(self, proj):
        """(N,V,U) NumPy array or tensor -> per-device tensors of the angles each device owns."""
        parts = []
        for dev, loc in zip(self.devices, self.local):
            if isinstance(proj, np.ndarray):
                part = torch.from_numpy(np.ascontiguousarray(proj[loc])).to(dev)
            else:
                part = proj[torch.as_tensor(loc, device=proj.device, dtype=torch.long)].to(dev)
            parts.append(part.contiguous())
        return parts

    def atb(self, proj):
        """Sum over the devices of each device's backprojection -> (Z,Y,X) tensor on device 0."""
        parts = self.scatter_projections(proj)
        partial = []
        for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
            with torch.cuda.device(dev):
                upd = plan.new_volume()
                if len(self.local[i]):
                    plan.atb(parts[i], 0, len(self.local[i]), upd, _lib.PAX_ATB_STORE)
                partial.append(upd)
        _all_reduce(partial)
        with torch.cuda.device(self.devices[0]):
            return self.plans[0].unpack(partial[0])

    # ------------------------------------------------------------------ OS-SART
    def ossart(self, proj, niter, lmbda=1.0, lmbda_red=0.99, noneg=True, init=None, w_variant="matlab",
               v_variant="A", computel2=False, verbose=False):
        from .algorithms.ossart import _v_variant_b
        G, nsub = len(self.devices), len(self.blocks)
        b = self.scatter_projections(proj)
        # V_s: partial sums over each device's angles, summed over the devices
        V = []
        for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
            with torch.cuda.device(dev):
                v = torch.zeros((nsub, plan.ny, plan.nx), dtype=torch.float32, device=dev)
                if v_variant == "A":
                    for s, (first, count) in enumerate(self.ranges[i]):
                        if count:
                            plan.compute_v(first, count, out=v[s])
                elif v_variant == "B":
                    v.copy_(torch.from_numpy(_v_variant_b(self.geo, self.angles, self.blocks)).to(dev))
                    v /= G
                else:
                    raise ValueError(f"unknown V variant {v_variant!r}")
                V.append(v)
        _all_reduce(V)
        inv_v, x, upd, resid, sumsq = [], [], [], [], []
        for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
            with torch.cuda.device(dev):
                inv_v.append(torch.where(V[i] > 1e-12, 1.0 / V[i], torch.zeros_like(V[i])).contiguous())
                x.append(plan.new_volume())
                upd.append(plan.new_volume())
                mc = max((c for _, c in self.ranges[i]), default=0)
                resid.append(torch.empty((max(mc, 1), plan.nv, plan.nu), dtype=torch.float32, device=dev))
                sumsq.append(torch.zeros(1, dtype=torch.float64, device=dev))
        if init is not None:
            with torch.cuda.device(self.devices[0]):
                self.plans[0].pack(init.to(self.devices[0]), out=x[0])
            for i in range(1, G):
                x[i].copy_(x[0], non_blocking=True)
        lam, l2 = float(lmbda), []
        for it in range(int(niter)):
            for s_ in sumsq:
                s_.zero_()
            for s in range(nsub):
                for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
                    first, count = self.ranges[i][s]
                    with torch.cuda.device(dev):
                        if count:
                            plan.ax_residual(x[i], b[i][first:first + count], first, count, resid[i], w_variant,
                                             sumsq[i] if computel2 else None)
                            plan.atb(resid[i], first, count, upd[i], _lib.PAX_ATB_STORE)
                        else:
                            upd[i].zero_()
                _all_reduce(upd)
                for i, (dev, plan) in enumerate(zip(self.devices, self.plans)):
                    with torch.cuda.device(dev):
                        plan.sart_update(x[i], upd[i], inv_v[i][s], lam, noneg)
            if computel2:
                l2.append(float(np.sqrt(sum(float(s_.item()) for s_ in sumsq))))
            if verbose:
                print(f"OS-SART iteration {it + 1}/{niter}, lambda={lam:.4f} ({G} devices)")
            lam *= lmbda_red
        with torch.cuda.device(self.devices[0]):
            out = self.plans[0].unpack(x[0])
        for plan in self.plans:
            plan.check_fault()
        return (out, np.array(l2)) if computel2 else out
