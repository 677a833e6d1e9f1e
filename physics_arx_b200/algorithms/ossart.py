"""Device-resident OS-SART (``tigre.algorithms.ossart``; reference call site
pCT_CBCT_Reconstruction.py:101-103, algorithm SURVEY.md A.6).

    x <- max(0, x + lambda * V_s^-1 * Atb_FDK(W_s * (b_s - Ax(x))))    per subset s
    lambda <- lambda * lmbda_red                                        per iteration

TIGRE runs each Ax/Atb host-array in / host-array out and does the residual, weight
and update arithmetic in NumPy between PCIe round trips.  Here the
iterate, the projections and the V maps stay in HBM for the whole run: one kernel
forms W*(b - Ax(x)) (W analytic, never materialised), one kernel backprojects and
applies the update.  With a ``torch.distributed`` group the angles of every subset are
interleaved over the ranks, each rank backprojects its residuals into a partial update,
the partials are summed with one NCCL all-reduce per subset and every rank applies the
(non-linear) clip after the sum.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..geometry import order_subsets
from ..projector import Plan, _device_from_gpuids, _to_cuda
from ..dist import (band_axis, choose_shard, col_window_geometry, row_window_geometry, shard_blocks,
                    shard_rows, subset_geometry)


def _vp(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def _vp_int(v):
    import ctypes
    return ctypes.c_void_p(int(v) if v else 0)


def _cur_stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_NVTX = __import__("os").environ.get("PAX_NVTX", "") == "1"


def os_environ(name):
    import os
    return os.environ.get(name, "")


def _dist_info(group, distributed):
    import torch.distributed as dist
    if distributed is False or not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    if distributed is None and group is None:
        return None, 0, 1
    g = group if group is not None else dist.group.WORLD
    return g, dist.get_rank(g), dist.get_world_size(g)


class OSSARTState:
    """Set-up products of one OS-SART problem (the role of TIGRE's IterativeReconAlg)."""

    def __init__(self, proj, geo, angles, blocksize=20, OrderStrategy=None, w_variant="matlab",
                 v_variant="A", group=None, distributed=None, device=None, seed=0, shard="auto",
                 fused_update=None):
        import torch.distributed as dist
        angles = np.asarray(angles, dtype=np.float64)
        if angles.ndim == 2:
            geo.check_geo(angles)
            angles = angles[:, 0]
        n = angles.shape[0]
        self.group, self.rank, self.world = _dist_info(group, distributed)
        self.geo, self.n = geo, n
        self.w_variant = w_variant
        self.blocks = order_subsets(n, blocksize, OrderStrategy, seed)
        dev = proj.device if proj is not None else torch.device(
            "cuda", torch.cuda.current_device() if device is None else int(device))
        nv_full, nu = (int(v) for v in geo.nDetector)
        self.v_variant = v_variant
        self.angles = angles
        # Multi-GPU scheme.  "gather" (default where symmetric memory works): the angles of every subset in
        # consecutive shares, every rank's residual projections broadcast into every rank's copy of the subset's
        # residual set, every rank backprojects ALL of them into its own rows of the volume (with the SART update
        # in the epilogue, as on one GPU) and broadcasts those rows.  63 MB of residuals cross the fabric per C2
        # subset instead of a 144 MB partial update per rank, there is no reduction and no separate update pass,
        # and the result is the single-GPU arithmetic voxel for voxel.  "reduce": every rank backprojects its own
        # residuals into a partial update of the whole volume and the partial updates are summed (fused
        # reduce + update + broadcast kernel, or NCCL all-reduce + update); used for explicit shard= modes.
        self.scheme, self.gsym, self.plan_full = "reduce", None, None
        import os
        if (self.world > 1 and shard == "auto" and fused_update is not False
                and os.environ.get("PAX_MGPU", "gather") == "gather" and os.environ.get("PAX_FUSED_UPDATE", "") != "0"):
            self._setup_gather(dev)
        # the ranks form an A x R grid: the angles of every subset over A groups, the detector over R bands
        # (of columns, or of rows with shard='rows')
        self.grid = choose_shard(self.blocks, self.world, shard) if self.world > 1 else (1, 1)
        if self.scheme == "gather" and (self.plan_full.projector != "fan" or os_environ("PAX_MGPU_GRID") != "auto"):
            # Angles only.  (PAX_MGPU_GRID=auto lets the ranks of an angle group take shares of the detector's
            # column tiles instead -- equal work for every rank where the angles do not divide evenly, their
            # residual columns ADDED into the symmetric set.  Measured at C2 on 8 GPUs: 4 x 2 shares 0.879 s per
            # volume, 8 x 1 with its 2-or-3-angle imbalance 0.860 s: the adds cost what the balance gains.)
            self.grid = (self.world, 1)
        A, R = self.grid
        self.axis = band_axis(shard)
        self.shard = "angles" if R == 1 else (self.axis if A == 1 else "grid")
        ra, rr = self.rank % A, self.rank // A
        self.v0, self.v1, self.u0, self.u1 = 0, nv_full, 0, nu
        if R > 1 and self.axis == "rows":
            self.v0, self.v1 = shard_rows(nv_full, rr, R)
        elif R > 1:
            self.u0, self.u1 = shard_rows(nu, rr, R, align=8)     # whole 32-byte sectors of a projection row
        local, self.ranges = shard_blocks(self.blocks, ra, A, contiguous=self.scheme == "gather")
        # a rank may own no angle at all (n < A): keep one dummy angle so a plan exists
        plan_idx = local if len(local) else np.array([0])
        sub_geo = subset_geometry(geo, plan_idx, n)
        self.tiles = None              # (part, parts): this rank's share of the detector's column tiles
        plan = None
        if R > 1 and self.scheme == "gather":
            # the R ranks of an angle group project the same angles on shares of the detector's column tiles; their
            # residuals (zeros in the columns of the others) are ADDED into the subset's symmetric residual set
            plan = Plan(sub_geo, angles[plan_idx], device=dev.index)
            plan.set_column_share(rr, R)
            self.tiles = (rr, R)
            self.u0, self.u1 = 0, nu
        elif R > 1 and self.axis == "cols":
            # Where the fan-plane projector runs, the bands are not contiguous: every rank keeps the WHOLE detector
            # and its kernel walks every R-th column tile, dealt out by decreasing chord length -- contiguous bands
            # of a half-fan detector differ by 1.5x in work.  The columns of the other ranks stay zero in the
            # residual, so the backprojections of the ranks still add up exactly.
            plan = Plan(sub_geo, angles[plan_idx], device=dev.index)
            if plan.projector == "fan" and shard != "cols":    # (shard='cols' asks for contiguous bands)
                plan.set_column_share(rr, R)
                self.tiles = (rr, R)
                self.u0, self.u1 = 0, nu
            else:
                plan = None
        empty = self.v1 <= self.v0 or self.u1 <= self.u0
        plan_geo = sub_geo
        if plan is None:
            if R > 1 and self.axis == "rows":       # a band of the detector is a detector of its own
                plan_geo = row_window_geometry(sub_geo, *((0, 1) if empty else (self.v0, self.v1)))
            elif R > 1:
                plan_geo = col_window_geometry(sub_geo, *((0, 1) if empty else (self.u0, self.u1)))
            plan = Plan(plan_geo, angles[plan_idx], device=dev.index)
        if empty:
            self.ranges = [(f, 0) for f, _ in self.ranges]
        self.local_idx = local
        self.plan = plan
        self.plan.w_geo = geo
        n_own = 0 if empty else len(local)
        rows, cols = self.v1 - self.v0, self.u1 - self.u0
        full = rows == nv_full and cols == nu
        if proj is None:                                    # caller fills the rank-local buffer
            self.b = torch.zeros((max(n_own, 1), max(rows, 1), max(cols, 1)), dtype=torch.float32,
                                 device=dev)[:n_own]
        elif len(local) == n and full and np.array_equal(local, np.arange(n)):
            self.b = proj                                   # ordered, single rank: no copy
        else:
            sel = proj[torch.as_tensor(local, device=dev, dtype=torch.long)] if n_own else proj[:0]
            self.b = sel[:, self.v0:self.v1, self.u0:self.u1].contiguous()
        p = self.plan
        self.inv_v = self.compute_inv_v()
        self.max_count = max((c for _, c in self.ranges), default=0)
        # (zeros: with a share of the column tiles the other ranks' columns are never written and must read 0)
        self.resid = torch.zeros((max(self.max_count, 1), p.nv, p.nu), dtype=torch.float32, device=dev)
        self.upd = p.new_volume() if (self.world > 1 and self.scheme == "reduce") else None
        self.fused = None              # symmetric-memory state of the fused reduce + update + broadcast
        if self.world > 1 and fused_update is not False and self.scheme == "reduce":
            self._setup_fused(dev, required=fused_update is True)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.launches = 0              # kernels of libpax launched by subset_step (bench claim)
        self.kernel_events = None      # set to [] to record CUDA events around the Ax kernel
        self.phase_events = {}         # ... and around the other phases of a subset step (bench breakdown)

    def compute_inv_v(self):
        """1 / V_s per subset (A.5).  V_s is a sum over the subset's angles: in the "reduce" scheme the ranks'
        partial sums are all-reduced; in the "gather" scheme every rank evaluates the whole subset itself."""
        import torch.distributed as dist
        p = self.plan
        dev = p.device
        V = torch.zeros((len(self.blocks), p.ny, p.nx), dtype=torch.float32, device=dev)
        if self.v_variant == "A":
            if self.scheme == "gather":
                for s, (gfirst, glen) in enumerate(self.gsym["granges"]):
                    self.plan_full.compute_v(gfirst, glen, out=V[s])
            else:
                for s, (first, count) in enumerate(self.ranges):
                    if count:
                        p.compute_v(first, count, out=V[s])
                if self.tiles is not None:      # every rank of an angle group computed the whole detector's map
                    V /= self.tiles[1]
        elif self.v_variant == "B":
            V.copy_(torch.from_numpy(_v_variant_b(self.geo, self.angles, self.blocks)).to(dev))
            if self.world > 1 and self.scheme == "reduce":
                V /= self.world
        else:
            raise ValueError(f"unknown V variant {self.v_variant!r}")
        if self.world > 1 and self.scheme == "reduce":
            dist.all_reduce(V, group=self.group)
        return torch.where(V > 1e-12, 1.0 / V, torch.zeros_like(V)).contiguous()

    # --------------------------------------------- gather scheme (csrc/plan.cu: pax_bcast_chunks; atb.cu row shares)
    def _setup_gather(self, dev):
        """Symmetric buffers for the iterate and for one subset's residual projections, plus a plan over ALL
        angles in subset order (the backprojection of a subset runs over every rank's residuals).  Leaves
        scheme == "reduce" when symmetric memory cannot be set up on every rank."""
        import ctypes
        import torch.distributed as dist
        g = self.group if self.group is not None else dist.group.WORLD
        if self.world > _lib.PAX_MAX_PEERS:
            return
        ok, st = 1, None
        try:
            import torch.distributed._symmetric_memory as symm
            if dist.get_backend(g) != "nccl":
                raise RuntimeError("the gather scheme needs the nccl backend")
            order = np.concatenate([np.asarray(b) for b in self.blocks])
            with torch.cuda.device(dev):
                full = Plan(subset_geometry(self.geo, order, self.n), self.angles[order], device=dev.index)
                full.w_geo = self.geo
                maxb = max(len(b) for b in self.blocks)
                npix = full.nv * full.nu
                if (npix * 4) % 16 or (full.nxp_nzp_bytes % 16):
                    raise RuntimeError("projection / volume rows are not multiples of 16 bytes")
                x = symm.empty(full.vol_elems, dtype=torch.float32, device=dev)
                r = symm.empty(maxb * npix, dtype=torch.float32, device=dev)
                x.zero_()
                r.zero_()
                hx, hr = symm.rendezvous(x, g), symm.rendezvous(r, g)
            mc_x = int(getattr(hx, "multicast_ptr", 0) or 0)
            mc_r = int(getattr(hr, "multicast_ptr", 0) or 0)
            use_mc = bool(mc_x and mc_r) and os_environ("PAX_FUSED_UPDATE") != "p2p"
            arr = ctypes.c_uint64 * self.world
            pos, granges = 0, []
            for b in self.blocks:
                granges.append((pos, len(b)))
                pos += len(b)
            st = dict(x=x, r=r.view(maxb, full.nv, full.nu), hx=hx, hr=hr, mc_x=mc_x if use_mc else 0,
                      mc_r=mc_r if use_mc else 0, x_peers=arr(*[int(v) for v in hx.buffer_ptrs]),
                      r_peers=arr(*[int(v) for v in hr.buffer_ptrs]), granges=granges,
                      kind="multicast" if use_mc else "peer pointers", npix=npix)
        except Exception as e:                     # noqa: BLE001 -- any failure means "not available here"
            ok = 0
            self.gather_error = repr(e)
        try:                                       # every rank must take the same path
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=g)
            ok = int(flag.item())
        except Exception:                          # noqa: BLE001
            ok = 0
        if ok:
            self.scheme, self.gsym, self.plan_full = "gather", st, full

    # ------------------------------------------------- fused collective (csrc/plan.cu: k_sart_update_fused)
    def _setup_fused(self, dev, required=False):
        """Symmetric (peer-mapped) buffers for the partial update and the iterate, so that one kernel can
        reduce, update and broadcast over NVLink (include/pax.h: pax_sart_update_fused).  Falls back to
        all-reduce + local update when symmetric memory cannot be set up (no peer access, old torch)."""
        import ctypes
        import os
        # PAX_FUSED_UPDATE: unset -> on (it skips the columns no ray of the subset sees, about half of the plane,
        # which an all-reduce cannot); "0" off: NCCL all-reduce + local update; "1" on (multicast when the fabric
        # has it); "p2p" on, with peer pointers
        import torch.distributed as dist
        mode = os.environ.get("PAX_FUSED_UPDATE", "")
        if not required and mode == "0":
            return
        if self.world > _lib.PAX_MAX_PEERS:        # the kernel's peer table (csrc/plan.cu); larger domains use NCCL
            if required:
                raise ValueError(f"fused update supports at most {_lib.PAX_MAX_PEERS} ranks, got {self.world}")
            return
        g = self.group if self.group is not None else dist.group.WORLD
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            if dist.get_backend(g) != "nccl":
                raise RuntimeError("fused update needs the nccl backend")
            n = self.plan.vol_elems
            with torch.cuda.device(dev):
                upd = symm.empty(n, dtype=torch.float32, device=dev)
                x = symm.empty(n, dtype=torch.float32, device=dev)
                upd.zero_()
                x.zero_()
                hu, hx = symm.rendezvous(upd, g), symm.rendezvous(x, g)
            mc_u = int(getattr(hu, "multicast_ptr", 0) or 0)
            mc_x = int(getattr(hx, "multicast_ptr", 0) or 0)
            use_mc = bool(mc_u and mc_x) and mode != "p2p"
            arr = ctypes.c_uint64 * self.world
            self.fused = dict(upd=upd, x=x, hu=hu, hx=hx, mc_u=mc_u if use_mc else 0, mc_x=mc_x if use_mc else 0,
                              x_peers=arr(*[int(v) for v in hx.buffer_ptrs]),
                              upd_peers=arr(*[int(v) for v in hu.buffer_ptrs]),
                              kind="multicast" if use_mc else "peer pointers")
            self.upd = upd
        except Exception as e:                     # noqa: BLE001 -- any failure means "not available here"
            if required:
                raise
            ok = 0
            self.fused = None
            self.fused_error = repr(e)
        # every rank must take the same path: a rank that fell back would sit in all_reduce while the
        # others wait in the symmetric-memory barrier
        try:
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=g)
            if int(flag.item()) == 0 and self.fused is not None:
                self.fused = None
                self.fused_error = "another rank could not set up symmetric memory"
                self.upd = self.plan.new_volume()
        except Exception:                          # noqa: BLE001 -- no working collective: nothing to agree on
            pass

    def new_iterate(self):
        """Zeroed resident volume for x.  With the fused collective it is THE symmetric buffer the peers
        write into (one per OSSARTState; a second call returns the same tensor, zeroed)."""
        if self.scheme == "gather":
            self.gsym["x"].zero_()
            return self.gsym["x"]
        if self.fused is not None:
            self.fused["x"].zero_()
            return self.fused["x"]
        return self.plan.new_volume()

    def _timed(self, name, fn):
        """Run fn(); with kernel_events enabled also bracket it with CUDA events under `name` (bench breakdown)."""
        if self.kernel_events is None:
            if _NVTX:                  # PAX_NVTX=1: named ranges around the phases of a subset step (nsys / ncu)
                torch.cuda.nvtx.range_push("pax:" + name)
                try:
                    return fn()
                finally:
                    torch.cuda.nvtx.range_pop()
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.phase_events.setdefault(name, []).append((e0, e1))
        return out

    def subset_step(self, x_res, s, lam, noneg=True, computel2=False):
        import torch.distributed as dist
        p = self.plan
        first, count = self.ranges[s]
        ss = self.sumsq if computel2 else None
        if self.scheme == "gather":
            return self._subset_step_gather(x_res, s, lam, noneg, ss)
        if count:
            ev = None
            if self.kernel_events is not None:     # bench: time the dominant kernel in place
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            p.ax_residual(x_res, self.b[first:first + count], first, count, self.resid,
                          self.w_variant, ss)
            if ev is not None:
                ev[1].record()
                self.kernel_events.append((ev[0], ev[1], count))
            self.launches += p.ax_launches(1)                  # (quad re-pack +) projector
        if self.world == 1:
            if count <= _lib.PAX_ATB_SART_MAX:
                self._timed("atb", lambda: p.atb(self.resid, first, count, x_res, _lib.PAX_ATB_SART,
                                                 self.inv_v[s], lam, noneg))
                self.launches += 1
            else:
                # a block larger than the fused kernel's angle table (TIGRE takes any blocksize, SIRT
                # included): backproject into a scratch update (STORE, then ADD chunks inside pax_atb)
                # and apply the update separately -- the path the multi-GPU branch below always takes
                if self.upd is None:
                    self.upd = p.new_volume()
                p.atb(self.resid, first, count, self.upd, _lib.PAX_ATB_STORE)
                p.sart_update(x_res, self.upd, self.inv_v[s], lam, noneg)
                self.launches += 1 + -(-count // _lib.PAX_ATB_SART_MAX)
            return
        if count:
            self._timed("atb", lambda: p.atb(self.resid, first, count, self.upd, _lib.PAX_ATB_STORE))
            self.launches += 1
        else:
            self.upd.zero_()
        f = self.fused
        if f is not None and x_res.data_ptr() == f["x"].data_ptr():
            # one kernel: reduce the partial updates over the ranks, update my slice, write it to every rank
            self._timed("barrier", lambda: f["hu"].barrier(channel=0))      # every rank's partial update is complete
            self._timed("update", lambda: _lib.check(p.lib.pax_sart_update_fused(
                p._h, _vp(x_res), _vp_int(f["mc_x"]), _vp_int(f["mc_u"]), f["x_peers"], f["upd_peers"],
                self.rank, self.world, _vp(self.inv_v[s]), float(lam), int(bool(noneg)), _cur_stream()),
                "pax_sart_update_fused"))
            self._timed("barrier", lambda: f["hx"].barrier(channel=0))      # every slice has landed everywhere
            self.launches += 1
            return
        self._timed("update", lambda: (dist.all_reduce(self.upd, group=self.group),
                                       p.sart_update(x_res, self.upd, self.inv_v[s], lam, noneg)))
        self.launches += 1

    def _subset_step_gather(self, x_res, s, lam, noneg, ss):
        p, f = self.plan, self.gsym
        if x_res.data_ptr() != f["x"].data_ptr():
            raise ValueError("the gather scheme iterates on the symmetric buffer new_iterate() returns")
        first, count = self.ranges[s]
        gfirst, glen = f["granges"][s]
        blk = self.blocks[s]
        A, R = self.grid
        off = len(blk) * (self.rank % A) // A              # my angle group's place among the subset's projections
        lib, npix = p.lib, f["npix"]
        if count:
            ev = None
            if self.kernel_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            # alone in my angle group: straight into my slice of the symmetric set; with column shares: into a
            # local buffer whose other columns stay zero
            dst = f["r"][off:off + count] if R == 1 else self.resid[:count]
            p.ax_residual(x_res, self.b[first:first + count], first, count, dst, self.w_variant, ss)
            if ev is not None:
                ev[1].record()
                self.kernel_events.append((ev[0], ev[1], count))
            self.launches += p.ax_launches(1)
            # my residual projections into every rank's copy of the subset's set (added, when the ranks of a
            # group each hold some of a projection's columns)
            if R == 1:
                self._timed("gather", lambda: _lib.check(lib.pax_bcast_chunks(
                    _vp(f["r"]), _vp_int(f["mc_r"]), f["r_peers"], self.world, npix * 4, off, 1, count,
                    _cur_stream()), "pax_bcast_chunks"))
            else:
                self._timed("gather", lambda: _lib.check(lib.pax_bcast_chunks2(
                    _vp(self.resid), 0, _vp_int(f["mc_r"]), f["r_peers"], self.world, npix * 4, off, 1, count, 1,
                    _cur_stream()), "pax_bcast_chunks2"))
            self.launches += 1
        self._timed("barrier", lambda: f["hr"].barrier(channel=0))          # every rank's residuals have landed
        # all of the subset's angles into MY rows of the volume, SART update in the epilogue (as on one GPU) ...
        self._timed("atb", lambda: self.plan_full.atb(f["r"], gfirst, glen, x_res, _lib.PAX_ATB_SART, self.inv_v[s],
                                                      lam, noneg, rows=(self.rank, self.world)))
        # ... and those rows to every rank
        self._timed("update", lambda: _lib.check(lib.pax_bcast_xrows(
            self.plan_full._h, _vp(x_res), _vp_int(f["mc_x"]), f["x_peers"], self.world, _vp(self.inv_v[s]),
            self.rank, self.world, _cur_stream()), "pax_bcast_xrows"))
        if R > 1:
            f["r"][:glen].zero_()                          # the next subset's residuals are added into zeros
        self._timed("barrier", lambda: f["hx"].barrier(channel=0))          # every row has landed everywhere
        self.launches += 2

    def iteration(self, x_res, lam, noneg=True, computel2=False):
        for s in range(len(self.blocks)):
            self.subset_step(x_res, s, lam, noneg, computel2)


def _v_variant_b(geo, angles, blocks):
    """2021-Python analytic V (SURVEY.md A.5 variant B); host NumPy, tiny."""
    g, rows = geo.pack(angles)
    ny, nx = int(g[1]), int(g[2])
    dv = np.asarray(geo.dVoxel, float)
    sv = np.asarray(geo.sVoxel, float)
    x = -sv[2] / 2 + dv[2] / 2 + dv[2] * np.arange(nx) + rows[0, 5]
    y = -sv[1] / 2 + dv[1] / 2 + dv[1] * np.arange(ny) + rows[0, 4]
    yy, xx = np.meshgrid(y, x, indexing="ij")
    out = np.zeros((len(blocks), ny, nx), dtype=np.float32)
    for s, idx in enumerate(blocks):
        acc = np.zeros((ny, nx))
        for i in idx:
            A = rows[i, 0] + np.pi / 2
            acc += (rows[i, 2] / (rows[i, 2] + yy * np.sin(-A) - xx * np.cos(-A))) ** 2
        out[s] = acc
    return out


# W / V definitions ossart() uses when the caller names none (the reference's TIGRE checkout is
# un-pinned and the definitions changed between builds).  'matlab' / 'A' -- current upstream's, and the pair a
# half-fan scan needs -- unless set_default_variants() chose otherwise: compat.install() selects the mid-2021
# Python package's ('python2021' / 'B'), the build the reference script was written against.
_DEFAULT_VARIANTS = {"w_variant": "matlab", "v_variant": "A"}


def set_default_variants(w_variant=None, v_variant=None):
    """Choose the W / V definitions of ossart() calls that do not name them; returns the previous pair."""
    prev = (_DEFAULT_VARIANTS["w_variant"], _DEFAULT_VARIANTS["v_variant"])
    if w_variant is not None:
        if w_variant not in ("matlab", "python2021"):
            raise ValueError(f"unknown W variant {w_variant!r}")
        _DEFAULT_VARIANTS["w_variant"] = w_variant
    if v_variant is not None:
        if v_variant not in ("A", "B"):
            raise ValueError(f"unknown V variant {v_variant!r}")
        _DEFAULT_VARIANTS["v_variant"] = v_variant
    return prev


def ossart(proj, geo, angles, niter, **kwargs):
    """OS-SART with TIGRE's signature and defaults (A.6).

    kwargs: blocksize=20, lmbda=1, lmbda_red=0.99, OrderStrategy=None, init=None,
    noneg=True, verbose=False, computel2=False, Quameasopts=None (a subset of RMSE / CC / MSSIM / UQI, measured
    between the iterate before and after every iteration: utilities/quality.py), gpuids=None; build extras: w_variant,
    v_variant (default: set_default_variants), group / distributed / shard (sharding over a torch.distributed group; dist.py),
    fused_update (the multi-GPU update as one kernel over symmetric memory; include/pax.h).
    Returns the (Z,Y,X) float32 volume (NumPy in -> NumPy out); with computel2 also the
    per-iteration ||b - Ax||_2.
    """
    blocksize = kwargs.pop("blocksize", 20)
    lmbda = float(kwargs.pop("lmbda", 1.0))
    lmbda_red = float(kwargs.pop("lmbda_red", 0.99))
    order = kwargs.pop("OrderStrategy", None)
    init = kwargs.pop("init", None)
    noneg = kwargs.pop("noneg", True)
    verbose = kwargs.pop("verbose", False)
    computel2 = kwargs.pop("computel2", False)
    gpuids = kwargs.pop("gpuids", None)
    quameas = kwargs.pop("Quameasopts", None)
    if quameas is not None:
        from ..utilities.quality import check_opts
        quameas = check_opts(quameas)
    state_kw = {k: kwargs.pop(k) for k in ("w_variant", "v_variant", "group", "distributed", "shard",
                                           "fused_update") if k in kwargs}
    for k, v in _DEFAULT_VARIANTS.items():
        if state_kw.get(k) is None:
            state_kw[k] = v
    if kwargs:
        raise TypeError(f"ossart: unexpected keyword arguments {sorted(kwargs)}")
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    import torch.distributed as dist

    nang = np.asarray(angles).shape[0]
    nv, nu = (int(v) for v in geo.nDetector)
    from ..projector import _multi
    devs = _multi(gpuids)
    if devs is not None and not any(k in state_kw for k in ("group", "distributed")):
        # several devices in ONE process (what TIGRE does with such a gpuids; reference :31,:103)
        if isinstance(proj, np.ndarray) and proj.dtype != np.float32:
            raise TypeError(f"proj must be float32, got {proj.dtype}")
        if isinstance(proj, torch.Tensor) and proj.dtype != torch.float32:
            raise TypeError(f"proj must be float32, got {proj.dtype}")
        if tuple(proj.shape) != (nang, nv, nu):
            raise ValueError(f"proj shape {tuple(proj.shape)} != (N, *geo.nDetector) {(nang, nv, nu)}")
        if isinstance(init, str):
            raise NotImplementedError(f"init={init!r} is not on this path (reference uses the default, zeros)")
        if quameas is not None:
            raise NotImplementedError("Quameasopts with several devices in one process: pass a single device")
        from ..multidev import MultiDevice
        md = MultiDevice(geo, angles, devs, blocksize=blocksize, OrderStrategy=order)
        x0 = None if init is None else _to_cuda(init, md.devices[0], "init")[0]
        res = md.ossart(proj, niter, lmbda, lmbda_red, noneg, x0, state_kw["w_variant"],
                        state_kw["v_variant"], computel2, verbose)
        was_np = isinstance(proj, np.ndarray)
        if computel2:
            return (res[0].cpu().numpy() if was_np else res[0]), res[1]
        return res.cpu().numpy() if was_np else res
    dev = torch.device("cuda", _device_from_gpuids(gpuids))
    b, was_np = _to_cuda(proj, dev, "proj")
    if tuple(b.shape) != (nang, nv, nu):
        raise ValueError(f"proj shape {tuple(b.shape)} != (N, *geo.nDetector) {(nang, nv, nu)}")
    st = OSSARTState(b, geo, angles, blocksize, order, **state_kw)
    p = st.plan
    if init is None:
        x = st.new_iterate()
    elif isinstance(init, str):
        raise NotImplementedError(f"init={init!r} is not on this path (reference uses the default, zeros)")
    else:
        x0, _ = _to_cuda(init, b.device, "init")
        x = p.pack(x0, out=st.new_iterate())
    lam = lmbda
    l2, lq = [], []
    for it in range(int(niter)):
        if computel2:
            st.sumsq.zero_()
        prev = p.unpack(x) if quameas is not None else None      # Measure_Quality(res_prev, res) per iteration
        st.iteration(x, lam, noneg, computel2)
        if quameas is not None:
            from ..utilities.quality import Measure_Quality
            lq.append(Measure_Quality(prev, p.unpack(x), quameas))
        if computel2:
            ss = st.sumsq.clone()
            if st.world > 1:
                dist.all_reduce(ss, group=st.group)
            l2.append(float(ss.sqrt().item()))
        if verbose and st.rank == 0:
            print(f"OS-SART iteration {it + 1}/{niter}, lambda={lam:.4f}")
        lam *= lmbda_red
    out = p.unpack(x)
    p.check_fault()                    # the fan-plane projector's self-check (csrc/ax_fan.cu)
    out = out.cpu().numpy() if was_np else out
    extra = ([np.array(l2)] if computel2 else []) + ([np.array(lq).T] if quameas is not None else [])
    return (out, *extra) if extra else out        # TIGRE: (res, l2) / (res, quality[measure, iteration])


OS_SART = ossart
