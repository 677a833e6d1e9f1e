"""NumPy model of the event-driven fan-plane projector (csrc/ax_fan.cu, round 2) for ONE detector column.

Design aid, not product and not oracle: it restates the kernel's arithmetic (fp32 where the kernel is fp32)
so that the event formulation, the crossing predicate and the effect of the epoch length on rounding can be
checked on the CPU against a float64 evaluation of the same Riemann sums (TIGRE's lattice).

    python tools/fan_event_model.py [--ep 256] [--u 300] [--angle 0.7] [--small]

Formulation.  For the rows of a column that share n (a segment) the in-plane sample positions are shared:
G_i(k) = bilinear_xy(slice k) at step i.  With the running sums  A_k(j) = sum_{i<=j} G_i(k)  and
C_k(j) = sum_{i<=j} (i - ref) G_i(k)  restarted at every epoch start, the line integral of row v is

    sum over epochs [ f_K(e1) ]  -  sum over crossings of a slice boundary k at step j
                                     dir * ( tau_k(v) * D2A_k(j-1) + s(v) * D2C_k(j-1) )

    f_k   = A_k + tau_k (A_{k+1} - A_k) + s (C_{k+1} - C_k)        (pair k at the epoch's last step)
    D2X_k = X_{k-1} - 2 X_k + X_{k+1},  tau_k(v) = z_ref(v) - k,  s(v) = dz per step,  dir = +1 upwards

i.e. a row touches the tables once per slice-boundary crossing and once per epoch, not once per step.
"Row v is above boundary k at step i" is DEFINED as  v >= ceil(fma(a_k, 1/i, -b))  with a_k = (k - sz) n / vz,
b = dz0 / vz: monotone in i and k by construction, so events enumerated per boundary (kernel: by the lanes
that own the slices) and the pair looked up at an epoch end always agree.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

f32 = np.float32
ZOFF = 4


def fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def frame(geo, theta):
    """padded index frame of one angle (plan.cu: pax_plan_create)"""
    nz, ny, nx = (int(v) for v in geo.nVoxel)
    dz, dy, dx = (float(v) for v in geo.dVoxel)
    nv, nu = (int(v) for v in geo.nDetector)
    dv, du = (float(v) for v in geo.dDetector)
    ov, ou = (float(v) for v in np.asarray(geo.offDetector, float))
    c, s = np.cos(theta), np.sin(theta)
    DSD, DSO = float(geo.DSD), float(geo.DSO)
    bx, by, bz = 0.5 * nx * dx - 0.5 * dx, 0.5 * ny * dy - 0.5 * dy, 0.5 * nz * dz - 0.5 * dz
    u0, v0 = (0.5 - 0.5 * nu) * du + ou, (0.5 - 0.5 * nv) * dv + ov
    Sx, Sy = DSO * c, DSO * s
    Px, Py, Pz = -(DSD - DSO) * c - u0 * s, -(DSD - DSO) * s + u0 * c, v0
    return dict(sx=(Sx + bx) / dx + 1, sy=(Sy + by) / dy + 1, sz=bz / dz + ZOFF,
                ox=(Px + bx) / dx + 1, oy=(Py + by) / dy + 1, oz=(Pz + bz) / dz + ZOFF,
                ux=-s * du / dx, uy=c * du / dy, vz=dv / dz, nv=nv, nu=nu,
                hix=nx + 2.0, hiy=ny + 2.0, loz=ZOFF - 1.0, hiz=nz + ZOFF * 1.0,
                d=(dx, dy, dz), acc=float(geo.accuracy))


def resident(vol):
    nz, ny, nx = vol.shape
    nzp = ((nz + ZOFF + 1) + 3) // 4 * 4
    r = np.zeros((ny + 2, nx + 2, nzp), f32)
    r[1:-1, 1:-1, ZOFF:ZOFF + nz] = np.transpose(vol, (1, 2, 0))
    return r


def slab(s, d, lo, hi, t0, t1):
    if d == 0.0:
        return (t0, t1) if lo <= s < hi else (1.0, 0.0)
    ta, tb = (lo - s) / d, (hi - s) / d
    return max(t0, min(ta, tb)), min(t1, max(ta, tb))


def column_reference(res, F, u):
    """float64 Riemann sums of every row of column u on TIGRE's lattice (unit step length)."""
    nv = F["nv"]
    dx, dy = F["ox"] + u * F["ux"] - F["sx"], F["oy"] + u * F["uy"] - F["sy"]
    out = np.zeros(nv)
    R = res.astype(np.float64)
    for v in range(nv):
        dz = F["oz"] + v * F["vz"] - F["sz"]
        n = int(np.ceil(np.sqrt(dx * dx + dy * dy + dz * dz) / F["acc"]))
        t0, t1 = slab(F["sx"], dx / n, 0.0, F["hix"], 0.0, float(n))
        t0, t1 = slab(F["sy"], dy / n, 0.0, F["hiy"], t0, t1)
        t0, t1 = slab(F["sz"], dz / n, F["loz"], F["hiz"], t0, t1)
        if t0 > t1:
            continue
        i = np.arange(int(np.ceil(t0)), int(np.floor(t1)) + 1, dtype=np.float64)
        if i.size == 0:
            continue
        x, y, z = F["sx"] + i * dx / n, F["sy"] + i * dy / n, F["sz"] + i * dz / n
        x0 = np.clip(np.floor(x).astype(int), 0, res.shape[1] - 2)
        y0 = np.clip(np.floor(y).astype(int), 0, res.shape[0] - 2)
        z0 = np.clip(np.floor(z).astype(int), 0, res.shape[2] - 2)
        fx, fy, fz = x - x0, y - y0, z - z0
        acc = 0.0
        for a, wy in ((0, 1 - fy), (1, fy)):
            for b, wx in ((0, 1 - fx), (1, fx)):
                for c, wz in ((0, 1 - fz), (1, fz)):
                    acc = acc + wy * wx * wz * R[y0 + a, x0 + b, z0 + c]
        out[v] = acc.sum()
    return out


def column_events(res, F, u, ep=256, stats=None):
    """the kernel's algorithm, fp32 tables, one column"""
    nv = F["nv"]
    nzp = res.shape[2]
    sx, sy, sz, vz = F["sx"], F["sy"], F["sz"], F["vz"]
    dx, dy = F["ox"] + u * F["ux"] - sx, F["oy"] + u * F["uy"] - sy
    dz0 = F["oz"] - sz
    rows = np.arange(nv)
    nrow = np.ceil(np.sqrt(dx * dx + dy * dy + (dz0 + rows * vz) ** 2) / F["acc"]).astype(int)
    starts = np.flatnonzero(np.r_[True, nrow[1:] != nrow[:-1]])
    ends = np.r_[starts[1:], nv]
    out = np.zeros(nv, f32)
    b32 = f32(dz0 / vz)
    szf = f32(sz)
    n_ev = 0
    for ra, rb in zip(starts, ends):
        n = int(nrow[ra])
        ddx, ddy = dx / n, dy / n
        t0, t1 = slab(sx, ddx, 0.0, F["hix"], 0.0, float(n))
        t0, t1 = slab(sy, ddy, 0.0, F["hiy"], t0, t1)
        if t0 > t1:
            continue
        I0, I1 = int(np.ceil(t0)), int(np.floor(t1))
        if I0 > I1:
            continue
        nvz = f32(n / vz)
        v = np.arange(ra, rb)
        s_v = ((dz0 + v * vz) / n)                                   # double; kernel: fp32 from double
        s32 = s_v.astype(f32)
        k_all = np.arange(nzp)
        a_k = ((k_all.astype(f32) - szf) * nvz).astype(f32)          # __fmul_rn(__fsub_rn(k, szf), nvz)
        for e0 in range(I0, I1 + 1, ep):
            e1 = min(e0 + ep - 1, I1)
            ii = np.arange(e0, e1 + 1)
            ref = e0 + ep // 2
            # in-plane positions: epoch origin split in double, steps added in fp32
            ex, ey = sx + e0 * ddx, sy + e0 * ddy
            exi, eyi = np.floor(ex), np.floor(ey)
            px = fma32((ii - e0).astype(f32), f32(ddx), f32(ex - exi))
            py = fma32((ii - e0).astype(f32), f32(ddy), f32(ey - eyi))
            cx = np.clip(int(exi) + np.floor(px).astype(int), 0, res.shape[1] - 2)
            cy = np.clip(int(eyi) + np.floor(py).astype(int), 0, res.shape[0] - 2)
            fx = np.clip(px - (cx - int(exi)).astype(f32), 0, 1).astype(f32)
            fy = np.clip(py - (cy - int(eyi)).astype(f32), 0, 1).astype(f32)
            gx, gy = f32(1) - fx, f32(1) - fy
            w00, w01, w10 = gx * gy, fx * gy, gx * fy
            w11 = f32(1) - w00 - w01 - w10
            G = (w00[:, None] * res[cy, cx] + w01[:, None] * res[cy, cx + 1]
                 + w10[:, None] * res[cy + 1, cx] + w11[:, None] * res[cy + 1, cx + 1]).astype(f32)
            A = np.cumsum(G, axis=0, dtype=f32)
            C = np.cumsum(((ii - ref).astype(f32))[:, None] * G, axis=0, dtype=f32)
            # row constants of the epoch
            zref = (sz + ref * s_v).astype(f32)                       # z of the row at the reference step
            # ---- crossings at steps e0+1 .. e1 (prefix through the step before)
            ri = (f32(1) / ii.astype(f32)).astype(f32)                 # __frcp_rn
            cV = np.ceil(fma32(a_k[None, :], ri[:, None], -b32)).astype(np.int64)    # [step, k]
            above = v[None, :, None] >= cV[:, None, :]                 # [step, row, k]
            flip = above[1:] != above[:-1]                             # crossing at step ii[t+1]
            tt, rr, kk = np.nonzero(flip)
            ok = (kk >= 1) & (kk <= nzp - 2)
            tt, rr, kk = tt[ok], rr[ok], kk[ok]
            d2a = A[tt, kk - 1] - f32(2) * A[tt, kk] + A[tt, kk + 1]
            d2c = C[tt, kk - 1] - f32(2) * C[tt, kk] + C[tt, kk + 1]
            direc = np.where(above[tt + 1, rr, kk], f32(1), f32(-1))
            tau = zref[rr] - kk.astype(f32)
            val = -direc * (tau * d2a + s32[rr] * d2c)
            np.add.at(out, v[rr], val.astype(f32))
            n_ev += tt.size
            # ---- close-out: pair of every row at e1 from the same predicate
            K = above[-1].sum(axis=1) - 1                              # largest k with above (k=0 always above)
            inside = (K >= 0) & (K <= nzp - 2)
            Kc = np.clip(K, 0, nzp - 2)
            tauK = zref - Kc.astype(f32)
            fK = A[-1, Kc] + tauK * (A[-1, Kc + 1] - A[-1, Kc]) + s32 * (C[-1, Kc + 1] - C[-1, Kc])
            out[ra:rb] += np.where(inside, fK, 0).astype(f32)
    if stats is not None:
        stats["events"] = n_ev
        stats["runs"] = len(starts)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ep", type=int, nargs="+", default=[128, 256, 4096])
    ap.add_argument("--u", type=int, nargs="+", default=[5, 300, 700, 1000])
    ap.add_argument("--angle", type=float, default=0.7)
    ap.add_argument("--small", action="store_true")
    args = ap.parse_args()
    from physics_arx_b200 import phantom
    cfg = phantom.config_c2()
    geo = cfg["geo"]
    if args.small:
        geo = phantom.make_geometry((32, 128, 128), (10.0, 3.9, 3.9), (192, 256), (1.552, 1.552),
                                    off_detector=(0.0, -160.0))
    vol = phantom.lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    res = resident(vol)
    F = frame(geo, args.angle)
    for u in args.u:
        u = min(u, F["nu"] - 1)
        ref = column_reference(res, F, u)
        for ep in args.ep:
            st = {}
            got = column_events(res, F, u, ep=ep, stats=st)
            err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
            print(f"u={u:5d} ep={ep:5d} runs={st['runs']} events/row={st['events'] / F['nv']:.1f} "
                  f"rel L2 {err:.3e}  max abs {np.abs(got - ref).max():.3e} (|ref| max {np.abs(ref).max():.3e})")


if __name__ == "__main__":
    main()
