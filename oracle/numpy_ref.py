"""Second, independent restatement of Ax / Atb in vectorised NumPy (float64).

TEST INFRASTRUCTURE ONLY.  Exists to cross-check oracle.c on tiny cases: it shares
no code with it (geometry is rebuilt here from the world-frame definitions of
SURVEY.md A.2 rather than from the index-space frame oracle.c uses).
"""
from __future__ import annotations

import numpy as np


def _trilinear_zero(vol, x, y, z):
    nz, ny, nx = vol.shape
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    z0 = np.floor(z).astype(np.int64)
    fx, fy, fz = x - x0, y - y0, z - z0
    out = np.zeros_like(x)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                ok = (xi >= 0) & (xi < nx) & (yi >= 0) & (yi < ny) & (zi >= 0) & (zi < nz)
                w = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                val = vol[np.clip(zi, 0, nz - 1), np.clip(yi, 0, ny - 1), np.clip(xi, 0, nx - 1)]
                out += np.where(ok, w * val, 0.0)
    return out


def ax(vol, geo, angles):
    """Ray-driven trilinear projector on TIGRE's lattice (A.3), all samples i=0..n."""
    vol = np.asarray(vol, dtype=np.float64)
    g, rows = geo.pack(angles)
    nz, ny, nx = vol.shape
    nv, nu = int(g[3]), int(g[4])
    dz, dy, dx, dv, du, acc = g[5], g[6], g[7], g[8], g[9], g[10]
    d = np.array([dx, dy, dz])
    out = np.zeros((rows.shape[0], nv, nu))
    uu = (np.arange(nu) - nu / 2 + 0.5) * du
    vv = (np.arange(nv) - nv / 2 + 0.5) * dv
    for ia, (th, DSD, DSO, oz, oy, ox, ov, ou) in enumerate(rows):
        c, s = np.cos(th), np.sin(th)
        src = np.array([DSO * c, DSO * s, 0.0])
        U, V = np.meshgrid(uu + ou, vv + ov, indexing="xy")  # (nv,nu)
        px = -(DSD - DSO) * c - U * s
        py = -(DSD - DSO) * s + U * c
        pz = V
        P = np.stack([px, py, pz], axis=-1)
        shift = np.array([-ox + nx * dx / 2 - dx / 2, -oy + ny * dy / 2 - dy / 2,
                          -oz + nz * dz / 2 - dz / 2])
        Sq = (src + shift) / d
        Pq = (P + shift) / d
        D = Pq - Sq
        L = np.linalg.norm(D, axis=-1)
        n = np.ceil(L / acc)
        step = D / n[..., None]
        nmax = int(n.max())
        acc_sum = np.zeros((nv, nu))
        for i in range(nmax + 1):
            q = Sq + i * step
            val = _trilinear_zero(vol, q[..., 0], q[..., 1], q[..., 2])
            acc_sum += np.where(i <= n, val, 0.0)
        out[ia] = acc_sum * np.linalg.norm(step * d, axis=-1)
    return out


def atb(proj, geo, angles):
    """Voxel-driven FDK-weighted backprojector (A.4)."""
    proj = np.asarray(proj, dtype=np.float64)
    g, rows = geo.pack(angles)
    nz, ny, nx = (int(v) for v in g[:3])
    nv, nu = int(g[3]), int(g[4])
    dz, dy, dx, dv, du = g[5], g[6], g[7], g[8], g[9]
    out = np.zeros((nz, ny, nx))
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for ia, (th, DSD, DSO, oz, oy, ox, ov, ou) in enumerate(rows):
        x = -nx * dx / 2 + (i + 0.5) * dx + ox
        y = -ny * dy / 2 + (j + 0.5) * dy + oy
        z = -nz * dz / 2 + (k + 0.5) * dz + oz
        c, s = np.cos(th), np.sin(th)
        xr = c * x + s * y
        yr = -s * x + c * y
        t = DSD / (DSO - xr)
        u = (yr * t - ou) / du + nu / 2 - 0.5
        v = (z * t - ov) / dv + nv / 2 - 0.5
        w = (DSO / (DSO - xr)) ** 2
        u0 = np.floor(u).astype(np.int64)
        v0 = np.floor(v).astype(np.int64)
        fu, fv = u - u0, v - v0
        val = np.zeros_like(u)
        p = proj[ia]
        for dvv in (0, 1):
            for duu in (0, 1):
                ui, vi = u0 + duu, v0 + dvv
                ok = (ui >= 0) & (ui < nu) & (vi >= 0) & (vi < nv)
                ww = (fu if duu else 1 - fu) * (fv if dvv else 1 - fv)
                val += np.where(ok, ww * p[np.clip(vi, 0, nv - 1), np.clip(ui, 0, nu - 1)], 0.0)
        out += w * val
    return out
