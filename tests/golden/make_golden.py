"""Generate tests/golden/*.npz from the float64 oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path; these are REGRESSION pins of
the oracle, not reference outputs.  Inputs are stored with the outputs so the GPU tests can run
on a box where neither this script nor /root/reference exists.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle as O  # noqa: E402
from physics_arx_b200.phantom import artifact_volume, lung_phantom, make_geometry  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def geo_params(geo):
    return dict(nVoxel=np.asarray(geo.nVoxel), dVoxel=np.asarray(geo.dVoxel, float),
                nDetector=np.asarray(geo.nDetector), dDetector=np.asarray(geo.dDetector, float),
                DSD=float(geo.DSD), DSO=float(geo.DSO), offDetector=np.asarray(geo.offDetector, float),
                offOrigin=np.asarray(geo.offOrigin, float))


def main():
    # case 1: half-fan, anisotropic, truncated -- Ax / Atb / W / V
    geo = make_geometry((10, 28, 28), (16.0, 7.0, 7.0), (24, 48), (6.0, 4.0), off_detector=(0.0, -40.0))
    angles = np.linspace(0, 2 * np.pi, 12)
    rng = np.random.default_rng(2021)
    vol = rng.random((10, 28, 28)).astype(np.float32)
    proj_in = rng.random((12, 24, 48)).astype(np.float32)
    blocks = [np.arange(0, 5), np.arange(5, 10), np.arange(10, 12)]
    np.savez_compressed(
        os.path.join(HERE, "halffan_ops.npz"), angles=angles, vol=vol, proj_in=proj_in,
        ax=O.ax(vol, geo, angles).astype(np.float64), atb=O.atb(proj_in, geo, angles),
        w_matlab=O.compute_w(geo, angles, "matlab"), w_py=O.compute_w(geo, angles, "python2021"),
        v_a=O.compute_v(geo, angles, blocks, "A"), **geo_params(geo))
    # case 2: the whole loop -- scatter add, Ax, noise, 3-iteration OS-SART
    geo = make_geometry((12, 32, 32), (12.0, 7.0, 7.0), (32, 64), (7.0, 5.0), off_detector=(0.0, -30.0))
    angles = np.linspace(0, 2 * np.pi, 20)
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    p = O.ax(ct.astype(np.float64) + 0.8 * art.astype(np.float64), geo, angles)
    noisy = O.ctnoise_add(p.astype(np.float32), Poisson=1e8, Gaussian=np.array([0.0, 4.0]), seed=11)
    rec = O.ossart(noisy.astype(np.float32), geo, angles, 3, blocksize=6)
    np.savez_compressed(os.path.join(HERE, "loop.npz"), angles=angles, ct=ct, art=art, scatter_scale=0.8,
                        I0=1e8, sigma=4.0, seed=11, niter=3, blocksize=6, proj=p, noisy=noisy, recon=rec,
                        **geo_params(geo))
    # case 3: noise stream words and samples
    pin = (rng.random(512) * 200).astype(np.float32)
    np.savez_compressed(
        os.path.join(HERE, "noise.npz"), p=pin,
        philox=np.stack([O.philox(7, i) for i in (0, 1, 2, 4294967296 + 5)]),
        hi=O.ctnoise_add(pin, Poisson=1e8, Gaussian=np.array([0.0, 4.0]), seed=7, maxval=200.0),
        mid=O.ctnoise_add(pin, Poisson=1e4, Gaussian=np.array([1.0, 2.0]), seed=7, maxval=200.0),
        lo=O.ctnoise_add(pin, Poisson=300.0, Gaussian=np.array([0.0, 0.5]), seed=7, maxval=200.0, index0=1000))
    # case 4: fan-plane regime (fine detector rows over thick slices, half-fan): Ax and a residual pass
    geo = make_geometry((20, 36, 36), (2.5, 1.0, 1.0), (224, 56), (0.45, 1.5), off_detector=(0.0, -12.0))
    angles = np.array([0.0, 0.7, 1.9, 3.3, 4.6, 5.8])
    vol = rng.random((20, 36, 36)).astype(np.float32)
    ax = O.ax(vol, geo, angles)
    b = (ax * (1.0 + 0.1 * rng.standard_normal(ax.shape))).astype(np.float32)
    w = O.compute_w(geo, angles, "matlab")
    np.savez_compressed(os.path.join(HERE, "fan_ops.npz"), angles=angles, vol=vol, ax=ax, b=b,
                        resid=w * (b.astype(np.float64) - ax), **geo_params(geo))
    # case 5: artifact induction (CBCT_pCT_Artifact_adding.sh steps 1-4, one variation)
    z, y, x = np.meshgrid(np.arange(10), np.arange(22), np.arange(26), indexing="ij")
    pct = np.round(-800 + 900 * np.exp(-((x - 13) ** 2 + (y - 11) ** 2) / 60.0)
                   + 40 * rng.standard_normal(x.shape)).astype(np.int16)
    z, y, x = np.meshgrid(np.arange(24), np.arange(30), np.arange(28), indexing="ij")
    cbct = np.round(-750 + 800 * np.exp(-((x - 14) ** 2 + (y - 16) ** 2) / 80.0)
                    + 60 * rng.standard_normal(x.shape) + 3.0 * z).astype(np.int16)
    pg = dict(spacing=(1.2, 1.2, 2.5), origin=(-15.0, -13.0, -12.0))
    cg = dict(spacing=(1.0, 1.0, 1.0), origin=(-13.0, -16.0, -11.0))
    m = O.index_map(pg["origin"], pg["spacing"], np.eye(3), cg["origin"], cg["spacing"], np.eye(3))
    w1 = O.resample_linear(cbct.astype(np.float64), m, pct.shape, cast="short")
    hist = O.plahe(w1, 5, 0.5, 0.5)
    added = pct.astype(np.float64) + hist
    nz = int(pct.shape[0] * 2.5 / 1.2)
    m2 = O.index_map(pg["origin"], (1.2, 1.2, 1.2), np.eye(3), pg["origin"], pg["spacing"], np.eye(3))
    rsc = O.rescale_intensity(O.resample_linear(added, m2, (nz,) + pct.shape[1:]), 0, 1)
    np.savez_compressed(os.path.join(HERE, "artifact.npz"), pct=pct, cbct=cbct, pct_spacing=pg["spacing"],
                        pct_origin=pg["origin"], cbct_spacing=cg["spacing"], cbct_origin=cg["origin"],
                        cbct_w1=w1.astype(np.int16), hist=hist.astype(np.float32), rsc=rsc.astype(np.float32))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
