"""C4 (shared-projection variation fan-out) and C5 (per-angle respiratory phase) drivers vs the
oracle run the long way round (one Ax per variation / per angle)."""
import numpy as np
import pytest

from conftest import rel_l2

HU = 4095.0


def _geo():
    from physics_arx_b200.phantom import make_geometry
    return make_geometry((12, 28, 28), (12.0, 8.0, 8.0), (28, 56), (8.0, 5.0), off_detector=(0.0, -25.0))


def test_respiratory_phases_and_phase_volumes_host_logic():
    from physics_arx_b200.drivers import config_c4, phase_volumes, respiratory_phases
    ph = respiratory_phases(256, 10, 60.0, 4.0)
    assert ph.shape == (256,) and ph.min() == 0 and ph.max() == 9
    assert np.all(np.diff(ph) % 10 <= 1)                 # phases advance one at a time
    assert len(np.unique(ph[:18])) == 10                 # one breath = 4 s = ~17 projections
    vol = np.zeros((16, 4, 4), np.float32)
    vol[5] = 1.0
    pv = phase_volumes(vol, dz_mm=2.5, n_phases=10, amplitude_mm=15.0)
    assert np.array_equal(pv[0], vol)                    # phase 0 is the reference position
    assert abs(np.argmax(pv[5][:, 0, 0]) - 11) <= 0      # peak inhale: 15 mm = 6 voxels
    assert np.allclose(pv.sum(axis=1), 1.0, atol=1e-5)   # an interior slab keeps its mass
    c4 = config_c4()
    assert len(c4["scatter_scales"]) * len(c4["I0s"]) * len(c4["strides"]) == 32


@pytest.mark.gpu
def test_c4_variations_match_oracle(oracle):
    from physics_arx_b200.drivers import VariationBatch
    from physics_arx_b200.phantom import artifact_volume, lung_phantom
    geo = _geo()
    angles = np.linspace(0, 2 * np.pi, 20)
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    vb = VariationBatch(geo, angles, niter=2, blocksize=5, sigma=4.0)
    res = vb.run(ct, art, scatter_scales=(0.0, 1.5), I0s=(1e5, 1e8), strides=(1, 2), seed=10)
    assert len(res) == 8
    pct, part = oracle.ax(ct, geo, angles), oracle.ax(art, geo, angles)
    assert rel_l2(vb.p_ct.cpu().numpy(), pct) <= 1e-4
    assert rel_l2(vb.p_art.cpu().numpy(), part) <= 1e-4
    npix = int(np.prod(geo.nDetector))
    for r in res:
        s, i0, stride, i = r["scatter_scale"], r["I0"], r["stride"], r["index"]
        p = (pct + s * part)[::stride].astype(np.float32)
        m = float(p.max())
        noisy = np.stack([oracle.ctnoise_add(p[k], Poisson=i0, Gaussian=np.array([0.0, 4.0]), seed=10 + i,
                                             index0=k * stride * npix, maxval=m) for k in range(p.shape[0])])
        ref = oracle.ossart(noisy.astype(np.float32), geo, angles[::stride], 2, blocksize=5)
        hu = np.abs(r["volume"] - ref) * HU
        assert hu.mean() <= 1.0 and hu.max() <= 5.0, (r["index"], hu.mean(), hu.max())
    # rank/world split: disjoint cover
    idx = sorted(v["index"] for w in range(3) for v in vb.run(ct, art, (0.0, 1.5), (1e5, 1e8), (1, 2), seed=10,
                                                               rank=w, world=3))
    assert idx == list(range(8))


@pytest.mark.gpu
def test_c5_motion_projections_match_oracle(oracle):
    import physics_arx_b200 as pax
    from physics_arx_b200.drivers import motion_projections, phase_volumes, respiratory_phases
    from physics_arx_b200.phantom import lung_phantom
    geo = _geo()
    angles = np.linspace(0, 2 * np.pi, 24)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    pv = phase_volumes(vol, dz_mm=float(geo.dVoxel[0]), n_phases=4, amplitude_mm=15.0)
    ph = respiratory_phases(24, 4, scan_s=12.0, period_s=3.0)
    got = motion_projections(pv, ph, geo, angles).cpu().numpy()
    ref = np.stack([oracle.ax(pv[ph[k]], geo, angles[k:k + 1])[0] for k in range(24)])
    assert rel_l2(got, ref) <= 1e-4
    # motion-corrupted data reconstructs to something between the phases (sanity, not parity)
    rec = pax.algorithms.ossart(got, geo, angles, 3, blocksize=6)
    assert rec.shape == vol.shape and np.isfinite(rec).all() and rec.min() >= 0
    ref_rec = oracle.ossart(got, geo, angles, 3, blocksize=6)
    hu = np.abs(rec - ref_rec) * HU
    assert hu.mean() <= 1.0 and hu.max() <= 5.0
    with pytest.raises(ValueError):
        motion_projections(pv, ph[:-1], geo, angles)
