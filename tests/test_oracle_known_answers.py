"""Pins for the oracle itself.  The reference ships no golden vectors for this path
(parity unpinned), so the oracle is anchored on answers that are true
of any correct cone-beam projector / backprojector / noise model, and on an independent
NumPy restatement (oracle/numpy_ref.py)."""
import numpy as np
import pytest

from conftest import rel_l2
from physics_arx_b200.phantom import make_geometry


def _geo(n=15, ndet=21, d=4.0, dd=8.0):
    return make_geometry((n, n, n), (d, d, d), (ndet, ndet), (dd, dd))


def test_c_oracle_equals_numpy_restatement(oracle):
    from oracle import numpy_ref
    geo = make_geometry((10, 14, 12), (9.0, 5.0, 7.0), (18, 22), (12.0, 9.0), off_detector=(5.0, -20.0))
    geo.offOrigin = np.array((3.0, -4.0, 6.0))
    ang = np.array([0.0, 0.7, 2.1, 3.9, 5.5])
    rng = np.random.default_rng(0)
    vol = rng.random((10, 14, 12))
    assert rel_l2(oracle.ax(vol, geo, ang), numpy_ref.ax(vol, geo, ang)) < 1e-12
    proj = rng.random((5, 18, 22))
    assert rel_l2(oracle.atb(proj, geo, ang), numpy_ref.atb(proj, geo, ang)) < 1e-12


def _chord_box(src, pix, half):
    d = pix - src
    t0, t1 = 0.0, 1.0
    for k in range(3):
        if d[k] == 0:
            if abs(src[k]) > half[k]:
                return 0.0
            continue
        a, b = sorted(((-half[k] - src[k]) / d[k], (half[k] - src[k]) / d[k]))
        t0, t1 = max(t0, a), min(t1, b)
    return max(t1 - t0, 0.0) * np.linalg.norm(d)


def test_uniform_box_gives_chord_length(oracle):
    """(i) Ax of a uniform box = c * chord +- one step * c (plus the half-voxel trilinear rim)."""
    geo = _geo(n=16, ndet=24, d=8.0, dd=10.0)
    c = 0.37
    vol = np.full((16, 16, 16), c)
    th = 0.4
    p = oracle.ax(vol, geo, [th])[0]
    src = np.array([1000 * np.cos(th), 1000 * np.sin(th), 0.0])
    half = np.array([64.0, 64.0, 64.0])
    for v, u in ((12, 12), (11, 9), (14, 13), (8, 15)):
        uu, vv = (u - 12 + 0.5) * 10.0, (v - 12 + 0.5) * 10.0
        pix = np.array([-500 * np.cos(th) - uu * np.sin(th), -500 * np.sin(th) + uu * np.cos(th), vv])
        chord = _chord_box(src, pix, half)
        assert chord > 100
        assert abs(p[v, u] - c * chord) <= c * 0.5 * 8.0 * 1.05 + 1e-9


def test_w_times_ax_ones_is_about_one(oracle):
    """(ii) W*Ax(ones) ~ 1/1.1 inside the FOV (the W box is 1.1x the volume in plane)."""
    geo = _geo(n=16, ndet=24, d=8.0, dd=6.0)
    ang = np.linspace(0, 2 * np.pi, 5, endpoint=False)
    r = oracle.compute_w(geo, ang) * oracle.ax(np.ones((16, 16, 16)), geo, ang)
    centre = r[:, 10:14, 10:14]
    assert np.all(centre > 0.80) and np.all(centre < 1.0)


def test_w_is_zero_where_ray_misses(oracle):
    geo = make_geometry((8, 8, 8), (4.0, 4.0, 4.0), (9, 41), (4.0, 20.0))
    w = oracle.compute_w(geo, [0.0])[0]
    assert np.all(w[:, 0] == 0) and np.all(w[:, -1] == 0) and np.all(w[:, 20] > 0)
    # chord through the box at the central ray: in-plane size 1.1 * 32 mm
    assert abs(1.0 / w[4, 20] - 35.2) < 0.05


def test_isocentre_point_projects_to_detector_centre(oracle):
    """(iii) a voxel at the isocentre peaks at the detector centre (+offDetector) for all angles."""
    geo = _geo()
    vol = np.zeros((15, 15, 15))
    vol[7, 7, 7] = 1.0
    ang = np.linspace(0, 2 * np.pi, 7, endpoint=False)
    p = oracle.ax(vol, geo, ang)
    for a in range(len(ang)):
        assert np.unravel_index(np.argmax(p[a]), p[a].shape) == (10, 10)
    geo.offDetector = np.array((16.0, -24.0))     # detector moves +2 px in v, -3 px in u
    p = oracle.ax(vol, geo, ang)
    for a in range(len(ang)):
        assert np.unravel_index(np.argmax(p[a]), p[a].shape) == (8, 13)


def test_offaxis_point_follows_the_sinogram_trace(oracle):
    """(iv) a point at (x0,0,0) lands at u = DSD * (-x0 sin th) / (DSO - x0 cos th)."""
    geo = _geo(n=15, ndet=41, d=4.0, dd=2.0)
    geo.DSD, geo.DSO = 1500.0, 1000.0
    vol = np.zeros((15, 15, 15))
    vol[7, 7, 12] = 1.0                       # x0 = +20 mm
    ang = np.linspace(0, 2 * np.pi, 9, endpoint=False)
    p = oracle.ax(vol, geo, ang)
    for a, th in enumerate(ang):
        u_mm = 1500.0 * (-20.0 * np.sin(th)) / (1000.0 - 20.0 * np.cos(th))
        u_px = u_mm / 2.0 + 41 / 2 - 0.5
        prof = p[a, 20]
        centroid = (prof * np.arange(41)).sum() / prof.sum()
        assert abs(centroid - u_px) < 0.6


def test_atb_of_ones_at_isocentre_counts_angles(oracle):
    """(v) Atb of all-ones projections at the isocentre voxel = N (weight 1 at x'=0)."""
    geo = _geo()
    ang = np.linspace(0, 2 * np.pi, 11, endpoint=False)
    x = oracle.atb(np.ones((11, 21, 21)), geo, ang)
    assert abs(x[7, 7, 7] - 11.0) < 1e-9
    assert np.allclose(x, oracle.atb(None, geo, ang), atol=1e-9)   # the analytic 'ones' branch


def test_atb_fdk_weight(oracle):
    geo = _geo()
    x = oracle.atb(np.ones((1, 21, 21)), geo, [0.0])
    # voxel (7,7,9) sits at x = +8 mm: weight (DSO/(DSO-x))^2
    assert abs(x[7, 7, 9] - (1000.0 / 992.0) ** 2) < 1e-9


def test_sphere_converges_to_analytic_chords(oracle):
    """(vi) Ax of a sphere against closed-form chord lengths; error shrinks with voxel size."""
    errs = []
    for n in (16, 32, 64):
        d = 128.0 / n
        geo = make_geometry((n, n, n), (d, d, d), (17, 17), (8.0, 8.0))
        z, y, x = np.meshgrid(*(((np.arange(n) - n / 2 + 0.5) * d,) * 3), indexing="ij")
        R = 40.0
        # anti-aliased sphere: fraction of the voxel inside, from a distance ramp
        vol = np.clip((R - np.sqrt(x * x + y * y + z * z)) / d + 0.5, 0, 1)
        p = oracle.ax(vol, geo, [0.3])[0]
        th = 0.3
        src = np.array([1000 * np.cos(th), 1000 * np.sin(th), 0.0])
        ref = np.zeros_like(p)
        for v in range(17):
            for u in range(17):
                uu, vv = (u - 8) * 8.0, (v - 8) * 8.0
                pix = np.array([-500 * np.cos(th) - uu * np.sin(th), -500 * np.sin(th) + uu * np.cos(th), vv])
                dirv = (pix - src) / np.linalg.norm(pix - src)
                bq = np.dot(src, dirv)
                disc = bq * bq - (np.dot(src, src) - R * R)
                ref[v, u] = 2 * np.sqrt(disc) if disc > 0 else 0.0
        errs.append(np.abs(p - ref).mean() / ref.max())
    assert errs[2] < errs[1] < errs[0] and errs[2] < 0.005


def test_ossart_residual_decreases_and_reconstructs(oracle):
    """(vii) ||b - Ax|| decreases over the first iterations on noise-free data."""
    from physics_arx_b200.phantom import lung_phantom
    geo = make_geometry((8, 24, 24), (12.0, 8.0, 8.0), (20, 48), (8.0, 6.0))
    ang = np.linspace(0, 2 * np.pi, 30, endpoint=False)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0).astype(np.float64)
    b = oracle.ax(vol, geo, ang)
    x, l2 = oracle.ossart(b, geo, ang, 4, blocksize=6, return_l2=True)
    assert np.all(np.diff(l2) < 0)
    assert np.all(x >= 0)
    inner = (slice(2, 6), slice(6, 18), slice(6, 18))
    assert np.abs(x[inner] - vol[inner]).mean() < 0.05


def test_ossart_half_fan_leaves_unseen_voxels_untouched(oracle):
    """Half-fan / truncated geometry: V == 0 => no update (SURVEY.md hard part 4)."""
    geo = make_geometry((6, 32, 32), (10.0, 10.0, 10.0), (12, 20), (6.0, 6.0), off_detector=(0.0, -80.0))
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    b = np.ones((16, 12, 20))
    x = oracle.ossart(b, geo, ang, 1, blocksize=4)
    assert np.all(np.isfinite(x)) and np.all(x >= 0)
    # the detector spans u in [-140,-20] mm: voxels within ~13 mm of the axis are in no fan
    assert np.all(x[:, 15:17, 15:17] == 0) and x.max() > 0


def test_philox_known_answer_vectors(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    assert oracle.philox_raw([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox_raw([0xffffffff] * 4, [0xffffffff] * 2) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox_raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    assert list(oracle.philox(0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


@pytest.mark.parametrize("I0,sigma", [(1e8, 4.0), (1e5, 0.5), (300.0, 1.0)])
def test_noise_moments(oracle, I0, sigma):
    """(viii) E[p'] ~ p and Var[p'] ~ m^2 (1/I + sigma^2/I^2), I = I0 exp(-p/m)."""
    n = 200000
    m = 250.0
    for pval in (50.0, 200.0):
        p = np.full(n, pval)
        out = oracle.ctnoise_add(p, Poisson=I0, Gaussian=np.array([0.0, sigma]), seed=3, maxval=m)
        I = I0 * np.exp(-pval / m)
        var = m * m * (1.0 / I + sigma ** 2 / I ** 2)
        assert abs(out.mean() - pval) < 6 * np.sqrt(var / n) + m * 0.6 / I
        assert abs(out.var() / var - 1.0) < 0.03


def test_noise_poisson_branches_match_numpy_statistics(oracle):
    """Counts recovered from p' follow Poisson(lambda): mean == var == lambda, both branches."""
    n = 200000
    for lam in (3.0, 80.0, 511.0, 513.0, 5000.0):
        p = np.zeros(n)
        out = oracle.ctnoise_add(p, Poisson=lam, Gaussian=np.array([0.0, 0.0]), seed=11, maxval=1.0)
        k = np.exp(-out) * lam
        assert abs(k.mean() / lam - 1) < 5 / np.sqrt(lam * n) + 1e-3
        assert abs(k.var() / lam - 1) < 0.03
        ref = np.random.default_rng(0).poisson(lam, n)
        assert abs(np.mean(k > lam + np.sqrt(lam)) - np.mean(ref > lam + np.sqrt(lam))) < 0.01


def test_noise_is_a_function_of_global_index_only(oracle):
    rng = np.random.default_rng(5)
    p = rng.random(1000) * 100
    full = oracle.ctnoise_add(p, Poisson=1e6, Gaussian=np.array([0.0, 2.0]), seed=9, maxval=100.0)
    a = oracle.ctnoise_add(p[:400], Poisson=1e6, Gaussian=np.array([0.0, 2.0]), seed=9, maxval=100.0)
    b = oracle.ctnoise_add(p[400:], Poisson=1e6, Gaussian=np.array([0.0, 2.0]), seed=9, maxval=100.0,
                           index0=400)
    assert np.array_equal(full, np.concatenate([a, b]))


def test_ax_sample_count_is_positive_and_bounded(oracle):
    geo = _geo()
    _, cnt = oracle.ax(np.ones((15, 15, 15)), geo, [0.0, 1.0], return_counts=True)
    assert np.all(cnt > 0)
    # every ray takes at most ~ (diagonal + 2) / accuracy samples inside the box
    assert np.all(cnt <= 21 * 21 * (np.sqrt(3) * 17 / 0.5 + 2))
