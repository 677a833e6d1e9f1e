"""CUDA path vs the float64 oracle on the same seeded inputs, through the C ABI (libpax.so).

Tolerances are BASELINE.json's: projections relative L2 <= 1e-4; volumes after the same
iteration count mean |dHU| <= 1 and max |dHU| <= 5 with HU = 4095 x - 1000
(statistics/compute_sCT_stats_psCBCT.py:115-116 of the reference).
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

PROJ_TOL = 1e-4
HU = 4095.0


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _cases():
    from physics_arx_b200.phantom import make_geometry
    out = []
    # isotropic, detector larger than the object
    out.append(("iso", make_geometry((20, 24, 28), (8.0, 8.0, 8.0), (40, 48), (10.0, 10.0)),
                np.linspace(0, 2 * np.pi, 7, endpoint=False)))
    # anisotropic planning-CT-like spacing, half-fan offset, truncated in z and in plane
    out.append(("halffan", make_geometry((12, 40, 40), (20.0, 6.0, 6.0), (36, 64), (5.0, 3.0),
                                          off_detector=(0.0, -60.0)),
                np.linspace(0, 2 * np.pi, 9)))
    # odd sizes, offsets, short detector (rows < 32 exercises the 8-row warp shape)
    g = make_geometry((9, 17, 13), (7.0, 5.0, 9.0), (11, 21), (14.0, 9.0), off_detector=(6.0, 11.0))
    g.offOrigin = np.array((4.0, -7.0, 9.0))
    out.append(("odd", g, np.array([0.1, 1.3, 2.9, 4.4, 5.9])))
    return out


@pytest.mark.parametrize("name,geo,angles", _cases(), ids=[c[0] for c in _cases()])
def test_ax_matches_oracle(oracle, name, geo, angles):
    import physics_arx_b200 as pax
    rng = np.random.default_rng(1)
    vol = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
    ref = oracle.ax(vol, geo, angles)
    got = pax.Ax(vol, geo, angles, "interpolated")
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert rel_l2(got, ref) <= PROJ_TOL


@pytest.mark.parametrize("name,geo,angles", _cases(), ids=[c[0] for c in _cases()])
def test_atb_matches_oracle(oracle, name, geo, angles):
    import physics_arx_b200 as pax
    rng = np.random.default_rng(2)
    proj = rng.random((len(angles),) + tuple(int(v) for v in geo.nDetector)).astype(np.float32)
    ref = oracle.atb(proj, geo, angles)
    got = pax.Atb(proj, geo, angles, "FDK")
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert rel_l2(got, ref) <= 1e-5


@pytest.mark.parametrize("name,geo,angles", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("variant", ["matlab", "python2021"])
def test_w_v_match_oracle(oracle, name, geo, angles, variant):
    torch = _torch()
    from physics_arx_b200.projector import Plan
    plan = Plan(geo, angles)
    w = plan.compute_w(w_variant=variant).cpu().numpy()
    wref = oracle.compute_w(geo, angles, variant)
    assert np.array_equal(w == 0, wref == 0) or np.mean((w == 0) != (wref == 0)) < 1e-3
    assert rel_l2(w, wref) <= 1e-5
    blocks = [np.arange(0, len(angles) // 2), np.arange(len(angles) // 2, len(angles))]
    vref = oracle.compute_v(geo, angles, blocks, "A")
    for s, blk in enumerate(blocks):
        v = plan.compute_v(int(blk[0]), len(blk)).cpu().numpy()
        assert rel_l2(v, vref[s]) <= 1e-5


def test_scatter_epilogue_is_linear(oracle):
    """Ax(ct) + s*Ax(art) in one two-channel pass == oracle Ax(ct + s*art) (AddImages.cxx:57-62)."""
    import physics_arx_b200 as pax
    name, geo, angles = _cases()[1]
    rng = np.random.default_rng(3)
    shape = tuple(int(v) for v in geo.nVoxel)
    ct = rng.random(shape).astype(np.float32)
    art = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    ref = oracle.ax(ct.astype(np.float64) + 0.7 * art.astype(np.float64), geo, angles)
    got = pax.Ax(ct, geo, angles, "interpolated", scatter=(art, 0.7))
    assert rel_l2(got, ref) <= PROJ_TOL


def test_ctnoise_matches_oracle(oracle):
    from physics_arx_b200.utilities import CTnoise
    rng = np.random.default_rng(4)
    p = (rng.random((5, 16, 24)) * 300).astype(np.float32)
    for I0, sig in ((1e8, 4.0), (1e5, 0.5), (200.0, 2.0)):
        ref = oracle.ctnoise_add(p, Poisson=I0, Gaussian=np.array([0.0, sig]), seed=7)
        got = CTnoise.add(p, Poisson=I0, Gaussian=np.array([0.0, sig]), seed=7)
        assert got.dtype == np.float32
        # same Philox words on both sides; libm differences can flip a rounding tie
        assert rel_l2(got, ref) <= 1e-5, (I0, rel_l2(got, ref))
    with pytest.raises(ValueError):
        CTnoise.add(p, Poisson=np.array([1.0, 2.0]))
    with pytest.raises(ValueError):
        CTnoise.add(p, Poisson=1e5, Gaussian=np.array([0.0, 1.0, 2.0]))


def test_ossart_matches_oracle(oracle):
    import physics_arx_b200 as pax
    from physics_arx_b200.phantom import lung_phantom, make_geometry
    geo = make_geometry((16, 32, 32), (10.0, 6.0, 6.0), (40, 64), (6.0, 5.0))
    angles = np.linspace(0, 2 * np.pi, 24)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    b = oracle.ax(vol, geo, angles).astype(np.float32)
    ref, l2ref = oracle.ossart(b, geo, angles, 3, blocksize=5, return_l2=True)
    got, l2 = pax.algorithms.ossart(b, geo, angles, 3, blocksize=5, computel2=True)
    d = np.abs(got.astype(np.float64) - ref) * HU
    assert d.mean() <= 1.0 and d.max() <= 5.0, (d.mean(), d.max())
    assert np.allclose(l2, l2ref, rtol=1e-3)
    # residual norm decreases over the first iterations on noise-free data (known answer vii)
    assert l2[0] > l2[1] > l2[2]


def test_error_behaviour():
    import physics_arx_b200 as pax
    _, geo, angles = _cases()[0]
    shape = tuple(int(v) for v in geo.nVoxel)
    with pytest.raises(TypeError):
        pax.Ax(np.zeros(shape, dtype=np.float64), geo, angles, "interpolated")
    with pytest.raises(ValueError):
        pax.Ax(np.zeros((3, 3, 3), dtype=np.float32), geo, angles, "interpolated")
    with pytest.raises(ValueError):
        pax.Atb(np.zeros((2, 3, 3), dtype=np.float32), geo, angles, "FDK")
    with pytest.raises(ValueError):
        pax.Ax(np.zeros(shape, dtype=np.float32), geo, angles, "Siddon")


def test_torch_tensor_roundtrip_stays_on_device():
    torch = _torch()
    import physics_arx_b200 as pax
    _, geo, angles = _cases()[0]
    shape = tuple(int(v) for v in geo.nVoxel)
    vol = torch.rand(shape, device="cuda", dtype=torch.float32)
    p = pax.Ax(vol, geo, angles, "interpolated")
    assert isinstance(p, torch.Tensor) and p.is_cuda
    x = pax.Atb(p, geo, angles, "FDK")
    assert isinstance(x, torch.Tensor) and x.is_cuda and tuple(x.shape) == shape


def test_pipeline_fused_max_and_image_domain_scatter(oracle):
    """SCBCTPipeline.simulate: max(p) from the Ax epilogue == max of the stored projections, and
    pack+axpy+Ax == oracle Ax(ct + s*art) (the reference's AddImages -> Ax order)."""
    torch = _torch()
    from physics_arx_b200.phantom import artifact_volume, lung_phantom
    from physics_arx_b200.pipeline import SCBCTPipeline
    _, geo, angles = _cases()[1]
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    pipe = SCBCTPipeline(geo, angles, niter=1, blocksize=4, I0=1e8, sigma=0.0, scatter_scale=0.7)
    p = pipe.plan
    p.pack(torch.from_numpy(ct).cuda(), out=pipe.ct_res)
    p.pack(torch.from_numpy(art).cuda(), out=pipe.art_res)
    p.axpy(pipe.ct_res, pipe.art_res, 0.7)
    m = torch.zeros(1, device="cuda")
    proj = p.ax(pipe.ct_res, maxout=m)
    assert float(m.item()) == float(proj.max().item())
    ref = oracle.ax(ct.astype(np.float64) + 0.7 * art.astype(np.float64), geo, angles)
    assert rel_l2(proj.cpu().numpy(), ref) <= PROJ_TOL


def test_plain_and_quad_gather_backends_agree(oracle):
    """The workspace-less kernel (PAX_AX_KERNEL=plain) is the same arithmetic on the same lattice."""
    torch = _torch()
    from physics_arx_b200.projector import Plan
    _, geo, angles = _cases()[1]
    rng = np.random.default_rng(5)
    vol = torch.from_numpy(rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)).cuda()
    plan = Plan(geo, angles)
    res = plan.pack(vol)
    a = plan.ax(res).cpu().numpy()
    plan.use_workspace = False
    b = plan.ax(res).cpu().numpy()
    assert rel_l2(a, b) <= 2e-6
    assert rel_l2(a, oracle.ax(vol.cpu().numpy(), geo, angles)) <= PROJ_TOL


# ----------------------------------------------------------------------- fan-plane projector
def _fan_cases():
    """Geometries for the FAN kernel (csrc/ax_fan.cu): fine detector rows over thick slices so that
    runs of equal n are long (the C2 regime, scaled down), plus the generic cases, where runs are a
    few rows long and the kernel is merely correct."""
    from physics_arx_b200.phantom import make_geometry
    out = list(_cases()[:2])
    # C2-like: 2.5 mm slices, 0.5 mm rows, half-fan, z-truncated, several row tiles and runs
    out.append(("c2like", make_geometry((24, 48, 48), (2.5, 1.0, 1.0), (300, 72), (0.5, 1.6),
                                         off_detector=(0.0, -20.0)),
                np.linspace(0, 2 * np.pi, 5, endpoint=False) + 0.3))
    # vertical detector offset (rays all tilted one way), volume offset, odd counts
    g = make_geometry((19, 33, 29), (3.0, 1.5, 2.0), (157, 41), (0.7, 2.0), off_detector=(25.0, 5.0))
    g.offOrigin = np.array((6.0, -4.0, 3.0))
    out.append(("tilted", g, np.array([0.0, 0.9, 2.2, 3.9, 5.5])))
    return out


@pytest.mark.parametrize("name,geo,angles", _fan_cases(), ids=[c[0] for c in _fan_cases()])
def test_fan_projector_matches_oracle(oracle, name, geo, angles):
    torch = _torch()
    from physics_arx_b200.projector import Plan
    rng = np.random.default_rng(11)
    vol = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
    plan = Plan(geo, angles, projector="fan")
    assert plan.projector == "fan" and plan.fan_info()["rows_per_tile"] >= 1
    res = plan.pack(torch.from_numpy(vol).cuda())
    ns = torch.zeros(1, dtype=torch.int64, device="cuda")
    m = torch.zeros(1, device="cuda")
    got = plan.ax(res, nsamples=ns, maxout=m)
    ref, nref = oracle.ax(vol, geo, angles, return_counts=True)
    assert rel_l2(got.cpu().numpy(), ref) <= PROJ_TOL, rel_l2(got.cpu().numpy(), ref)
    assert rel_l2(got.cpu().numpy(), ref) <= 2e-6            # it is the same sum, re-associated
    assert plan.fault() == 0
    assert float(m.item()) == float(got.max().item())
    assert abs(int(ns.item()) - int(nref.sum())) <= 1e-5 * nref.sum() + 2 * got.numel() / len(angles)
    # and the ray kernel on the same plan gives the same projections
    plan.set_projector("ray")
    ray = plan.ax(res)
    assert rel_l2(got.cpu().numpy(), ray.cpu().numpy()) <= 2e-6


def test_fan_projector_residual_and_two_channel(oracle):
    torch = _torch()
    from physics_arx_b200.projector import Plan
    name, geo, angles = _fan_cases()[2]
    rng = np.random.default_rng(12)
    shape = tuple(int(v) for v in geo.nVoxel)
    ct = rng.random(shape).astype(np.float32)
    art = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    plan = Plan(geo, angles, projector="fan")
    r0, r1 = plan.pack(torch.from_numpy(ct).cuda()), plan.pack(torch.from_numpy(art).cuda())
    ref = oracle.ax(ct.astype(np.float64) + 0.7 * art.astype(np.float64), geo, angles)
    got = plan.ax(r0, res1=r1, a0=1.0, a1=0.7)
    assert rel_l2(got.cpu().numpy(), ref) <= 2e-6
    p1 = torch.empty_like(got)
    p0 = plan.ax(r0, res1=r1, out1=p1)
    assert rel_l2((p0 + 0.7 * p1).cpu().numpy(), ref) <= 2e-6
    # residual epilogue: W*(b - Ax(x)) and the running sum of squares, against the ray kernel
    b = (got * 1.1).contiguous()
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    out = torch.empty_like(got)
    plan.ax_residual(r0, b, 0, len(angles), out, sumsq=ss)
    w = plan.compute_w()
    want = w * (b - plan.ax(r0))
    assert rel_l2(out.cpu().numpy(), want.cpu().numpy()) <= 1e-6
    assert abs(float(ss.item()) - float(((b - plan.ax(r0)) ** 2).sum().item())) <= 1e-5 * float(ss.item())


def test_ossart_with_fan_projector_matches_oracle(oracle):
    import physics_arx_b200 as pax
    from physics_arx_b200.phantom import lung_phantom, make_geometry
    import os
    geo = make_geometry((20, 40, 40), (2.5, 1.0, 1.0), (200, 64), (0.5, 1.5), off_detector=(0.0, -15.0))
    angles = np.linspace(0, 2 * np.pi, 24)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    b = oracle.ax(vol, geo, angles).astype(np.float32)
    ref = oracle.ossart(b, geo, angles, 3, blocksize=5)
    os.environ["PAX_AX_KERNEL"] = "fan"
    try:
        got = pax.algorithms.ossart(b, geo, angles, 3, blocksize=5)
    finally:
        del os.environ["PAX_AX_KERNEL"]
    d = np.abs(got.astype(np.float64) - ref) * HU
    assert d.mean() <= 1.0 and d.max() <= 5.0, (d.mean(), d.max())


def test_full_size_c2_projection_matches_oracle(oracle):
    """BASELINE.json's full C2 geometry (512x512x128, 1024x768 half-fan): three angles of the phantom's
    projections against the float64 oracle, for both projector kernels."""
    torch = _torch()
    from physics_arx_b200.phantom import config_c2, lung_phantom
    from physics_arx_b200.projector import Plan
    cfg = config_c2()
    geo = cfg["geo"]
    angles = cfg["angles"][[3, 250, 601]]
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    ref = oracle.ax(vol, geo, angles)
    plan = Plan(geo, angles)
    assert plan.projector == "fan"                     # what AUTO resolves to at C2
    res = plan.pack(torch.from_numpy(vol).cuda())
    for kind in ("fan", "ray"):
        plan.set_projector(kind)
        got = plan.ax(res).cpu().numpy()
        err = rel_l2(got, ref)
        print(kind, "rel L2 vs oracle", err, "max abs", np.abs(got - ref).max())
        assert plan.fault() == 0
        assert err <= PROJ_TOL, (kind, err)
        assert err <= 5e-6, (kind, err)


@pytest.mark.parametrize("seed", list(range(14)))
def test_fan_and_ray_kernels_agree_on_random_geometries(seed):
    """Fuzz: random volume / detector sizes, spacings, offsets, distances and angles; the fan-plane kernel
    (forced) against the ray-march kernel, which the float64 oracle pins on the fixed cases above."""
    torch = _torch()
    from physics_arx_b200 import _lib
    from physics_arx_b200.phantom import make_geometry
    from physics_arx_b200.projector import Plan
    rng = np.random.default_rng(1000 + seed)
    nvox = (int(rng.integers(6, 40)), int(rng.integers(12, 56)), int(rng.integers(12, 56)))
    dvox = (float(rng.uniform(1.0, 6.0)), float(rng.uniform(0.8, 3.0)), float(rng.uniform(0.8, 3.0)))
    ndet = (int(rng.integers(20, 400)), int(rng.integers(8, 40)))
    ext_z = nvox[0] * dvox[0]
    ddet = (float(rng.uniform(0.3, 1.6) * ext_z * 1.5 / ndet[0]), float(rng.uniform(1.0, 4.0)))
    DSO = float(rng.uniform(500, 1200))
    DSD = DSO + float(rng.uniform(300, 800))
    off = (float(rng.uniform(-0.3, 0.3) * ndet[0] * ddet[0]), float(rng.uniform(-0.4, 0.4) * ndet[1] * ddet[1]))
    geo = make_geometry(nvox, dvox, ndet, ddet, DSD=DSD, DSO=DSO, off_detector=off)
    geo.offOrigin = np.array((rng.uniform(-0.2, 0.2) * ext_z, rng.uniform(-10, 10), rng.uniform(-10, 10)))
    if seed % 3 == 0:
        geo.accuracy = float(rng.choice([0.25, 0.4, 1.0]))
    angles = np.sort(rng.uniform(0, 2 * np.pi, size=4))
    if seed % 4 == 1:
        angles[0] = 0.0
        angles[1] = np.pi / 2          # axis-aligned views
    plan = Plan(geo, angles)
    if plan.fan_info()["rows_per_tile"] < 1:
        with pytest.raises(ValueError):
            plan.set_projector("fan")  # the kernel refuses geometries it cannot tile
        return
    vol = torch.from_numpy(rng.random(nvox).astype(np.float32)).cuda()
    res = plan.pack(vol)
    plan.set_projector("ray")
    ns_r = torch.zeros(1, dtype=torch.int64, device="cuda")
    pr = plan.ax(res, nsamples=ns_r)
    plan.set_projector("fan")
    ns_f = torch.zeros(1, dtype=torch.int64, device="cuda")
    pf = plan.ax(res, nsamples=ns_f)
    assert plan.fault() == 0, plan.fan_info()
    den = float(pr.double().norm())
    assert den > 0
    err = float((pf - pr).double().norm()) / den
    assert err <= 2e-5, (err, plan.fan_info())
    assert float((pf - pr).abs().max()) <= 1e-3 * float(pr.max())
    assert abs(int(ns_f.item()) - int(ns_r.item())) <= 2e-4 * int(ns_r.item()) + 4 * pr.numel() // len(angles)


def test_c1_like_pipeline_matches_oracle_after_five_iterations(oracle):
    """BASELINE.json's C1 recipe at half size (64^3 @ 4 mm, 96 angles, 128x128 detector @ 4 mm, scatter +
    noise, 5-iteration OS-SART with blocksize 20): the reconstructed volume within the stated tolerance of
    the float64 oracle run through the same sequence."""
    _torch()
    from physics_arx_b200.phantom import artifact_volume, lung_phantom, make_geometry
    from physics_arx_b200.pipeline import SCBCTPipeline
    geo = make_geometry((64, 64, 64), (4.0, 4.0, 4.0), (128, 128), (4.0, 4.0))
    angles = np.linspace(0, 2 * np.pi, 96)
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    pipe = SCBCTPipeline(geo, angles, niter=5, blocksize=20, I0=1e8, sigma=4.0, scatter_scale=0.5, seed=5)
    got = pipe.run_host(ct, art).numpy()
    p = oracle.ax(ct.astype(np.float64) + 0.5 * art.astype(np.float64), geo, angles)
    p = oracle.ctnoise_add(p.astype(np.float32), Poisson=1e8, Gaussian=np.array([0.0, 4.0]), seed=5)
    ref = oracle.ossart(p.astype(np.float32), geo, angles, 5, blocksize=20)
    d = np.abs(got.astype(np.float64) - ref) * HU
    assert d.mean() <= 1.0 and d.max() <= 5.0, (d.mean(), d.max())
    assert d.mean() <= 0.01, d.mean()          # in practice three orders of magnitude inside the tolerance


@pytest.mark.parametrize("parts", [2, 3, 4])
def test_fan_projector_split_along_the_chords(oracle, parts, monkeypatch):
    """A small launch (a rank's few angles in a multi-GPU run) is cut into parts along the chords and summed by
    a finish kernel (include/pax.h: pax_ax_fan_scratch_bytes).  Forced here: same projections, same residual."""
    torch = _torch()
    from physics_arx_b200.projector import Plan
    name, geo, angles = _fan_cases()[2]
    rng = np.random.default_rng(13)
    vol = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
    plan = Plan(geo, angles, projector="fan")
    res = plan.pack(torch.from_numpy(vol).cuda())
    whole = plan.ax(res).cpu().numpy()
    monkeypatch.setenv("PAX_FAN_SPLIT", str(parts))
    m = torch.zeros(1, device="cuda")
    split = plan.ax(res, maxout=m)
    assert plan._work is not None and plan.fault() == 0
    assert rel_l2(split.cpu().numpy(), whole) <= 5e-7
    assert rel_l2(split.cpu().numpy(), oracle.ax(vol, geo, angles)) <= 2e-6
    assert float(m.item()) == float(split.max().item())
    b = (split * 1.1).contiguous()
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    out = torch.empty_like(split)
    plan.ax_residual(res, b, 0, len(angles), out, sumsq=ss)
    monkeypatch.delenv("PAX_FAN_SPLIT")
    want = torch.empty_like(split)
    plan.ax_residual(res, b, 0, len(angles), want)
    assert rel_l2(out.cpu().numpy(), want.cpu().numpy()) <= 1e-5
    assert abs(float(ss.item()) - float(((b - split) ** 2).sum().item())) <= 1e-4 * float(ss.item())


@pytest.mark.parametrize("rows", [72, 1100])
def test_fan_queued_launch_equals_grid_launch(oracle, rows, monkeypatch):
    """The fan projector hands its columns to the warps at run time from per-SM queues (csrc/ax_fan.cu,
    k_ax_fan_queued); PAX_FAN_QUEUED=0 at plan creation selects the plain grid launch of the same per-column walk.
    Both must give the same bits -- also with more detector rows than one warp walks (two row bands), where the
    queue's blocks carry the band as well."""
    torch = _torch()
    from physics_arx_b200.phantom import make_geometry
    from physics_arx_b200.projector import Plan
    geo = make_geometry((24, 48, 48), (2.5, 1.0, 1.0), (rows, 72), (150.0 / rows, 1.6), off_detector=(0.0, -20.0))
    angles = np.linspace(0, 2 * np.pi, 7, endpoint=False) + 0.3
    rng = np.random.default_rng(17)
    vol = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
    out = {}
    for queued in ("1", "0"):
        monkeypatch.setenv("PAX_FAN_QUEUED", queued)
        plan = Plan(geo, angles, projector="fan")
        res = plan.pack(torch.from_numpy(vol).cuda())
        ns = torch.zeros(1, dtype=torch.int64, device="cuda")
        p = plan.ax(res, nsamples=ns)
        b = (p * 1.05).contiguous()
        r = torch.empty_like(p)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.ax_residual(res, b, 0, len(angles), r, sumsq=ss)
        assert plan.fault() == 0
        out[queued] = (p.cpu().numpy(), r.cpu().numpy(), int(ns.item()), float(ss.item()))
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
    assert out["1"][2] == out["0"][2]
    assert abs(out["1"][3] - out["0"][3]) <= 1e-9 * out["0"][3]          # (atomic adds in a different order)
    if rows <= 300:
        assert rel_l2(out["1"][0], oracle.ax(vol, geo, angles)) <= 2e-6
