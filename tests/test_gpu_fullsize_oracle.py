"""The CUDA path against the float64 oracle at BASELINE.json's FULL C2 size (512x512x128 volume, 1024x768
half-fan detector), for the pieces of one OS-SART subset step beyond the plain Ax that
test_gpu_parity.test_full_size_c2_projection_matches_oracle pins: Atb, V, the residual epilogue
W.(b - Ax(x)) and one whole subset update (reference call site pCT_CBCT_Reconstruction.py:103; SURVEY.md
A.4-A.6).  Three angles keep the oracle in seconds.  Plus two size-independent sanity checks of SURVEY.md
section 4 (1)/(2): the <Ax x, y> / <x, Atb y> pairing and a hypothesis fuzz of the fan-plane kernel against
the oracle itself.  All through the C ABI (libpax.so)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

HU = 4095.0
PICK = [5, 230, 611]            # a shallow, an oblique and a steep view of the half-fan scan


@pytest.fixture(scope="module")
def c2(oracle):
    import torch
    from physics_arx_b200.phantom import config_c2, lung_phantom
    from physics_arx_b200.projector import Plan
    assert torch.cuda.is_available()
    cfg = config_c2()
    geo = cfg["geo"]
    angles = cfg["angles"][PICK]
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    plan = Plan(geo, angles)
    assert plan.projector == "fan"
    res = plan.pack(torch.from_numpy(vol).cuda())
    b_ref = oracle.ax(vol, geo, angles)                         # float64 projections of the phantom
    return dict(geo=geo, angles=angles, vol=vol, plan=plan, res=res, b_ref=b_ref, torch=torch, oracle=oracle)


def test_atb_matches_oracle_at_full_size(c2):
    torch, plan, oracle = c2["torch"], c2["plan"], c2["oracle"]
    rng = np.random.default_rng(21)
    proj = rng.random(c2["b_ref"].shape).astype(np.float32)
    ref = oracle.atb(proj, c2["geo"], c2["angles"])
    out = plan.new_volume()
    from physics_arx_b200 import _lib
    plan.atb(torch.from_numpy(proj).cuda(), 0, len(PICK), out, _lib.PAX_ATB_STORE)
    got = plan.unpack(out).cpu().numpy()
    # fp32 detector coordinates of ~10^3 pixels carry ~1e-4 pixel of rounding: 1.7e-5 measured at this size
    # (1e-5 holds on the small geometries of test_gpu_parity)
    assert rel_l2(got, ref) <= 5e-5, rel_l2(got, ref)


def test_v_matches_oracle_at_full_size(c2):
    plan, oracle = c2["plan"], c2["oracle"]
    ref = oracle.compute_v(c2["geo"], c2["angles"], [np.arange(len(PICK))], variant="A")[0]
    got = plan.compute_v(0, len(PICK)).cpu().numpy()
    assert rel_l2(got, ref) <= 1e-5, rel_l2(got, ref)
    # half-fan: a part of the plane is seen by none of the three views, and V must be exactly 0 there
    assert (ref == 0).any() and np.array_equal(got == 0, ref == 0)


def test_residual_epilogue_matches_oracle_at_full_size(c2):
    """W.(b - Ax(x)) with x = 0.9 phantom and b = Ax(phantom): the residual is 10 % of the projections, so
    a relative error e of Ax shows up as 9e here."""
    torch, plan, oracle = c2["torch"], c2["plan"], c2["oracle"]
    geo, angles = c2["geo"], c2["angles"]
    b = c2["b_ref"].astype(np.float32)
    x = c2["res"] * 0.9
    out = torch.empty((len(PICK), plan.nv, plan.nu), dtype=torch.float32, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.ax_residual(x, torch.from_numpy(b).cuda(), 0, len(PICK), out, sumsq=ss)
    w = oracle.compute_w(geo, angles, "matlab")
    r = b.astype(np.float64) - 0.9 * c2["b_ref"]                 # Ax is linear
    ref = w * r
    assert rel_l2(out.cpu().numpy(), ref) <= 1e-4, rel_l2(out.cpu().numpy(), ref)
    assert abs(float(ss.item()) - float((r * r).sum())) <= 1e-4 * float((r * r).sum())
    assert plan.fault() == 0


def test_one_subset_step_matches_oracle_at_full_size(c2):
    """x after Ax -> residual -> Atb -> update of ONE 3-angle subset, from x0 = 0.9 phantom, against the
    oracle's OS-SART run for one iteration on the same subset."""
    torch, oracle = c2["torch"], c2["oracle"]
    import physics_arx_b200 as pax
    geo, angles = c2["geo"], c2["angles"]
    b = c2["b_ref"].astype(np.float32)
    x0 = (0.9 * c2["vol"]).astype(np.float32)
    ref = oracle.ossart(b, geo, angles, 1, blocksize=len(PICK), init=x0)
    got = pax.algorithms.ossart(b, geo, angles, 1, blocksize=len(PICK), init=x0)
    d = np.abs(got.astype(np.float64) - ref) * HU
    print("one C2 subset step: mean |dHU|", d.mean(), "max", d.max())
    assert d.mean() <= 1.0 and d.max() <= 5.0, (d.mean(), d.max())
    assert d.mean() <= 0.01, d.mean()
    moved = np.abs(ref - x0).max() * HU
    assert moved > 10.0, moved                                   # the step did something


def test_ax_atb_pairing_sanity():
    """<Ax x, y> against <x, Atb y>: Atb 'FDK' is the weighted, not the matched, transpose (A.4), so the two
    inner products differ by a geometry factor; the CUDA pair must show the SAME factor as the oracle pair
    (and a factor of order one: a wrong axis or a missing weight shows up as orders of magnitude)."""
    import torch
    from oracle import cpu_oracle as oracle
    import physics_arx_b200 as pax
    from physics_arx_b200.phantom import make_geometry
    geo = make_geometry((24, 40, 40), (2.5, 1.0, 1.0), (160, 64), (0.5, 1.5), off_detector=(0.0, -10.0))
    angles = np.linspace(0, 2 * np.pi, 7, endpoint=False)
    rng = np.random.default_rng(31)
    x = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
    y = rng.random((len(angles),) + tuple(int(v) for v in geo.nDetector)).astype(np.float32)
    lhs = float((pax.Ax(x, geo, angles).astype(np.float64) * y).sum())
    rhs = float((x.astype(np.float64) * pax.Atb(y, geo, angles)).sum())
    lhs_o = float((oracle.ax(x, geo, angles) * y).sum())
    rhs_o = float((x * oracle.atb(y, geo, angles)).sum())
    assert abs(lhs / rhs - lhs_o / rhs_o) <= 1e-5 * abs(lhs_o / rhs_o), (lhs / rhs, lhs_o / rhs_o)
    assert 0.05 < lhs / rhs < 20.0, lhs / rhs
    assert torch.cuda.is_available()


def _fuzz_case(seed):
    from physics_arx_b200.phantom import make_geometry
    rng = np.random.default_rng(seed)
    nvox = (int(rng.integers(6, 28)), int(rng.integers(10, 36)), int(rng.integers(10, 36)))
    dvox = (float(rng.uniform(1.5, 5.0)), float(rng.uniform(0.8, 2.5)), float(rng.uniform(0.8, 2.5)))
    ndet = (int(rng.integers(24, 260)), int(rng.integers(6, 28)))
    ext_z = nvox[0] * dvox[0]
    ddet = (float(rng.uniform(0.4, 1.5) * ext_z * 1.5 / ndet[0]), float(rng.uniform(1.0, 4.0)))
    DSO = float(rng.uniform(500, 1200))
    DSD = DSO + float(rng.uniform(300, 800))
    off = (float(rng.uniform(-0.3, 0.3) * ndet[0] * ddet[0]), float(rng.uniform(-0.4, 0.4) * ndet[1] * ddet[1]))
    geo = make_geometry(nvox, dvox, ndet, ddet, DSD=DSD, DSO=DSO, off_detector=off)
    geo.offOrigin = np.array((rng.uniform(-0.2, 0.2) * ext_z, rng.uniform(-8, 8), rng.uniform(-8, 8)))
    geo.accuracy = float(rng.choice([0.25, 0.5, 0.5, 1.0]))
    angles = np.sort(rng.uniform(0, 2 * np.pi, size=3))
    return geo, angles, rng.random(nvox).astype(np.float32)


def test_fan_projector_matches_oracle_on_random_geometries(oracle):
    """Hypothesis-driven fuzz of the fan-plane kernel (forced) against the float64 ORACLE (the fixed-seed fuzz
    in test_gpu_parity compares it with the ray kernel only)."""
    import torch
    from hypothesis import HealthCheck, given, settings, strategies as stg
    from physics_arx_b200.projector import Plan
    assert torch.cuda.is_available()
    seen = []

    @settings(max_examples=20, deadline=None, derandomize=True,
              suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
    @given(stg.integers(min_value=0, max_value=10**6))
    def run(seed):
        geo, angles, vol = _fuzz_case(seed)
        plan = Plan(geo, angles)
        if plan.fan_info()["rows_per_tile"] < 1:
            return
        plan.set_projector("fan")
        got = plan.ax(plan.pack(torch.from_numpy(vol).cuda())).cpu().numpy()
        assert plan.fault() == 0
        ref = oracle.ax(vol, geo, angles)
        if np.linalg.norm(ref) == 0:
            return
        err = rel_l2(got, ref)
        seen.append(err)
        assert err <= 1e-4, (seed, err)
        assert err <= 5e-6, (seed, err)

    run()
    assert len(seen) >= 10


@pytest.mark.slow
def test_c1_full_size_pipeline_matches_oracle(oracle):
    """BASELINE.json's configuration C1 at FULL size (128^3 @ 2 mm, 256 angles, 256x256 detector @ 2 mm,
    scatter + noise induction, 5-iteration OS-SART with blocksize 20) -- "the reference's CPU-runnable case" --
    through the same sequence on the float64 oracle: about a minute of CPU."""
    from physics_arx_b200.phantom import artifact_volume, config_c1, lung_phantom
    from physics_arx_b200.pipeline import SCBCTPipeline
    cfg = config_c1()
    geo, angles = cfg["geo"], cfg["angles"]
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    pipe = SCBCTPipeline(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], I0=cfg["I0"], sigma=cfg["sigma"],
                         scatter_scale=0.5, seed=5)
    got = pipe.run_host(ct, art).numpy()
    p = oracle.ax(ct.astype(np.float64) + 0.5 * art.astype(np.float64), geo, angles)
    p = oracle.ctnoise_add(p.astype(np.float32), Poisson=cfg["I0"], Gaussian=np.array([0.0, cfg["sigma"]]), seed=5)
    ref = oracle.ossart(p.astype(np.float32), geo, angles, cfg["niter"], blocksize=cfg["blocksize"])
    d = np.abs(got.astype(np.float64) - ref) * HU
    print("C1 full size: mean |dHU|", d.mean(), "max", d.max())
    assert d.mean() <= 1.0 and d.max() <= 5.0, (d.mean(), d.max())
    assert d.mean() <= 0.01, d.mean()
