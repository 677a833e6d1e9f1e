"""Size-independent properties of the CUDA path at BASELINE.json's full C2 size (512x512x128 volume,
1024x768 half-fan detector), where the float64 oracle is too slow to run more than a few angles:
linearity of Ax, agreement of the two projector kernels, W.Ax(ones) ~ 1, determinism and descent of
OS-SART.  All through the C ABI (libpax.so)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import torch
    from physics_arx_b200.phantom import artifact_volume, config_c2, lung_phantom
    from physics_arx_b200.projector import Plan
    assert torch.cuda.is_available()
    cfg = config_c2()
    geo = cfg["geo"]
    angles = cfg["angles"][[0, 97, 164, 333, 480, 655]]      # incl. the axis-aligned first and last views
    plan = Plan(geo, angles)
    ct = torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda()
    art = torch.from_numpy(artifact_volume(geo.nVoxel, seed=1)).cuda()
    return dict(geo=geo, angles=angles, plan=plan, ct=ct, art=art, torch=torch)


def test_c2_resolves_to_the_fan_projector_and_fits(c2):
    plan = c2["plan"]
    info = plan.fan_info()
    assert plan.projector == "fan" and info["rows_per_tile"] >= 32 and info["mean_run"] >= 48


def test_ax_is_linear_at_full_size(c2):
    torch, plan = c2["torch"], c2["plan"]
    x, y = plan.pack(c2["ct"]), plan.pack(c2["art"])
    px, py = plan.ax(x), plan.ax(y)
    mix = plan.pack(0.3 * c2["ct"] + 0.7 * c2["art"])
    pm = plan.ax(mix)
    ref = 0.3 * px + 0.7 * py
    err = float((pm - ref).double().norm() / ref.double().norm())
    assert err <= 1e-5, err
    # and through the two-channel epilogue in one call
    p2 = plan.ax(x, res1=y, a0=0.3, a1=0.7)
    assert float((p2 - ref).double().norm() / ref.double().norm()) <= 1e-5
    assert plan.fault() == 0


def test_fan_and_ray_kernels_agree_at_full_size(c2):
    plan = c2["plan"]
    x = plan.pack(c2["ct"])
    plan.set_projector("fan")
    pf = plan.ax(x)
    plan.set_projector("ray")
    pr = plan.ax(x)
    plan.set_projector("auto")
    err = float((pf - pr).double().norm() / pr.double().norm())
    assert err <= 2e-5, err                     # the ray kernel's own error vs the oracle is 4e-6
    assert plan.fault() == 0


def test_w_times_ax_of_ones_is_one_inside_the_field_of_view(c2):
    """Known answer (ii) of SURVEY.md 8(c): W = 1/chord of the box enlarged 1.1x in plane (MATLAB computeW),
    so a ray that crosses the volume through two opposite side faces gives W * Ax(ones) = 1/1.1; rays that
    clip a corner of the enlarged box give less, none gives more than 1."""
    torch, plan, geo = c2["torch"], c2["plan"], c2["geo"]
    ones = plan.pack(torch.ones(tuple(int(v) for v in geo.nVoxel), device="cuda"))
    p = plan.ax(ones)
    w = plan.compute_w()
    r = (w * p)[:, 300:468, :]                  # central rows: rays enter and leave through the sides
    hit = p[:, 300:468, :] > 100.0              # chords longer than 100 mm
    vals = r[hit]
    assert float(vals.min()) > 0.5 and float(vals.max()) <= 1.0 + 1e-4
    assert abs(float(vals.median()) - 1.0 / 1.1) < 0.01


def test_ossart_is_deterministic_and_descends_at_full_size(c2):
    torch, geo = c2["torch"], c2["geo"]
    from physics_arx_b200.algorithms.ossart import OSSARTState
    from physics_arx_b200.phantom import config_c2
    angles = config_c2()["angles"][::16]        # 41 views over the full rotation
    st = OSSARTState(None, geo, angles, blocksize=20)
    p = st.plan
    x_true = p.pack(c2["ct"])
    p.ax(x_true, 0, len(angles), out=st.b)

    def run():
        x = p.new_volume()
        l2 = []
        lam = 1.0
        for _ in range(2):
            st.sumsq.zero_()
            st.iteration(x, lam, True, True)
            l2.append(float(st.sumsq.sqrt().item()))
            lam *= 0.99
        return x, l2

    x1, l2a = run()
    x2, l2b = run()
    assert torch.equal(x1, x2)                  # no atomics on the volume path
    assert l2a[1] < 0.6 * l2a[0]                # residual norm drops (known answer vii)
    assert np.allclose(l2a, l2b, rtol=1e-6)
    assert float(x1.min()) >= 0.0               # noneg clip
    assert p.fault() == 0
