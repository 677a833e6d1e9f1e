"""Artifact-Induction tools: the oracle's restatement of the ITK classes the
reference's command-line tools instantiate (known answers, CPU), and the CUDA versions against it (GPU)."""
import numpy as np
import pytest

from conftest import rel_l2


def _imgs(seed=0):
    """A small planning CT (thick slices) and a CBCT on a different, shifted, finer grid."""
    from physics_arx_b200.io_nrrd import NrrdImage
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(14), np.arange(30), np.arange(34), indexing="ij")
    pct = (-800 + 900 * np.exp(-((x - 17) ** 2 + (y - 15) ** 2) / 90.0) + 40 * rng.standard_normal(x.shape))
    pct = NrrdImage(np.round(pct).astype(np.int16), spacing=(1.2, 1.2, 2.5), origin=(-20.0, -18.0, -17.0))
    z, y, x = np.meshgrid(np.arange(30), np.arange(40), np.arange(36), indexing="ij")
    cb = (-750 + 800 * np.exp(-((x - 18) ** 2 + (y - 21) ** 2) / 120.0) + 60 * rng.standard_normal(x.shape)
          + 3.0 * z)
    cbct = NrrdImage(np.round(cb).astype(np.int16), spacing=(1.0, 1.0, 1.0), origin=(-17.0, -21.0, -14.0))
    return pct, cbct


# ------------------------------------------------------------------------- oracle known answers
def test_oracle_resample_identity_and_ramp(oracle):
    rng = np.random.default_rng(1)
    a = rng.random((5, 6, 7))
    m = oracle.index_map((0, 0, 0), (1, 1, 1), np.eye(3), (0, 0, 0), (1, 1, 1), np.eye(3))
    assert np.array_equal(oracle.resample_linear(a, m, a.shape), a)
    # a linear ramp is reproduced exactly by linear interpolation wherever the sample is inside
    z, y, x = np.meshgrid(np.arange(8.0), np.arange(9.0), np.arange(10.0), indexing="ij")
    ramp = 2 * x - 3 * y + 0.5 * z
    m = oracle.index_map((1.3, 0.7, 2.1), (0.5, 0.8, 0.6), np.eye(3), (0, 0, 0), (1, 1, 1), np.eye(3))
    out = oracle.resample_linear(ramp, m, (6, 7, 8))
    k, j, i = np.meshgrid(np.arange(6.0), np.arange(7.0), np.arange(8.0), indexing="ij")
    want = 2 * (1.3 + 0.5 * i) - 3 * (0.7 + 0.8 * j) + 0.5 * (2.1 + 0.6 * k)
    assert np.allclose(out, want, atol=1e-12)


def test_oracle_resample_border_rules(oracle):
    """ITK: inside means continuous index in [-0.5, n-0.5); inside, the outer half voxel repeats the edge."""
    a = np.arange(4.0).reshape(1, 1, 4) + 10.0
    for shift, want in ((-0.5, [10.0]), (-0.51, [0.0]), (3.49, [13.0]), (3.5, [0.0]), (2.75, [12.75])):
        m = oracle.index_map((shift, 0, 0), (1, 1, 1), np.eye(3), (0, 0, 0), (1, 1, 1), np.eye(3))
        assert np.allclose(oracle.resample_linear(a, m, (1, 1, 1)).ravel(), want), shift
    # short pixels: truncation toward zero, like static_cast<short>
    b = np.array([[[-3.0, 4.0]]])
    m = oracle.index_map((0.25, 0, 0), (1, 1, 1), np.eye(3), (0, 0, 0), (1, 1, 1), np.eye(3))
    assert oracle.resample_linear(b, m, (1, 1, 1), cast="short").ravel()[0] == -1.0      # -1.25 -> -1
    m = oracle.index_map((0.9, 0, 0), (1, 1, 1), np.eye(3), (0, 0, 0), (1, 1, 1), np.eye(3))
    assert oracle.resample_linear(b, m, (1, 1, 1), cast="short").ravel()[0] == 3.0       # 3.3 -> 3


def test_oracle_rescale_intensity(oracle):
    rng = np.random.default_rng(2)
    a = rng.normal(size=(4, 5, 6)) * 300 - 500
    r = oracle.rescale_intensity(a, 0, 1)
    assert r.min() == 0.0 and abs(r.max() - 1.0) < 1e-15
    assert np.allclose(r, (a - a.min()) / (a.max() - a.min()), atol=1e-14)
    assert np.all(oracle.rescale_intensity(np.full((2, 2, 2), 7.0), 0, 1) == 0.0)


def test_oracle_plahe_known_answers(oracle):
    rng = np.random.default_rng(3)
    a = rng.random((7, 8, 9)) * 1000 - 300
    # alpha = beta = 1: F(u,v) = u, the filter is the identity
    assert np.allclose(oracle.plahe(a, 2, 1.0, 1.0), a, atol=1e-9)
    # alpha = 0, beta = 0: classical adaptive histogram equalisation = rank within the window
    out = oracle.plahe(a, 1, 0.0, 0.0)
    k, j, i = 3, 4, 5
    win = a[k - 1:k + 2, j - 1:j + 2, i - 1:i + 2]
    rank = (np.sum(win < a[k, j, i]) - np.sum(win > a[k, j, i])) / win.size
    lo, hi = a.min(), a.max()
    assert abs(out[k, j, i] - ((hi - lo) * (0.5 * rank + 0.5) + lo)) < 1e-9
    # alpha = 1, beta = 0: unsharp masking, out = in - windowed mean + mid-grey
    out = oracle.plahe(a, 1, 1.0, 0.0)
    assert abs(out[k, j, i] - (a[k, j, i] - win.mean() + 0.5 * (hi + lo))) < 1e-9
    # at a corner the window is clipped to the image (N = 8, not 27)
    c = a[:2, :2, :2]
    assert abs(out[0, 0, 0] - (a[0, 0, 0] - c.mean() + 0.5 * (hi + lo))) < 1e-9
    # alpha = 0.5: F = 0.5 sgn(u-v) sqrt(2|u-v|) + beta v, checked term by term at one voxel
    out = oracle.plahe(a, 1, 0.5, 0.5)
    u, v = (a[k, j, i] - lo) / (hi - lo) - 0.5, (win - lo) / (hi - lo) - 0.5
    f = 0.5 * np.sign(u - v) * np.sqrt(np.abs(2 * (u - v))) + 0.5 * v
    assert abs(out[k, j, i] - ((hi - lo) * (f.mean() + 0.5) + lo)) < 1e-9


def test_index_map_matches_itk_physical_space():
    from physics_arx_b200.artifact_induction import index_map
    d = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])        # 90 degree in-plane rotation
    m = index_map((5.0, 6.0, 7.0), (2.0, 2.0, 3.0), d, (1.0, 2.0, 3.0), (0.5, 0.5, 1.5), np.eye(3)).reshape(3, 4)
    idx = np.array([3.0, 4.0, 2.0])
    phys = np.array([5.0, 6.0, 7.0]) + d @ (np.array([2.0, 2.0, 3.0]) * idx)
    want = (phys - np.array([1.0, 2.0, 3.0])) / np.array([0.5, 0.5, 1.5])
    assert np.allclose(m[:, :3] @ idx + m[:, 3], want)


# ---------------------------------------------------------------------------- CUDA vs oracle
@pytest.mark.gpu
def test_gpu_resample_matches_oracle(oracle):
    from physics_arx_b200 import artifact_induction as ai
    pct, cbct = _imgs()
    for fn, cast in ((ai.ResampleImages, "short"), (ai.ResampleImagesFloat, "float")):
        got = fn(pct, cbct)
        m = oracle.index_map(pct.origin, pct.spacing, pct.direction, cbct.origin, cbct.spacing, cbct.direction)
        ref = oracle.resample_linear(np.asarray(cbct.array, dtype=np.float64), m, pct.array.shape, cast=cast)
        g = got.data.cpu().numpy().astype(np.float64)
        assert got.size == pct.GetSize() and got.spacing == pct.spacing
        if cast == "short":
            assert np.mean(g != ref) < 1e-4 and np.abs(g - ref).max() <= 1.0      # a tie at an integer may flip
        else:
            assert np.allclose(g, ref, rtol=1e-6, atol=1e-3)
        assert (ref == 0).any() and (ref != 0).any()          # the CBCT does not cover the whole planning CT


@pytest.mark.gpu
@pytest.mark.parametrize("alpha,beta", [(0.0, 1.0), (1.0, 0.0), (0.5, 0.5), (0.5, 0.0), (0.0, 0.5), (0.5, 1.0),
                                        (1.0, 0.5), (1.0, 1.0), (0.3, 0.8)])
def test_gpu_plahe_matches_oracle(oracle, alpha, beta):
    from physics_arx_b200 import artifact_induction as ai
    pct, cbct = _imgs()
    img = ai.ResampleImages(pct, cbct)
    got = ai.AdaptiveHistogramMatch(img, 5, alpha, beta).data.cpu().numpy()
    ref = oracle.plahe(img.data.cpu().numpy(), 5, alpha, beta)
    span = ref.max() - ref.min()
    assert np.abs(got - ref).max() <= 2e-5 * span, (np.abs(got - ref).max(), span)
    if alpha == 1.0 and beta == 1.0:
        assert np.allclose(got, img.data.cpu().numpy(), atol=2e-5 * span)


@pytest.mark.gpu
def test_gpu_add_and_rescale_match_oracle(oracle):
    from physics_arx_b200 import artifact_induction as ai
    pct, cbct = _imgs()
    s = ai.AddImages(pct, cbct)
    m = oracle.index_map(pct.origin, pct.spacing, pct.direction, cbct.origin, cbct.spacing, cbct.direction)
    ref = pct.array.astype(np.float64) + oracle.resample_linear(cbct.array.astype(np.float64), m, pct.array.shape)
    assert np.allclose(s.data.cpu().numpy(), ref, rtol=1e-6, atol=1e-3)
    r = ai.RescaleImage(s, 0, 1, 1)
    nz = int(pct.array.shape[0] * 2.5 / 1.2)
    assert r.size == (34, 30, nz) and r.spacing == (1.2, 1.2, 1.2)
    m2 = oracle.index_map(pct.origin, (1.2, 1.2, 1.2), pct.direction, pct.origin, pct.spacing, pct.direction)
    ref2 = oracle.rescale_intensity(oracle.resample_linear(ref.astype(np.float32).astype(np.float64), m2,
                                                           (nz, 30, 34)), 0, 1)
    g = r.data.cpu().numpy()
    assert g.min() == 0.0 and g.max() == 1.0
    assert np.abs(g - ref2).max() <= 1e-6


@pytest.mark.gpu
def test_gpu_artifact_adding_chain(oracle):
    """CBCT_pCT_Artifact_adding.sh steps 1-4 for one patient and two variations, against the same chain
    of oracle calls."""
    from physics_arx_b200 import artifact_induction as ai
    pct, cbct = _imgs()
    out = ai.cbct_pct_artifact_adding(pct, cbct, variations=((0.5, 0.5), (0.0, 1.0)))
    assert {"CBCT_w1", "CBCT_Histogram_0.5_0.5", "pCT_Artifact_0_1", "rsc_pCT", "CBCT_rsc_w1",
            "pCT_Artifact_0.5_0.5_rsc"} <= set(out)
    m = oracle.index_map(pct.origin, pct.spacing, pct.direction, cbct.origin, cbct.spacing, cbct.direction)
    w1 = oracle.resample_linear(cbct.array.astype(np.float64), m, pct.array.shape, cast="short")
    hist = oracle.plahe(w1, 5, 0.5, 0.5)
    added = pct.array.astype(np.float64) + hist
    nz = int(pct.array.shape[0] * 2.5 / 1.2)
    m2 = oracle.index_map(pct.origin, (1.2, 1.2, 1.2), pct.direction, pct.origin, pct.spacing, pct.direction)
    ref = oracle.rescale_intensity(oracle.resample_linear(added, m2, (nz, 30, 34)), 0, 1)
    got = out["pCT_Artifact_0.5_0.5_rsc"].data.cpu().numpy()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= 1e-5 and np.abs(got - ref).max() <= 1e-4
    # the rescaled artifact volume is what the reconstruction script reads: in [0,1], float32
    assert got.dtype == np.float32 and got.min() == 0.0 and got.max() == 1.0
