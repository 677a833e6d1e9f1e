"""Output side of the path: the 128^3 resize + NPZ {CT, CBCT, RTSTRUCTS} packing of
preprocess/prep_test_data_pseudo_cbct_3D.py:23-64 and the Quameasopts quality measures that
pCT_CBCT_Reconstruction.py:99-103 prepares -- oracle restatements against known answers (CPU), device kernels
against the oracle (GPU)."""
import os

import numpy as np
import pytest


# ------------------------------------------------------------------------------ oracle, CPU
def test_resize_oracle_known_answers():
    from oracle.quality_oracle import resize_nd
    rng = np.random.default_rng(0)
    a = rng.random((8, 6, 4))
    assert np.array_equal(resize_nd(a, a.shape, 1), a)                       # same shape: identity
    # halving with order 1 samples at 2o + 1/2: the mean of the two neighbours
    h = resize_nd(a, (4, 6, 4), 1)
    assert np.allclose(h, 0.5 * (a[0::2] + a[1::2]))
    # order 0 on a mask keeps the label set, and at 2o + 1/2 takes floor(c + 1/2) = 2o + 1
    m = (rng.random((8, 8, 8)) > 0.5).astype(np.uint8)
    n0 = resize_nd(m, (4, 4, 4), 0)
    assert set(np.unique(n0)).issubset({0.0, 1.0}) and np.array_equal(n0, m[1::2, 1::2, 1::2])
    # a linear ramp survives order-1 resampling in the interior
    r = np.arange(10, dtype=np.float64)[:, None, None] * np.ones((10, 3, 3))
    up = resize_nd(r, (25, 3, 3), 1)
    c = 0.4 * (np.arange(25) + 0.5) - 0.5
    inside = (c >= 0) & (c <= 9)
    assert np.allclose(up[inside, 0, 0], c[inside])


def test_quality_oracle_known_answers():
    from oracle.quality_oracle import measure_quality
    rng = np.random.default_rng(1)
    a = rng.random((5, 6, 7))
    q = measure_quality(a, a, ["RMSE", "CC", "MSSIM", "UQI"])
    assert np.allclose(q, [0.0, 1.0, 1.0, 1.0])
    b = a + 0.25
    assert np.isclose(measure_quality(a, b, "RMSE")[0], 0.25)
    assert np.isclose(measure_quality(a, b, "CC")[0], 1.0)
    assert np.isclose(measure_quality(a, 1.0 - a, "CC")[0], -1.0)
    assert measure_quality(a, rng.random(a.shape), "UQI")[0] < 0.5


def test_quality_from_sums_matches_oracle():
    """host half of the device path (six sums -> measures) against the direct two-pass oracle"""
    from oracle.quality_oracle import measure_quality
    from physics_arx_b200.utilities.quality import check_opts, from_sums
    rng = np.random.default_rng(2)
    a, b = rng.random(4000) * 3, rng.random(4000) * 2 + 0.3
    sums = [a.size, a.sum(), b.sum(), (a * a).sum(), (b * b).sum(), (a * b).sum()]
    opts = ["RMSE", "CC", "MSSIM", "UQI"]
    got = from_sums(sums, min(a.min(), b.min()), max(a.max(), b.max()), opts)
    assert np.allclose(got, measure_quality(a, b, opts), rtol=1e-10)
    with pytest.raises(ValueError):
        check_opts(["PSNR"])


def test_combine_rtstructs_priority():
    from physics_arx_b200.dataprep import combine_rtstructs
    z = np.zeros((2, 2, 2))
    heart, lungs, cord, eso = z.copy(), z.copy(), z.copy(), z.copy()
    eso[0, 0, 0] = cord[0, 0, 0] = 1          # cord over eso
    heart[0, 0, 1] = eso[0, 0, 1] = 1         # heart over eso
    lungs[0, 1, 0] = heart[0, 1, 0] = 1       # lungs over heart
    eso[1, 1, 1] = 1
    rt = combine_rtstructs(heart, lungs, cord, eso)
    assert rt.dtype == np.uint8
    assert rt[0, 0, 0] == 3 and rt[0, 0, 1] == 2 and rt[0, 1, 0] == 1 and rt[1, 1, 1] == 4 and rt[1, 0, 0] == 0


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape,out,order", [((20, 24, 16), (8, 8, 8), 1), ((20, 24, 16), (8, 8, 8), 0),
                                             ((9, 7, 11), (16, 16, 16), 1), ((9, 7, 11), (16, 16, 16), 0),
                                             ((64, 64, 40), (32, 32, 32), 1), ((12, 12, 12), (12, 12, 12), 1)])
def test_resize_matches_oracle(shape, out, order):
    from oracle.quality_oracle import resize_nd
    from physics_arx_b200.dataprep import resize
    rng = np.random.default_rng(3)
    a = rng.random(shape).astype(np.float32) if order == 1 else (rng.random(shape) > 0.6).astype(np.float32)
    got = resize(a, out, order)
    ref = resize_nd(a.astype(np.float64), out, order)
    assert got.shape == tuple(out) and got.dtype == np.float32
    if order == 0:
        assert np.array_equal(got, ref.astype(np.float32))
    else:
        assert np.abs(got - ref).max() <= 1e-6


@pytest.mark.gpu
def test_measure_quality_matches_oracle():
    import torch
    from oracle.quality_oracle import measure_quality
    from physics_arx_b200.utilities.quality import Measure_Quality
    rng = np.random.default_rng(4)
    a = rng.random((40, 50, 60)).astype(np.float32)
    b = (a + 0.05 * rng.standard_normal(a.shape)).astype(np.float32)
    opts = ["RMSE", "CC", "MSSIM", "UQI"]
    got = Measure_Quality(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), opts)
    assert np.allclose(got, measure_quality(a, b, opts), rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
def test_ossart_quameasopts_and_computel2(oracle):
    """ossart(..., Quameasopts=[...], computel2=True): (res, l2, quality[measure, iteration]); the measures are
    those of the oracle's iterates."""
    import physics_arx_b200 as pax
    from oracle.quality_oracle import measure_quality
    from physics_arx_b200.phantom import lung_phantom, make_geometry
    geo = make_geometry((12, 24, 24), (8.0, 6.0, 6.0), (24, 40), (6.0, 5.0))
    angles = np.linspace(0, 2 * np.pi, 20)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    b = oracle.ax(vol, geo, angles).astype(np.float32)
    opts = ["RMSE", "CC", "MSSIM", "UQI"]
    res, l2, q = pax.algorithms.ossart(b, geo, angles, 3, blocksize=5, Quameasopts=opts, computel2=True)
    assert q.shape == (4, 3) and l2.shape == (3,)
    prev = np.zeros_like(vol, dtype=np.float64)
    for it in range(3):
        cur = oracle.ossart(b, geo, angles, it + 1, blocksize=5)
        want = measure_quality(prev, cur, opts)
        ok = np.isfinite(want)
        assert np.allclose(q[ok, it], want[ok], rtol=2e-3, atol=1e-5), (it, q[:, it], want)
        prev = cur
    assert q[0, 2] < q[0, 0]                     # the iterate moves less and less
    with pytest.raises(ValueError):
        pax.algorithms.ossart(b, geo, angles, 1, Quameasopts=["PSNR"])


@pytest.mark.gpu
def test_process_case_packs_what_the_dataset_reads(tmp_path):
    """A synthetic case folder with the reference's file names -> <case>.npz with CT / CBCT (float, 128^3,
    (Z,Y,X)) and RTSTRUCTS (uint8 labels), the keys data/cbct2ct3d_dataset.py:53-58 loads."""
    from oracle.quality_oracle import resize_nd
    from physics_arx_b200.dataprep import process_case
    from physics_arx_b200.io_nrrd import write_nrrd
    rng = np.random.default_rng(5)
    case = tmp_path / "in" / "P7"
    case.mkdir(parents=True)
    shape_zyx = (40, 64, 64)
    ct = (rng.random(shape_zyx) * 1.5).astype(np.float32)            # max > 1: gets normalised
    cbct = rng.random(shape_zyx).astype(np.float32) * 0.9
    masks = {k: (rng.random(shape_zyx) > 0.8).astype(np.uint8) for k in ("Heart", "Lungs", "Cord", "Eso")}
    write_nrrd(str(case / "P7_CT_plan_50.nrrd"), ct)
    write_nrrd(str(case / "P7_pCT_OSSART_0.5_0.5.nrrd"), cbct)
    for k, m in masks.items():
        write_nrrd(str(case / f"P7_{k}.nrrd"), m)
    out = process_case(str(tmp_path / "in"), "P7", str(tmp_path / "out"))
    assert os.path.basename(out) == "P7.npz"
    z = np.load(out)
    assert sorted(z.files) == ["CBCT", "CT", "RTSTRUCTS"]
    assert z["CT"].shape == z["CBCT"].shape == z["RTSTRUCTS"].shape == (128, 128, 128)
    assert z["RTSTRUCTS"].dtype == np.uint8 and set(np.unique(z["RTSTRUCTS"])).issubset({0, 1, 2, 3, 4})
    # CT: resized in (X,Y,Z) index order, swapped back to (Z,Y,X), divided by its max
    ref = np.swapaxes(resize_nd(np.transpose(ct, (2, 1, 0)).astype(np.float64), (128, 128, 128), 1), 0, -1)
    ref = ref / ref.max()
    assert np.abs(z["CT"] - ref).max() <= 2e-6 and abs(float(z["CT"].max()) - 1.0) <= 1e-6
    ref_b = np.swapaxes(resize_nd(np.transpose(cbct, (2, 1, 0)).astype(np.float64), (128, 128, 128), 1), 0, -1)
    assert np.abs(z["CBCT"] - ref_b).max() <= 2e-6                    # max <= 1: left alone
