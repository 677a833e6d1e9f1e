"""NRRD I/O (the file format on either side of the path) and the reference-shaped case driver."""
import os

import numpy as np
import pytest

REF_NRRD = "/root/reference/results/msk_aapm_stabilized_eso4/test_latest/npz_images/LCTSC-Test-S1-101_1_RTSTRUCTS.nrrd"


@pytest.mark.parametrize("enc", ["raw", "gzip"])
@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8])
def test_nrrd_roundtrip(tmp_path, enc, dtype):
    from physics_arx_b200.io_nrrd import read_nrrd, write_nrrd
    rng = np.random.default_rng(0)
    arr = (rng.random((5, 7, 9)) * 100).astype(dtype)
    d = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    p = tmp_path / "v.nrrd"
    write_nrrd(str(p), arr, spacing=(0.9766, 0.9766, 2.5), origin=(-250.0, -250.0, 10.5), direction=d, encoding=enc)
    img = read_nrrd(str(p))
    assert np.array_equal(img.array, arr) and img.array.dtype == np.dtype(dtype)
    assert img.GetSize() == (9, 7, 5)
    assert np.allclose(img.GetSpacing(), (0.9766, 0.9766, 2.5))
    assert np.allclose(img.GetOrigin(), (-250.0, -250.0, 10.5))
    assert np.allclose(img.direction, d)


@pytest.mark.skipif(not os.path.exists(REF_NRRD), reason="reference tree not present on this box")
def test_reads_the_reference_trees_own_nrrd():
    from physics_arx_b200.io_nrrd import read_nrrd
    img = read_nrrd(REF_NRRD)
    assert img.GetSize() == (128, 128, 128) and img.array.dtype == np.uint8
    assert img.GetSpacing() == (1.0, 1.0, 1.0) and img.GetOrigin() == (0.0, 0.0, 0.0)
    assert set(np.unique(img.array)).issubset({0, 1, 2, 3, 4})          # merged RTSTRUCT labels


def test_rejects_what_it_does_not_support(tmp_path):
    from physics_arx_b200.io_nrrd import read_nrrd, write_nrrd
    p = tmp_path / "bad.nrrd"
    p.write_bytes(b"NRRD0004\ntype: float\ndimension: 2\nsizes: 2 2\nencoding: raw\n\n" + b"\0" * 16)
    with pytest.raises(ValueError):
        read_nrrd(str(p))
    with pytest.raises(ValueError):
        write_nrrd(str(tmp_path / "x.nrrd"), np.zeros((2, 2)))


def test_case_geometry_matches_reference_script():
    from physics_arx_b200.reconstruct_cases import FNAMES, case_geometry
    geo = case_geometry((200, 180, 90), (1.1, 1.2, 2.5))
    assert geo.DSD == 1500 and geo.DSO == 1000.0
    assert tuple(geo.nDetector) == (400, 400) and np.allclose(geo.dDetector, 1.0)
    assert tuple(geo.nVoxel) == (90, 180, 200) and np.allclose(geo.dVoxel, (2.5, 1.2, 1.1))
    assert np.allclose(case_geometry((8, 8, 8), (1, 1, 1), half_fan=True).offDetector, (0, -160))
    assert len(FNAMES) == 7 and FNAMES[0] == '0.5_0.5_rsc.nrrd'


def _make_patient(root, n=20):
    from physics_arx_b200.io_nrrd import write_nrrd
    from physics_arx_b200.phantom import lung_phantom
    art = os.path.join(root, "P1", "Artifacts")
    os.makedirs(art)
    vol = lung_phantom((n, n, n), (8.0, 8.0, 8.0), seed=0)
    for a, b in (("0.5", "0.5"), ("1", "0")):
        write_nrrd(os.path.join(art, f"P1_pCT_Artifact_{a}_{b}_rsc.nrrd"), vol, spacing=(8.0, 8.0, 8.0),
                   origin=(1.0, 2.0, 3.0))
    return vol


def test_list_cases(tmp_path):
    from physics_arx_b200.reconstruct_cases import list_cases
    _make_patient(str(tmp_path))
    cases = list_cases(str(tmp_path))
    assert [c[0] for c in cases] == ["P1", "P1"] and len(cases) == 2


@pytest.mark.gpu
def test_case_driver_end_to_end(tmp_path, oracle):
    """Reference-shaped run on a tiny synthetic patient: NRRD in -> NPZ + NRRD out, vs the oracle."""
    from physics_arx_b200.io_nrrd import read_nrrd
    from physics_arx_b200.reconstruct_cases import case_geometry, pCT_CBCT_reconstruction
    vol = _make_patient(str(tmp_path))
    done = pCT_CBCT_reconstruction(str(tmp_path), niter=2, blocksize=10, n_angles=30, seed=4, verbose=False)
    assert len(done) == 2 and all(os.path.exists(f) for f in done)
    out = read_nrrd(done[0])
    assert out.GetSpacing() == (8.0, 8.0, 8.0) and out.GetOrigin() == (1.0, 2.0, 3.0)
    npz = np.load(os.path.join(str(tmp_path), "P1", "Artifacts", "P1_projections_0.5_0.5.npz"))
    assert npz["projection"].shape == (30, 400, 400)
    geo, angles = case_geometry((20, 20, 20), (8.0, 8.0, 8.0)), np.linspace(0, 2 * np.pi, 30)
    ref = oracle.ossart(npz["projection"], geo, angles, 2, blocksize=10)
    hu = np.abs(out.array - ref) * 4095.0
    assert hu.mean() <= 1.0 and hu.max() <= 5.0
