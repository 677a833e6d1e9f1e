"""CPU-side checks: the C-ABI library loads and exports every symbol include/pax.h declares
(no compute calls -- there is no GPU here), and the host logic (geometry bag, subset ordering,
angle sharding, error behaviour) does what the reference's call surface expects."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pax.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pax_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from physics_arx_b200.csrc import build as pax_build
    so = pax_build.build()
    lib = ctypes.CDLL(so)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libpax.so does not export {name}"
    from physics_arx_b200 import _lib
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    assert _lib.load().pax_version() == 100


def test_abi_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is checkable on CPU."""
    from physics_arx_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    geo = _lib.PaxGeo(4, 4, 4, 4, 4, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5)
    rows = np.array([[0.0, 1500.0, 1000.0, 0, 0, 0, 0, 0]], dtype=np.float64)
    assert lib.pax_plan_create(None, rows.ctypes.data_as(_lib.c_double_p), 1, ctypes.byref(h)) == -1
    assert b"NULL" in lib.pax_last_error()
    assert lib.pax_plan_create(ctypes.byref(geo), rows.ctypes.data_as(_lib.c_double_p), 0, ctypes.byref(h)) == -1
    bad = _lib.PaxGeo(4, 4, 4, 4, 4, 1.0, 1.0, -1.0, 1.0, 1.0, 0.5)
    assert lib.pax_plan_create(ctypes.byref(bad), rows.ctypes.data_as(_lib.c_double_p), 1, ctypes.byref(h)) == -1
    rows[0, 1] = 900.0   # DSD < DSO
    assert lib.pax_plan_create(ctypes.byref(geo), rows.ctypes.data_as(_lib.c_double_p), 1, ctypes.byref(h)) == -1
    assert b"DSD" in lib.pax_last_error()
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    assert lib.pax_plan_destroy(None) == 0


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import physics_arx_b200 as pax
    from physics_arx_b200 import _lib
    geo = pax.geometry_default()
    with pytest.raises(_lib.PaxError):
        pax.Ax(np.zeros((64, 64, 64), np.float32), geo, np.linspace(0, 1, 3), "interpolated")
    with pytest.raises(_lib.PaxError):
        pax.algorithms.ossart(np.zeros((3, 128, 128), np.float32), geo, np.linspace(0, 1, 3), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "physics_arx_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "cpu_oracle" not in src and "numpy_ref" not in src and "liboracle" not in src, f


def test_geometry_default_and_reference_overrides():
    """The attribute edits of pCT_CBCT_Reconstruction.py:67-78 work on the default geometry."""
    import physics_arx_b200.compat as tigre
    geo = tigre.geometry_default(high_resolution=False)
    assert geo.DSD == 1536.0 and geo.DSO == 1000.0 and geo.accuracy == 0.5 and geo.mode == "cone"
    assert tuple(geo.nVoxel) == (64, 64, 64) and tuple(geo.nDetector) == (128, 128)
    angles = np.linspace(0, 2 * np.pi, 500)
    geo.DSD = 1500
    geo.nDetector = np.array((400, 400))
    geo.sDetector = np.array((400, 400))
    geo.dDetector = geo.sDetector / geo.nDetector
    sz, spc = (200, 180, 90), (1.1, 1.1, 2.5)
    geo.nVoxel = np.array((sz[2], sz[1], sz[0]))
    geo.sVoxel = np.array((sz[2] * spc[2], sz[1] * spc[1], sz[0] * spc[0]))
    geo.dVoxel = geo.sVoxel / geo.nVoxel
    g, rows = geo.pack(angles)
    assert rows.shape == (500, 8) and np.all(rows[:, 1] == 1500) and np.all(rows[:, 2] == 1000)
    assert list(g[:5]) == [90, 180, 200, 400, 400] and g[10] == 0.5
    assert rows[0, 0] == 0 and abs(rows[-1, 0] - 2 * np.pi) < 1e-12     # endpoint included (:68)
    geo.offDetector = np.array((0, -160))                                 # the half-fan hook (:83)
    _, rows = geo.pack(angles)
    assert np.all(rows[:, 7] == -160) and np.all(rows[:, 6] == 0)


def test_geometry_validation():
    from physics_arx_b200 import geometry_default
    geo = geometry_default()
    with pytest.raises(ValueError):
        geo.check_geo(np.zeros((3, 2)))
    g2 = geo.copy(); g2.sVoxel = np.array((1.0, 2.0, 3.0))
    with pytest.raises(ValueError):
        g2.check_geo([0.0])
    g3 = geo.copy(); g3.rotDetector = np.array((0.1, 0, 0))
    with pytest.raises(NotImplementedError):
        g3.check_geo([0.0])
    g4 = geo.copy(); g4.mode = "parallel"
    with pytest.raises(NotImplementedError):
        g4.check_geo([0.0])
    with pytest.raises(NotImplementedError):
        geo.check_geo(np.array([[0.0, 0.1, 0.0]]))
    g5 = geo.copy(); g5.DSD = np.linspace(1500, 1600, 4)
    _, rows = g5.pack(np.zeros(4))
    assert np.allclose(rows[:, 1], np.linspace(1500, 1600, 4))
    with pytest.raises(ValueError):
        g5.pack(np.zeros(5))


def test_w_box_variants():
    from physics_arx_b200.phantom import config_c2
    geo = config_c2()["geo"]
    (hz, hy, hx), thr = geo.w_box("matlab")
    assert abs(hy - 275.0) < 1e-9 and abs(hx - 275.0) < 1e-9 and abs(hz - 160.0) < 1e-9
    assert abs(thr - 0.5 * 500 / 512) < 1e-12
    (hz, hy, hx), thr = geo.w_box("python2021")
    assert hz == 250.0 and hy == 250.0 and hx == 250.0 and abs(thr - 0.25 * 500 / 512) < 1e-12


def test_order_subsets_matches_reference_blocking():
    from physics_arx_b200 import order_subsets
    blocks = order_subsets(500, 20)          # reference: 500 angles, blocksize 20 -> 25 subsets
    assert len(blocks) == 25 and all(len(b) == 20 for b in blocks)
    blocks = order_subsets(656, 20)          # C2: 32 x 20 + 16
    assert len(blocks) == 33 and len(blocks[-1]) == 16
    assert np.array_equal(np.concatenate(blocks), np.arange(656))
    r = order_subsets(100, 7, "random", seed=1)
    assert sorted(np.concatenate(r).tolist()) == list(range(100))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_blocks_partitions_every_subset(world):
    from physics_arx_b200 import order_subsets
    from physics_arx_b200.dist import shard_blocks
    blocks = order_subsets(656, 20)
    seen = []
    for r in range(world):
        local, ranges = shard_blocks(blocks, r, world)
        assert len(ranges) == len(blocks)
        for (first, count), blk in zip(ranges, blocks):
            assert np.array_equal(local[first:first + count], blk[r::world])
            assert abs(count - len(blk) / world) < 1
        seen.append(local)
    assert sorted(np.concatenate(seen).tolist()) == list(range(656))


def test_contiguous_runs():
    from physics_arx_b200.pipeline import _contiguous_runs
    assert _contiguous_runs([0, 1, 2, 5, 6, 9]) == [(0, 3), (5, 2), (9, 1)]
    assert _contiguous_runs([]) == []


def test_phantoms_are_seeded_and_in_range():
    from physics_arx_b200.phantom import artifact_volume, config_c1, lung_phantom
    c = config_c1()
    a = lung_phantom((16, 32, 32), (8.0, 8.0, 8.0), seed=0)
    b = lung_phantom((16, 32, 32), (8.0, 8.0, 8.0), seed=0)
    assert np.array_equal(a, b) and a.dtype == np.float32 and a.min() >= 0 and a.max() <= 1
    assert len(np.unique(a)) > 5
    art = artifact_volume((16, 32, 32), seed=1)
    assert abs(art.mean()) < 1e-3 and np.abs(art).max() <= 0.0501
    assert len(c["angles"]) == 256 and c["niter"] == 5


@pytest.mark.parametrize("world", [1, 2, 3, 8, 40])
def test_shard_rows_partitions_the_detector(world):
    from physics_arx_b200.dist import shard_rows
    for nv in (768, 256, 100, 33, 7):
        bands = [shard_rows(nv, r, world) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == nv
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        assert all(v1 >= v0 for v0, v1 in bands)
    assert shard_rows(768, 3, 8) == (288, 384)


def test_choose_shard_auto():
    from physics_arx_b200 import order_subsets
    from physics_arx_b200.dist import choose_shard
    blocks = order_subsets(656, 20)                      # 32 x 20 + 16
    assert choose_shard(blocks, 2) == (2, 1) and choose_shard(blocks, 4) == (4, 1)      # plain angle sharding
    assert choose_shard(blocks, 8) == (4, 2)             # 5 (4) angles x half the detector columns on every rank
    assert choose_shard(order_subsets(40, 7), 4) == (1, 4)   # 7 and 5 angles share no factor with 4: bands only
    assert choose_shard(blocks, 8, "angles") == (8, 1) and choose_shard(blocks, 8, "rows") == (1, 8)
    assert choose_shard(blocks, 8, "cols") == (1, 8)
    from physics_arx_b200.dist import band_axis
    assert band_axis("auto") == band_axis("grid") == band_axis((4, 2)) == "cols" and band_axis("rows") == "rows"
    assert choose_shard(blocks, 8, (2, 4)) == (2, 4)
    with pytest.raises(ValueError):
        choose_shard(blocks, 8, (3, 2))
    with pytest.raises(ValueError):
        choose_shard(blocks, 2, "columns")


def test_row_band_is_a_detector_of_its_own(oracle):
    """Ax of a band == rows of the full Ax; Atb / V over the bands sum to the full Atb / V."""
    from conftest import rel_l2
    from physics_arx_b200.dist import row_window_geometry, shard_rows
    from physics_arx_b200.phantom import make_geometry
    geo = make_geometry((8, 16, 16), (14.0, 9.0, 9.0), (22, 30), (8.0, 7.0), off_detector=(3.0, -25.0))
    ang = np.array([0.2, 1.9, 3.3, 5.0])
    rng = np.random.default_rng(3)
    vol = rng.random((8, 16, 16))
    proj = rng.standard_normal((4, 22, 30))
    full_ax, full_atb = oracle.ax(vol, geo, ang), oracle.atb(proj, geo, ang)
    full_v = oracle.atb(None, geo, ang, vscale=geo.v_scale())
    full_w = oracle.compute_w(geo, ang)
    acc, vacc = 0.0, 0.0
    for r in range(3):
        v0, v1 = shard_rows(22, r, 3, align=4)
        gb = row_window_geometry(geo, v0, v1)
        assert rel_l2(oracle.ax(vol, gb, ang), full_ax[:, v0:v1]) < 1e-12
        acc = acc + oracle.atb(proj[:, v0:v1], gb, ang)
        vacc = vacc + oracle.atb(None, gb, ang, vscale=geo.v_scale())
        # W must come from the FULL geometry's box (its z size depends on sDetector)
        (hz, hy, hx), thr = geo.w_box("matlab")
        gb_w = gb.copy()
        assert gb.w_box("matlab")[0][0] != hz or v1 - v0 == 22
    assert rel_l2(acc, full_atb) < 1e-12
    assert rel_l2(vacc, full_v) < 1e-12
    assert full_w.shape == (4, 22, 30)


def test_column_band_is_a_detector_of_its_own(oracle):
    """Ax of a band of detector columns == those columns of the full Ax; Atb / V over the bands sum to the
    full Atb / V (what the grid sharding of the GPUs relies on)."""
    from conftest import rel_l2
    from physics_arx_b200.dist import col_window_geometry, shard_rows
    from physics_arx_b200.phantom import make_geometry
    geo = make_geometry((8, 16, 16), (14.0, 9.0, 9.0), (22, 30), (8.0, 7.0), off_detector=(3.0, -25.0))
    ang = np.array([0.2, 1.9, 3.3, 5.0])
    rng = np.random.default_rng(4)
    vol = rng.random((8, 16, 16))
    proj = rng.standard_normal((4, 22, 30))
    full_ax, full_atb = oracle.ax(vol, geo, ang), oracle.atb(proj, geo, ang)
    full_v = oracle.atb(None, geo, ang, vscale=geo.v_scale())
    acc, vacc = 0.0, 0.0
    for r in range(3):
        u0, u1 = shard_rows(30, r, 3, align=8)
        gb = col_window_geometry(geo, u0, u1)
        assert tuple(gb.nDetector) == (22, u1 - u0)
        assert rel_l2(oracle.ax(vol, gb, ang), full_ax[:, :, u0:u1]) < 1e-12
        acc = acc + oracle.atb(proj[:, :, u0:u1], gb, ang)
        vacc = vacc + oracle.atb(None, gb, ang, vscale=geo.v_scale())
        # the 'matlab' W box depends on sDetector's ROW extent only, so a band of columns keeps it (the plans of
        # the ranks take the box from the full geometry anyway: Plan.w_geo)
        assert np.allclose(gb.w_box("matlab")[0], geo.w_box("matlab")[0])
    assert rel_l2(acc, full_atb) < 1e-12
    assert rel_l2(vacc, full_v) < 1e-12


@pytest.mark.parametrize("world,nv,blocksize,n", [(8, 768, 20, 656), (4, 96, 6, 40), (6, 50, 9, 45), (2, 33, 5, 23)])
def test_grid_sharding_covers_every_ray_once(world, nv, blocksize, n):
    """(angle groups) x (detector bands): over all ranks, every (angle, band unit) of every subset is owned
    exactly once, and with the auto grid every rank gets the same number of angles in every subset."""
    from physics_arx_b200 import order_subsets
    from physics_arx_b200.dist import choose_shard, shard_blocks, shard_rows
    blocks = order_subsets(n, blocksize)
    A, R = choose_shard(blocks, world)
    assert A * R == world
    owned = np.zeros((n, nv), dtype=np.int32)
    per_rank = []
    for rank in range(world):
        ra, rr = rank % A, rank // A
        v0, v1 = shard_rows(nv, rr, R) if R > 1 else (0, nv)
        local, ranges = shard_blocks(blocks, ra, A)
        assert sum(c for _, c in ranges) == len(local)
        per_rank.append([c for _, c in ranges])
        owned[np.ix_(local, np.arange(v0, v1))] += 1
    assert (owned == 1).all()
    counts = np.array(per_rank)
    assert (counts == counts[0]).all()          # equal angle counts per subset on every rank
