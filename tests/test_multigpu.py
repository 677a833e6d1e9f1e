"""N-GPU angle-sharded run vs the 1-GPU run on the same inputs: the only
difference allowed is the summation order of the all-reduce.  Needs >= 2 GPUs (gpurun --gpus 2);
skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, shard):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from physics_arx_b200.phantom import artifact_volume, lung_phantom, make_geometry
    from physics_arx_b200.pipeline import SCBCTPipeline
    geo = make_geometry((16, 40, 40), (10.0, 6.0, 6.0), (40, 72), (6.0, 4.0), off_detector=(0.0, -30.0))
    if shard in ("tiles", "gather_tiles"):  # fine rows over thick slices: AUTO picks the fan-plane projector, and
        # the bands of a grid become shares of its column tiles (OSSARTState.tiles)
        geo = make_geometry((16, 40, 40), (10.0, 4.0, 4.0), (160, 72), (1.5, 4.0), off_detector=(0.0, -30.0))
    angles = np.linspace(0, 2 * np.pi, 23)
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    kw = dict(niter=3, blocksize=5, I0=1e8, sigma=4.0, scatter_scale=0.5, seed=3)
    want = "nccl"
    os.environ["PAX_FUSED_UPDATE"] = "0"    # the plain modes below exercise NCCL all-reduce + local update
    if shard == "p2p":                      # the peer-pointer variant of the fused update kernel
        os.environ["PAX_FUSED_UPDATE"] = "p2p"
        shard, want = "angles", "peer pointers"
    elif shard == "fused":                  # the fused kernel, multicast when the fabric has it
        os.environ["PAX_FUSED_UPDATE"] = "1"
        shard, want = "grid", "fused"
    elif shard == "tiles":
        del os.environ["PAX_FUSED_UPDATE"]
        shard, want = (1, world), "fused"
    elif shard == "gather_tiles":           # the gather scheme where the auto grid has column-tile shares (23 angles in
        del os.environ["PAX_FUSED_UPDATE"]  # blocks of 5 share no factor with the rank count: 1 x world)
        os.environ["PAX_MGPU_GRID"] = "auto"
        shard, want = "auto", "gather"
    elif shard == "default":                # what a caller gets: the gather scheme (residual all-gather, row-owned
        del os.environ["PAX_FUSED_UPDATE"]  # backprojection + update, row broadcast)
        shard, want = "auto", "gather"
    if shard == "grid":                     # 2 angle groups x (world/2) bands of detector columns
        shard = (2, world // 2)
    pipe = SCBCTPipeline(geo, angles, group=dist.group.WORLD, shard=shard, **kw)
    if want == "fused" and isinstance(shard, tuple) and shard[0] == 1:
        assert pipe.st.tiles == (dist.get_rank(), world) and pipe.plan.projector == "fan"
    elif shard != "auto":
        assert pipe.st.shard == (shard if isinstance(shard, str) else ("grid" if world > 2 else "angles"))
    sharded = pipe.run_host(ct, art).numpy().copy()
    coll = "nccl" if pipe.st.fused is None else pipe.st.fused["kind"]
    if want == "gather":
        assert pipe.st.scheme == "gather", getattr(pipe.st, "gather_error", "")
        coll = "gather, " + pipe.st.gsym["kind"] + f", grid {pipe.st.grid}, tiles {pipe.st.tiles}"
        if os.environ.get("PAX_MGPU_GRID") == "auto":
            assert pipe.plan.projector == "fan" and pipe.st.grid == (1, world)
            assert pipe.st.tiles == (dist.get_rank(), world)
        else:
            assert pipe.st.grid == (world, 1) and pipe.st.tiles is None
    elif want == "fused":
        assert pipe.st.fused is not None, getattr(pipe.st, "fused_error", "")
    else:
        assert coll == want, (coll, getattr(pipe.st, "fused_error", ""))
    ok, info = True, ""
    if rank == 0:
        single = SCBCTPipeline(geo, angles, **kw).run_host(ct, art).numpy()
        hu = np.abs(sharded - single) * 4095.0
        ok = bool(hu.mean() <= 0.05 and hu.max() <= 0.5)
        info = f"mean {hu.mean():.2e} HU max {hu.max():.2e} HU, collective: {coll}, " \
               f"{getattr(pipe.st, 'fused_error', '')}"
    # every rank must hold the same volume
    t = torch.from_numpy(sharded).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    ok &= bool(torch.equal(t, ref))
    q.put((rank, ok, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shard", ["angles", "rows", "cols", "grid", "p2p", "fused", "default", "tiles", "gather_tiles"])
def test_sharded_pipeline_equals_single_gpu(shard):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, shard)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    print(sorted(res)[0][2])


def test_gpuids_with_several_devices_in_one_process():
    """The reference passes EVERY GPU of the first name as ``gpuids`` (pCT_CBCT_Reconstruction.py:31,86,103);
    such a call is split over the devices inside one process (multidev.MultiDevice) and must give what one
    device gives."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import physics_arx_b200 as pax
    from physics_arx_b200.phantom import lung_phantom, make_geometry
    from physics_arx_b200.utilities import gpu
    ids = gpu.getGpuIds(gpu.getGpuNames()[0])
    assert len(ids) >= 2
    one = gpu.GpuIds(ids.name, ids.devices[:1])
    geo = make_geometry((16, 40, 40), (10.0, 4.0, 4.0), (120, 72), (1.5, 4.0), off_detector=(0.0, -30.0))
    angles = np.linspace(0, 2 * np.pi, 23)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    p1 = pax.Ax(vol, geo, angles, "interpolated", gpuids=one)
    pn = pax.Ax(vol, geo, angles, "interpolated", gpuids=ids)
    assert pn.shape == p1.shape and np.array_equal(pn, p1)          # same kernel, same angles: bit-identical
    b1 = pax.Atb(p1, geo, angles, "FDK", gpuids=one)
    bn = pax.Atb(p1, geo, angles, "FDK", gpuids=ids)
    assert np.abs(bn - b1).max() <= 1e-5 * np.abs(b1).max()          # summation order over devices only
    r1 = pax.algorithms.ossart(p1, geo, angles, 3, blocksize=5, gpuids=one)
    rn, l2 = pax.algorithms.ossart(p1, geo, angles, 3, blocksize=5, gpuids=ids, computel2=True)
    hu = np.abs(rn - r1) * 4095.0
    print(f"{len(ids)} devices in one process vs one: mean {hu.mean():.2e} HU, max {hu.max():.2e} HU")
    assert hu.mean() <= 0.05 and hu.max() <= 0.5
    assert len(l2) == 3 and l2[-1] < l2[0]
