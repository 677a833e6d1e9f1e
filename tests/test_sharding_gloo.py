"""World-size-2 gloo run of the multi-GPU host logic on CPU: each rank takes
its interleaved share of every OS-SART subset, backprojects it with the oracle, and the
all-reduced partial updates / V maps must equal the unsharded result."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_oracle as O
    from physics_arx_b200 import order_subsets
    from physics_arx_b200.dist import shard_blocks, subset_geometry
    from physics_arx_b200.phantom import make_geometry
    geo = make_geometry((6, 14, 14), (12.0, 9.0, 9.0), (10, 24), (9.0, 6.0), off_detector=(0.0, -20.0))
    n = 11
    geo.DSD = np.linspace(1480.0, 1520.0, n)            # per-angle attribute must follow the shard
    angles = np.linspace(0, 2 * np.pi, n)
    rng = np.random.default_rng(0)
    resid = rng.standard_normal((n, 10, 24))
    blocks = order_subsets(n, 4)
    local, ranges = shard_blocks(blocks, rank, world)
    ok = True
    for s, (first, count) in enumerate(ranges):
        idx = local[first:first + count]
        if count:
            gs = subset_geometry(geo, idx, n)
            part = O.atb(resid[idx], gs, angles[idx])
            vpart = O.atb(None, gs, angles[idx], vscale=geo.v_scale()).mean(axis=0)
        else:
            part = np.zeros((6, 14, 14))
            vpart = np.zeros((14, 14))
        t = torch.from_numpy(part)
        v = torch.from_numpy(vpart)
        dist.all_reduce(t)
        dist.all_reduce(v)
        gfull = subset_geometry(geo, blocks[s], n)
        full = O.atb(resid[blocks[s]], gfull, angles[blocks[s]])
        vfull = O.compute_v(geo, angles, blocks, "A")[s]
        ok &= bool(np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-12))
        ok &= bool(np.allclose(v.numpy(), vfull, rtol=1e-12, atol=1e-12))
    # noise depends on the global pixel index only: shards concatenate to the full result
    p = rng.random((n, 10, 24)) * 50
    full_noise = O.ctnoise_add(p, Poisson=1e6, Gaussian=np.array([0.0, 2.0]), seed=5, maxval=50.0)
    mine = np.concatenate([O.ctnoise_add(p[i], Poisson=1e6, Gaussian=np.array([0.0, 2.0]), seed=5,
                                         maxval=50.0, index0=int(i) * 240) for i in local])
    ok &= bool(np.array_equal(mine.reshape(-1, 10, 24), full_noise[local]))
    m = torch.tensor([float(p[local].max())], dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    ok &= bool(m.item() == p.max())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_sharded_update_equals_unsharded():
    import torch.multiprocessing as mp
    from oracle import cpu_oracle
    cpu_oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
