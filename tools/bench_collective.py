"""The per-subset collective of the sharded OS-SART loop at C2 size, both ways:
  (a) NCCL all-reduce of the partial update + local update kernel (pax_sart_update)
  (b) one fused reduce + update + broadcast kernel over peer memory (pax_sart_update_fused), with the
      two symmetric-memory barriers around it.
torchrun --nproc-per-node N tools/bench_collective.py"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200 import _lib  # noqa: E402
from physics_arx_b200.algorithms.ossart import OSSARTState  # noqa: E402
from physics_arx_b200.phantom import config_c2  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
cfg = config_c2()
geo = cfg["geo"]
angles = cfg["angles"][:: 656 // (2 * world)][: 2 * world]       # a couple of angles per rank: one subset
st = OSSARTState(None, geo, angles, blocksize=len(angles), group=dist.group.WORLD, distributed=True)
p = st.plan
x = st.new_iterate()
st.upd.normal_()
inv_v = torch.rand((p.ny, p.nx), device="cuda")


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def nccl():
    dist.all_reduce(st.upd)
    p.sart_update(x, st.upd, inv_v, 0.5, True)


def fused():
    f = st.fused
    f["hu"].barrier(channel=0)
    _lib.check(p.lib.pax_sart_update_fused(
        p._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(f["mc_x"] or 0), ctypes.c_void_p(f["mc_u"] or 0),
        f["x_peers"], f["upd_peers"], st.rank, st.world, ctypes.c_void_p(inv_v.data_ptr()), 0.5, 1,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "fused")
    f["hx"].barrier(channel=0)


ta = timed(nccl)
if rank == 0:
    print(f"{world} ranks, {st.upd.numel() * 4 / 1e6:.1f} MB update: all-reduce + update kernel {ta:.3f} ms")
if st.fused is None:
    if rank == 0:
        print("fused path not available:", getattr(st, "fused_error", ""))
else:
    st.upd.normal_()
    tb = timed(fused)
    if rank == 0:
        print(f"   fused kernel over {st.fused['kind']} (+ 2 barriers): {tb:.3f} ms")
    only = timed(lambda: st.fused["hu"].barrier(channel=0))
    if rank == 0:
        print(f"   one symmetric-memory barrier alone: {only:.3f} ms")
dist.destroy_process_group()
