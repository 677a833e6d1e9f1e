"""Time the per-subset collective of the sharded OS-SART loop: all-reduce of one resident volume
(C2: 35.9 M floats), and the pieces around it.  torchrun --nproc-per-node N tools/bench_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
n_full = 514 * 514 * 136


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


for frac in (1.0, 0.55, 0.25):
    n = int(n_full * frac)
    x = torch.randn(n, device="cuda")
    t = timed(lambda: dist.all_reduce(x))
    if rank == 0:
        print(f"all_reduce {n * 4 / 1e6:7.1f} MB on {world} ranks: {t:.3f} ms  "
              f"(bus {n * 4 * 2 * (world - 1) / world / t / 1e6:.0f} GB/s)")
    rs = torch.empty(n // world, device="cuda")
    t2 = timed(lambda: dist.reduce_scatter_tensor(rs, x[: n // world * world]))
    t3 = timed(lambda: dist.all_gather_into_tensor(x[: n // world * world], rs))
    if rank == 0:
        print(f"   reduce_scatter {t2:.3f} ms, all_gather {t3:.3f} ms")
dist.destroy_process_group()
