"""Ax + residual of a few angles of C2 (what a rank of a multi-GPU run launches per subset), best of 12.

    PAX_FAN_QUEUED=0|1|2 python tools/time_small_launch.py      (0: grid launch, 1: queue for unsplit launches, 2: for all)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200.phantom import config_c2, lung_phantom  # noqa: E402
from physics_arx_b200.projector import Plan  # noqa: E402

cfg = config_c2()
geo, angles = cfg["geo"], cfg["angles"]
vol = torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda()
line = []
for count in (2, 3, 5, 10, 20):
    p = Plan(geo, angles[100:100 + count])
    x = p.pack(vol)
    b = p.ax(x)
    x.mul_(0.9)
    out = torch.empty_like(b)
    best = 1e30
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.ax_residual(x, b, 0, count, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    line.append(f"{count} angles {best:.3f} ms")
    chk = float(out.double().abs().sum())
print("; ".join(line), f"(checksum of the last {chk:.6e})")
