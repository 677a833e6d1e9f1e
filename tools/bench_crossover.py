"""Where the fan-plane projector stops paying: C2's volume and scan with coarser and coarser detector rows
(same coverage), both kernels on 20 angles.   python tools/bench_crossover.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200.phantom import lung_phantom, make_geometry  # noqa: E402
from physics_arx_b200.projector import Plan  # noqa: E402

d_inplane = 500.0 / 512.0
vol = None
for rows, pitch in ((768, 0.388), (512, 0.582), (384, 0.776), (256, 1.164), (192, 1.552), (128, 2.328)):
    geo = make_geometry((128, 512, 512), (2.5, d_inplane, d_inplane), (rows, 1024), (pitch, 0.388),
                        off_detector=(0.0, -160.0))
    angles = np.linspace(0, 2 * np.pi, 656)[np.linspace(0, 655, 20).astype(int)]
    plan = Plan(geo, angles)
    if vol is None:
        vol = torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda()
    x = plan.pack(vol)
    info = plan.fan_info()
    auto = plan.projector
    res = {}
    for kind in ("fan", "ray"):
        if kind == "fan" and info["rows_per_tile"] < 1:
            continue
        plan.set_projector(kind)
        plan.ax(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.ax(x)
        e1.record()
        torch.cuda.synchronize()
        res[kind] = e0.elapsed_time(e1) / 3
    print(f"rows {rows:4d} pitch {pitch:5.3f} mm: mean run {info['mean_run']:6.1f}, worst-case fit {info['rows_per_tile']:3d} rows, "
          f"auto -> {auto}; " + ", ".join(f"{k} {v:6.2f} ms" for k, v in res.items()))
