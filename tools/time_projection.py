"""The simulation pass of C2 (all 656 angles, plain mode) against the same angles projected 20 at a time.

    python tools/time_projection.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200.phantom import config_c2, lung_phantom  # noqa: E402
from physics_arx_b200.projector import Plan  # noqa: E402

cfg = config_c2()
geo, angles = cfg["geo"], cfg["angles"]
p = Plan(geo, angles)
x = p.pack(torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda())
n = len(angles)
out = torch.empty((n, p.nv, p.nu), dtype=torch.float32, device="cuda")


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def by20():
    for f in range(0, n, 20):
        c = min(20, n - f)
        p.ax(x, f, c, out=out[f:f + c])


whole = timed(lambda: p.ax(x, 0, n, out=out))
ref = out.clone()
parts = timed(by20)
print(f"{n} angles in one call: {whole:.1f} ms; 20 at a time: {parts:.1f} ms; identical: {bool(torch.equal(ref, out))}")
