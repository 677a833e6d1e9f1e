"""Timing of BASELINE.json's other configurations on one B200 (they are parity-test cases, not the
headline): C1 (128^3, 256 angles, 5 iterations), C4 (32 variations of one CT at 128^3 from ONE two-channel
projection pass), C5 (10-phase motion scan + reconstruction).   python tools/bench_configs.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200.algorithms import ossart  # noqa: E402
from physics_arx_b200.drivers import (VariationBatch, config_c4, motion_projections, phase_volumes,  # noqa: E402
                                      respiratory_phases)
from physics_arx_b200.phantom import artifact_volume, config_c1, lung_phantom  # noqa: E402
from physics_arx_b200.pipeline import SCBCTPipeline  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cfg = config_c1()
geo, angles = cfg["geo"], cfg["angles"]
ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
art = artifact_volume(geo.nVoxel, seed=1)
ct_d, art_d = torch.from_numpy(ct).cuda(), torch.from_numpy(art).cuda()

pipe = SCBCTPipeline(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], I0=cfg["I0"], sigma=cfg["sigma"],
                     scatter_scale=0.5, seed=0)
t = timed(lambda: pipe.run_device(ct_d, art_d))
print(f"C1  one sCBCT volume (projector: {pipe.plan.projector}): {t:8.1f} ms")

c4 = config_c4()
vb = VariationBatch(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], sigma=cfg["sigma"])
n = len(c4["scatter_scales"]) * len(c4["I0s"]) * len(c4["strides"])
t = timed(lambda: vb.run(ct_d, art_d, c4["scatter_scales"], c4["I0s"], c4["strides"], to_numpy=False), reps=2)
print(f"C4  {n} variations from one two-channel projection pass: {t:8.1f} ms = {t / n:.1f} ms per variation "
      f"({1e3 * n / t:.1f} volumes/s)")

ph = respiratory_phases(len(angles))
pv = torch.from_numpy(phase_volumes(ct, float(geo.dVoxel[0]))).cuda()


def c5():
    proj = motion_projections(pv, ph, geo, angles)
    return ossart(proj, geo, angles, cfg["niter"], blocksize=cfg["blocksize"])


t = timed(c5)
print(f"C5  10-phase motion scan (256 angles) + {cfg['niter']}-iteration OS-SART: {t:8.1f} ms")
