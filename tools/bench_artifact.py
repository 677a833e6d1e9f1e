"""Time the Artifact-Induction kernels (csrc/image_ops.cu) on a planning-CT-sized volume, with the
float64 oracle on a sub-volume beside them.

    python tools/bench_artifact.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu_oracle  # noqa: E402
from physics_arx_b200 import artifact_induction as ai  # noqa: E402
from physics_arx_b200.io_nrrd import NrrdImage  # noqa: E402

rng = np.random.default_rng(0)
pct = NrrdImage(rng.integers(-1000, 1500, size=(128, 512, 512)).astype(np.int16), (0.9766, 0.9766, 2.5),
                (-250.0, -250.0, -160.0))
cbct = NrrdImage(rng.integers(-1000, 1500, size=(264, 384, 384)).astype(np.int16), (1.0, 1.0, 1.0),
                 (-192.0, -192.0, -132.0))
dp, dc = ai.DeviceImage.from_nrrd(pct), ai.DeviceImage.from_nrrd(cbct)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


nvox = dp.data.numel()
t, w1 = timed(lambda: ai.ResampleImages(dp, dc))
print(f"ResampleImages        {t:8.3f} ms  ({nvox * 8 / t / 1e6:.0f} GB/s of 8 B/voxel)")
for a, b in ((0.0, 1.0), (0.5, 0.5), (1.0, 0.0), (0.3, 0.8)):
    t, h = timed(lambda: ai.AdaptiveHistogramMatch(w1, 5, a, b))
    print(f"AdaptiveHistogramMatch a={a} b={b} {t:8.3f} ms  ({nvox * 1331 / t / 1e6:.0f} G window terms/s)")
t, s = timed(lambda: ai.AddImages(dp, h))
print(f"AddImages             {t:8.3f} ms")
t, r = timed(lambda: ai.RescaleImage(s, 0, 1, 1))
print(f"RescaleImage (iso)    {t:8.3f} ms -> {r.size}")
t, _ = timed(lambda: ai.cbct_pct_artifact_adding(dp, dc), reps=2)
print(f"steps 1-4, 7 variations: {t:8.3f} ms per patient")
# CPU baseline: the oracle on a 64^3 corner, scaled by voxel count
sub = w1.data[:64, :64, :64].cpu().numpy()
t0 = time.perf_counter()
cpu_oracle.plahe(sub, 5, 0.5, 0.5)
t1 = time.perf_counter()
print(f"oracle PL-AHE on 64^3: {t1 - t0:.2f} s with {os.cpu_count()} threads -> "
      f"{(t1 - t0) * nvox / sub.size:.0f} s extrapolated to the full volume")
