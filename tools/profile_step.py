"""One OS-SART subset step (Ax residual + Atb update) at full C2 size, for ncu captures.

    python tools/profile_step.py [--config C2] [--subsets 2] [--reps 2]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physics_arx_b200 import _lib  # noqa: E402
from physics_arx_b200.algorithms.ossart import OSSARTState  # noqa: E402
from physics_arx_b200.phantom import config_c1, config_c2, lung_phantom  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--subsets", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--compare", action="store_true", help="also run the other projector kernel and report the difference")
args = ap.parse_args()
cfg = {"C1": config_c1, "C2": config_c2}[args.config]()
geo, angles = cfg["geo"], cfg["angles"]
n = args.subsets * cfg["blocksize"]
angles = angles[np.linspace(0, len(angles) - 1, n).astype(int)]      # spread over the rotation
st = OSSARTState(None, geo, angles, cfg["blocksize"])
p = st.plan
vol = torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda()
x = p.pack(vol)
p.ax(x, 0, n, out=st.b)                  # "measured" projections of the phantom
x.mul_(0.9)                              # a non-trivial residual
torch.cuda.synchronize()
print("projector:", p.projector, p.fan_info())
if args.compare:
    first, count = st.ranges[0]
    kinds = ("fan", "ray") if p.fan_info()["rows_per_tile"] >= 1 else ("ray",)
    outs = {}
    for kind in kinds:
        p.set_projector(kind)
        ns = torch.zeros(1, dtype=torch.int64, device="cuda")
        o = p.ax(x, first, count, nsamples=ns)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = p.ax(x, first, count)
        e1.record()
        torch.cuda.synchronize()
        outs[kind] = o.double()
        print(f"{kind}: {e0.elapsed_time(e1):.3f} ms for {count} angles, {int(ns.item())} samples, "
              f"{int(ns.item()) / e0.elapsed_time(e1) / 1e6:.1f} Gsamples/s")
    if len(outs) == 2:
        d = outs["fan"] - outs["ray"]
        print(f"fan vs ray: rel L2 {float(d.norm() / outs['ray'].norm()):.3e}, max abs {float(d.abs().max()):.3e} "
              f"(max value {float(outs['ray'].max()):.3f})")
    print("fan fault flag:", p.fault())
    p.set_projector(os.environ.get("PAX_AX_KERNEL", "auto") if os.environ.get("PAX_AX_KERNEL") in ("fan", "ray") else "auto")
for r in range(args.reps):
    for s in range(len(st.blocks)):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        first, count = st.ranges[s]
        e0.record()
        p.ax_residual(x, st.b[first:first + count], first, count, st.resid)
        e1.record()
        p.atb(st.resid, first, count, x, _lib.PAX_ATB_SART, st.inv_v[s], 0.5, True)
        e2.record()
        torch.cuda.synchronize()
        print(f"rep {r} subset {s}: ax_residual {e0.elapsed_time(e1):.3f} ms, atb_sart {e1.elapsed_time(e2):.3f} ms")
