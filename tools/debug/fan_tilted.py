import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from oracle import cpu_oracle
from physics_arx_b200.projector import Plan
from test_gpu_parity import _fan_cases
name, geo, angles = _fan_cases()[3]
rng = np.random.default_rng(11)
vol = rng.random(tuple(int(v) for v in geo.nVoxel)).astype(np.float32)
plan = Plan(geo, angles, projector="fan")
print(plan.fan_info())
res = plan.pack(torch.from_numpy(vol).cuda())
got = plan.ax(res).cpu().numpy().astype(np.float64)
print("fault", plan.fault())
ref = cpu_oracle.ax(vol, geo, angles)
d = np.abs(got - ref)
print("rel", np.linalg.norm(got - ref) / np.linalg.norm(ref))
for a in range(len(angles)):
    print("angle", a, "max", d[a].max(), "at", np.unravel_index(d[a].argmax(), d[a].shape), "rel", np.linalg.norm(d[a]) / np.linalg.norm(ref[a]))
bad = np.argwhere(d > 1e-4 * ref.max())
print("n bad", len(bad), "of", d.size)
print(bad[:40])
rows = np.unique(bad[:, 1]); print("bad rows", rows)
cols = np.unique(bad[:, 2]); print("bad cols", cols)
for (a, v, u) in bad[:10]:
    print(a, v, u, got[a, v, u], ref[a, v, u])
