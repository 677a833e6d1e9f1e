"""``Quameasopts`` of ``tigre.algorithms.ossart``: per-iteration image quality measures between the iterate
before and after an iteration.  The reference prepares the list ``['RMSE', 'CC', 'MSSIM', 'UQI']`` and leaves the
keyword commented out of the call (pCT_CBCT_Reconstruction.py:99-103); TIGRE's ``Measure_Quality`` is not
vendored, so the formulas below are a recall of it (TIGRE-recall, parity unpinned):

    RMSE  = sqrt(mean((a - b)^2))
    CC    = Pearson correlation of the two volumes
    MSSIM = ((2 ma mb + C1)(2 cov + C2)) / ((ma^2 + mb^2 + C1)(va + vb + C2)),  C_i = (K_i d)^2, K = (0.01, 0.03),
            d = max(a, b) - min(a, b)         (TIGRE evaluates SSIM over the whole volume as one window)
    UQI   = 4 cov ma mb / ((va + vb)(ma^2 + mb^2))

All four are functions of six sums and the joint range, which one pass over the device volumes delivers
(csrc/image_ops.cu: pax_pair_sums, pax_minmax); nothing leaves the GPU but those few doubles.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..projector import _ptr, _stream

MEASURES = ("RMSE", "CC", "MSSIM", "UQI")


def check_opts(opts):
    opts = [opts] if isinstance(opts, str) else list(opts)
    for o in opts:
        if o not in MEASURES:
            raise ValueError(f"unknown quality measure {o!r}; expected a subset of {MEASURES}")
    return opts


def from_sums(sums, lo, hi, opts):
    """The measures from {n, sum a, sum b, sum a^2, sum b^2, sum ab} and the joint (min, max)."""
    n, sa, sb, saa, sbb, sab = (float(v) for v in sums)
    ma, mb = sa / n, sb / n
    va, vb = max(saa / n - ma * ma, 0.0), max(sbb / n - mb * mb, 0.0)       # population moments
    cov = sab / n - ma * mb
    out = []
    for o in check_opts(opts):
        if o == "RMSE":
            out.append(np.sqrt(max(saa - 2.0 * sab + sbb, 0.0) / n))
        elif o == "CC":
            den = np.sqrt(va * vb)
            out.append(cov / den if den > 0 else float("nan"))
        elif o == "MSSIM":
            d = float(hi) - float(lo)
            c1, c2 = (0.01 * d) ** 2, (0.03 * d) ** 2
            out.append(((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2)))
        elif o == "UQI":
            den = (va + vb) * (ma * ma + mb * mb)
            out.append(4 * cov * ma * mb / den if den > 0 else float("nan"))
    return np.array(out, dtype=np.float64)


def Measure_Quality(res_prev, res, QualMeasOpts):
    """Quality measures between two float32 volumes (CUDA tensors or NumPy arrays of equal shape)."""
    opts = check_opts(QualMeasOpts)
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    dev = res.device if isinstance(res, torch.Tensor) and res.is_cuda else torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(res_prev, dtype=torch.float32, device=dev).contiguous()
    b = torch.as_tensor(res, dtype=torch.float32, device=dev).contiguous()
    if a.shape != b.shape:
        raise ValueError(f"shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    lib = _lib.load()
    sums = torch.empty(6, dtype=torch.float64, device=dev)
    mma = torch.empty(2, dtype=torch.float32, device=dev)
    mmb = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.pax_pair_sums(_ptr(a), _ptr(b), a.numel(), _ptr(sums), _stream()), "pax_pair_sums")
        _lib.check(lib.pax_minmax(_ptr(a), a.numel(), _ptr(mma), _stream()), "pax_minmax")
        _lib.check(lib.pax_minmax(_ptr(b), b.numel(), _ptr(mmb), _stream()), "pax_minmax")
    mm = torch.stack([mma, mmb]).cpu().numpy()
    return from_sums(sums.cpu().numpy(), mm[:, 0].min(), mm[:, 1].max(), opts)
