"""``tigre.utilities.CTnoise.add`` (reference call site pCT_CBCT_Reconstruction.py:88;
algorithm SURVEY.md A.7):  I = I0 exp(-p/max p);  I <- Poisson(I) + N(mean, std);
p' = -log(I/I0) max p.

The reference draws from NumPy's unseeded global RNG, so no bit-level parity with it
exists; this build specifies its own stream -- one Philox4x32-10 block per pixel keyed by
(seed, global pixel index) -- which makes the result independent of how projections are
sharded over GPUs.  Arithmetic is double on the device (I0 = 1e8 counts do not fit fp32).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib
from ..projector import _ptr, _stream, _to_cuda, reduce_max


def add(projections, Poisson=1e5, Gaussian=np.array([0, 0.5]), seed=0, index0=0, group=None,
        inplace=False):
    """Add Poisson + Gaussian noise; NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out.

    ``index0``: global element index of ``projections[0,0,0]`` (for angle-sharded callers);
    ``group``: torch.distributed group over which the normalising max is taken.
    """
    if np.ndim(Poisson) != 0:
        raise ValueError("Poisson must be a scalar (max photon count I0)")
    Gaussian = np.asarray(Gaussian, dtype=np.float64)
    if Gaussian.shape != (2,):
        raise ValueError("Gaussian must be an array of shape (2,): (mean, std)")
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    p, was_np = _to_cuda(projections, torch.device("cuda", torch.cuda.current_device()), "projections")
    if not was_np and not inplace:
        p = p.clone()
    lib = _lib.load()
    m = reduce_max(p)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    with torch.cuda.device(p.device):
        _lib.check(lib.pax_ctnoise(_ptr(p), p.numel(), ctypes.c_uint64(int(index0)), 0, 0, float(Poisson),
                                   float(Gaussian[0]), float(Gaussian[1]), _ptr(m),
                                   ctypes.c_uint64(int(seed)), _stream()), "pax_ctnoise")
    return p.cpu().numpy() if was_np else p
