"""`pCT_CBCT_reconstruction(in_dir)` -- the reference's driver script, same behaviour, on libpax.

Restates Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py:18-110:
for every patient folder, for each of the seven artifact-added volumes `*_{a}_{b}_rsc.nrrd`
(:44-45): read NRRD -> float32 (z,y,x) (:60-63); geometry DSD 1500, 400x400 detector over 400 mm,
voxel grid from image size x spacing (:67-78); 500 angles over [0, 2 pi] (:68); Ax 'interpolated'
(:86); CTnoise.add(Poisson=1e8, Gaussian=[0,4]) (:88); save the noisy projections as NPZ (:91-92);
ossart(niter=30, blocksize=20) (:101-103); write the NRRD with the source's Direction / Spacing /
Origin (:105-110).  Differences: volumes stay on the device between the steps, NRRD I/O is
io_nrrd (SimpleITK is not required), and the function does not run at import.
"""
from __future__ import annotations

import os

import numpy as np

FNAMES = sorted(['0.5_0.5_rsc.nrrd', '0.5_0_rsc.nrrd', '0.5_1_rsc.nrrd', '0_0.5_rsc.nrrd', '0_1_rsc.nrrd',
                 '1_0.5_rsc.nrrd', '1_0_rsc.nrrd'])


def case_geometry(size_xyz, spacing_xyz, half_fan=False):
    """The geometry the reference builds for one volume (:67-83)."""
    from .geometry import geometry_default
    geo = geometry_default(high_resolution=False)
    geo.DSD = 1500
    geo.nDetector = np.array((400, 400))
    geo.sDetector = np.array((400, 400))
    geo.dDetector = geo.sDetector / geo.nDetector
    sz, spc = size_xyz, spacing_xyz
    geo.nVoxel = np.array((sz[2], sz[1], sz[0]))
    geo.sVoxel = np.array((sz[2] * float(spc[2]), sz[1] * float(spc[1]), sz[0] * float(spc[0])))
    geo.dVoxel = geo.sVoxel / geo.nVoxel
    if half_fan:                                   # the commented-out hook at :83
        geo.offDetector = np.array((0, -160))
    return geo


def list_cases(in_dir):
    """[(patient, caseID, path)] in the reference's order (:34-56)."""
    out = []
    for patient in sorted(os.listdir(in_dir)):
        pth = os.path.join(in_dir, patient, 'Artifacts')
        if not os.path.isdir(pth):
            continue
        artifacts = sorted(f for f in os.listdir(pth) if any(sfx in f for sfx in FNAMES))
        for file, artifact in zip(FNAMES, artifacts):
            out.append((patient, file[:-9], os.path.join(pth, artifact)))
    return out


def pCT_CBCT_reconstruction(in_dir, niter=30, blocksize=20, n_angles=500, Poisson=1e8,
                            Gaussian=np.array([0, 4]), seed=0, half_fan=False, save_projections=True,
                            verbose=True, w_variant=None, v_variant=None):
    """The reference script's loop (module docstring).  ``w_variant`` / ``v_variant`` choose between the W / V
    definitions of TIGRE's builds (the script's TIGRE checkout is un-pinned).  Default: those of
    the mid-2021 Python package the script was written against (``'python2021'`` / ``'B'``), as
    ``compat.install()`` gives the script itself; with ``half_fan`` current upstream's (``'matlab'`` / ``'A'``:
    MATLAB ``computeW``, V from ``Atb(ones)``), which an offset detector needs."""
    import torch

    from . import Ax
    from .algorithms import ossart
    from .io_nrrd import read_nrrd, write_nrrd
    from .utilities import CTnoise, gpu

    if w_variant is None:
        w_variant = "matlab" if half_fan else "python2021"
    if v_variant is None:
        v_variant = "A" if half_fan else "B"
    names = gpu.getGpuNames()
    if len(names) == 0:
        raise RuntimeError("Error: No gpu found")          # the reference prints this and then fails (:25-31)
    gpuids = gpu.getGpuIds(names[0])
    done = []
    for patient, caseID, infile in list_cases(in_dir):
        img = read_nrrd(infile)
        vol = torch.from_numpy(np.array(img.array, dtype=np.float32, order='C')).cuda()
        geo = case_geometry(img.GetSize(), img.GetSpacing(), half_fan)
        angles = np.linspace(0, 2 * np.pi, n_angles)
        projections = Ax(vol, geo, angles, 'interpolated', gpuids=gpuids)
        noise_projections = CTnoise.add(projections, Poisson=Poisson, Gaussian=Gaussian, seed=seed, inplace=True)
        if verbose:
            print(tuple(noise_projections.shape), float(noise_projections.min()), float(noise_projections.max()))
        if save_projections:
            np.savez(os.path.join(in_dir, patient, 'Artifacts', patient + '_projections_' + caseID + '.npz'),
                     projection=noise_projections.cpu().numpy())
        out_dir = os.path.join(in_dir, patient, 'Recons')
        os.makedirs(os.path.join(out_dir, 'QualMeasure'), exist_ok=True)
        rec = ossart(noise_projections, geo, angles, niter, blocksize=blocksize, gpuids=gpuids,
                     w_variant=w_variant, v_variant=v_variant)
        filename = os.path.join(out_dir, patient + '_pCT_OSSART_' + caseID + '.nrrd')
        write_nrrd(filename, rec.cpu().numpy(), img.GetSpacing(), img.GetOrigin(), img.direction, img.space)
        done.append(filename)
    return done


if __name__ == "__main__":
    import sys
    pCT_CBCT_reconstruction(sys.argv[1] if len(sys.argv) > 1 else '../Sample-Patients')
