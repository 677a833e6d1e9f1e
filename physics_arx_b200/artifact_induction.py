"""Artifact induction on the device: the reference's ITK command-line tools and the shell script
that chains them (Physics-based-Augmentation/Artifact-Induction/Cpp-Codes/*/src/*.cxx,
CBCT_pCT_Artifact_adding.sh:45-145), with the same names, argument meaning and outputs, but
without the NRRD round trip between every two steps.

    ResampleImages(ref, img)            ResampleImages.cxx        short pixels, linear, identity transform
    ResampleImagesFloat(ref, img)       ResampleImagesFloat.cxx   the same on float pixels
    AdaptiveHistogramMatch(img, r, a, b) AdaptiveHistogramMatch.cxx:30-42   PL-AHE (ITK adaptive histogram eq.)
    AddImages(img1, img2)               AddImages.cxx:40-62       img1 + resample(img2 -> img1 grid)
    RescaleImage(img, lo, hi, is_iso)   RescaleImage.cxx:44-107   iso-z resample, then rescale to [lo,hi]
    cbct_pct_artifact_adding(pCT, CBCT) CBCT_pCT_Artifact_adding.sh:45-145, steps 1-4

Images are ``NrrdImage`` (array (z,y,x) + spacing/origin/direction in (x,y,z) order, io_nrrd.py); arrays
live on the GPU between steps (``DeviceImage``).  All arithmetic is in libpax.so (csrc/image_ops.cu)
through the C ABI of include/pax.h; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .io_nrrd import NrrdImage
from .projector import _ptr, _stream, lincomb

# (alpha, beta) of the seven variations the script extracts and adds (CBCT_pCT_Artifact_adding.sh:57-111;
# the eighth, (1,1), is computed at :51-54 but never added: PL-AHE with alpha = beta = 1 is the identity)
VARIATIONS = ((0.0, 1.0), (1.0, 0.0), (0.5, 0.5), (0.5, 0.0), (0.0, 0.5), (0.5, 1.0), (1.0, 0.5))
RADIUS = 5


class DeviceImage:
    """(Z,Y,X) float32 CUDA tensor + the geometry an itk::Image carries."""

    def __init__(self, data, spacing, origin, direction=None):
        if not torch.cuda.is_available():
            raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
        self.data = data.to(torch.float32).contiguous()
        self.spacing = tuple(float(v) for v in spacing)
        self.origin = tuple(float(v) for v in origin)
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)

    @classmethod
    def from_nrrd(cls, img: NrrdImage):
        return cls(np.asarray(img.array, dtype=np.float32), img.spacing, img.origin, img.direction)

    def to_nrrd(self, dtype=np.float32) -> NrrdImage:
        return NrrdImage(self.data.cpu().numpy().astype(dtype), self.spacing, self.origin, self.direction)

    @property
    def size(self):                      # (x,y,z), like itk::Image::GetSize / sitk.GetSize
        return tuple(int(v) for v in self.data.shape[::-1])

    def like(self, data):
        return DeviceImage(data, self.spacing, self.origin, self.direction)


def _dev(img):
    return img if isinstance(img, DeviceImage) else DeviceImage.from_nrrd(img)


def index_map(dst_origin, dst_spacing, dst_direction, src_origin, src_spacing, src_direction):
    """12 doubles: continuous SOURCE index (x,y,z) of DESTINATION index (i,j,k) under the identity
    transform -- what itk::ResampleImageFilter evaluates per output voxel."""
    dd = np.asarray(dst_direction, dtype=np.float64).reshape(3, 3)
    sd = np.asarray(src_direction, dtype=np.float64).reshape(3, 3)
    a = np.linalg.inv(sd @ np.diag(np.asarray(src_spacing, dtype=np.float64)))
    lin = a @ dd @ np.diag(np.asarray(dst_spacing, dtype=np.float64))
    off = a @ (np.asarray(dst_origin, dtype=np.float64) - np.asarray(src_origin, dtype=np.float64))
    return np.ascontiguousarray(np.concatenate([lin, off[:, None]], axis=1).reshape(-1))


def _resample(src: DeviceImage, size_xyz, spacing, origin, direction, cast_short, default=0.0):
    lib = _lib.load()
    m = index_map(origin, spacing, direction, src.origin, src.spacing, src.direction)
    nx, ny, nz = (int(v) for v in size_xyz)
    out = torch.empty((nz, ny, nx), dtype=torch.float32, device=src.data.device)
    snz, sny, snx = (int(v) for v in src.data.shape)
    with torch.cuda.device(src.data.device):
        _lib.check(lib.pax_resample_linear(_ptr(src.data), snz, sny, snx, m.ctypes.data_as(_lib.c_double_p),
                                           _ptr(out), nz, ny, nx, float(default), int(bool(cast_short)),
                                           _stream()), "pax_resample_linear")
    return DeviceImage(out, spacing, origin, direction)


def ResampleImages(reference, image):
    """ResampleImages.cxx:87-102: `image` on the grid of `reference`; short pixels (the double result
    of the linear interpolation is truncated toward zero, ITK's static_cast), 0 outside."""
    ref, img = _dev(reference), _dev(image)
    return _resample(img, ref.size, ref.spacing, ref.origin, ref.direction, cast_short=True)


def ResampleImagesFloat(reference, image):
    """ResampleImagesFloat.cxx: the same on float pixels."""
    ref, img = _dev(reference), _dev(image)
    return _resample(img, ref.size, ref.spacing, ref.origin, ref.direction, cast_short=False)


def AddImages(image1, image2):
    """AddImages.cxx:40-62: image1 + (image2 resampled onto image1's grid, linear, 0 outside)."""
    a, b = _dev(image1), _dev(image2)
    b_on_a = _resample(b, a.size, a.spacing, a.origin, a.direction, cast_short=False)
    out = torch.empty_like(a.data)
    lincomb(out, a.data, b_on_a.data, 1.0, 1.0)
    return a.like(out)


def _minmax(t):
    lib = _lib.load()
    mm = torch.empty(2, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.pax_minmax(_ptr(t), t.numel(), _ptr(mm), _stream()), "pax_minmax")
    return mm


def RescaleImage(image, out_min=0, out_max=1, is_iso=1):
    """RescaleImage.cxx:44-107.  With is_iso == 1 (what the script always passes) the image is first
    resampled to slices as thick as the in-plane pixel (spacing z := spacing x, size z :=
    int(nz * sz / sx), linear, 0 outside), then rescaled to [out_min, out_max] with
    itk::RescaleIntensityImageFilter.  The reference tool feeds an EMPTY image to the rescaler when
    is_iso != 1 (:36-41 never assigns it); here that case rescales the input as it is."""
    lib = _lib.load()
    img = _dev(image)
    if int(is_iso) == 1:
        sx, sy, sz = img.spacing
        nx, ny, nz = img.size
        size = (int(nx * sx / sx), int(ny * sy / sy), int(nz * sz / sx))
        img = _resample(img, size, (sx, sy, sx), img.origin, img.direction, cast_short=False)
    else:
        img = img.like(img.data.clone())
    mm = _minmax(img.data)
    with torch.cuda.device(img.data.device):
        _lib.check(lib.pax_rescale_intensity(_ptr(img.data), img.data.numel(), _ptr(mm), float(int(out_min)),
                                             float(int(out_max)), _stream()), "pax_rescale_intensity")
    return img


def AdaptiveHistogramMatch(image, radius=RADIUS, alpha=0.5, beta=0.5):
    """AdaptiveHistogramMatch.cxx:30-42: itk::AdaptiveHistogramEqualizationImageFilter (power-law adaptive
    histogram equalisation) with a cubic window of the given radius."""
    lib = _lib.load()
    img = _dev(image)
    nz, ny, nx = (int(v) for v in img.data.shape)
    out = torch.empty_like(img.data)
    mm = _minmax(img.data)
    with torch.cuda.device(img.data.device):
        _lib.check(lib.pax_plahe(_ptr(img.data), nz, ny, nx, int(radius), float(alpha), float(beta), _ptr(mm),
                                 _ptr(out), _stream()), "pax_plahe")
    return img.like(out)


def _tag(v):
    return ("%g" % v)


def cbct_pct_artifact_adding(plan_ct, cbct, variations=VARIATIONS, radius=RADIUS):
    """Steps 1-4 of CBCT_pCT_Artifact_adding.sh for one patient, device resident.

    Returns a dict keyed like the script's output files (without the patient prefix):
    ``CBCT_w1``, ``CBCT_Histogram_a_b``, ``pCT_Artifact_a_b``, ``CBCT_rsc_w1``, ``rsc_pCT`` and
    ``pCT_Artifact_a_b_rsc`` -- the last being what pCT_CBCT_Reconstruction.py:40-56 reads.
    """
    pct = _dev(plan_ct)
    out = {}
    # 1) week-1 CBCT onto the planning-CT grid (short pixels)                          .sh:45-48
    cbct_w1 = ResampleImages(pct, cbct)
    out["CBCT_w1"] = cbct_w1
    # 2) PL-AHE "scatter" images for the (alpha, beta) variations                      .sh:51-87
    # 3) add them to the planning CT                                                   .sh:92-111
    for a, b in variations:
        key = f"{_tag(a)}_{_tag(b)}"
        hist = AdaptiveHistogramMatch(cbct_w1, radius, a, b)
        out[f"CBCT_Histogram_{key}"] = hist
        out[f"pCT_Artifact_{key}"] = AddImages(pct, hist)
    # 4) rescale everything to [0,1] on iso-z grids                                    .sh:115-145
    out["CBCT_rsc_w1"] = RescaleImage(cbct_w1, 0, 1, 1)
    out["rsc_pCT"] = ResampleImagesFloat(out["CBCT_rsc_w1"], RescaleImage(pct, 0, 1, 1))
    for a, b in variations:
        key = f"{_tag(a)}_{_tag(b)}"
        out[f"pCT_Artifact_{key}_rsc"] = RescaleImage(out[f"pCT_Artifact_{key}"], 0, 1, 1)
    return out


def write_artifact_volumes(volumes, out_dir, patient):
    """Write the dict of cbct_pct_artifact_adding() under <out_dir>/<patient>/Artifacts/ with the script's
    file names (CBCT_pCT_Artifact_adding.sh:36-38,45-145), so that reconstruct_cases.pCT_CBCT_reconstruction
    (and the reference's own script) find the seven `*_rsc.nrrd` inputs."""
    import os

    from .io_nrrd import write_nrrd
    art = os.path.join(out_dir, patient, "Artifacts")
    os.makedirs(art, exist_ok=True)
    os.makedirs(os.path.join(out_dir, patient, "Recons"), exist_ok=True)
    paths = {}
    for key, img in volumes.items():
        dtype = np.int16 if key == "CBCT_w1" else np.float32          # ResampleImages writes short pixels
        path = os.path.join(art, f"{patient}_{key}.nrrd")
        write_nrrd(path, img.data.cpu().numpy().astype(dtype), img.spacing, img.origin, img.direction)
        paths[key] = path
    return paths
