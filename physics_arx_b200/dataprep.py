"""Output side of the path: what turns the reconstructed sCBCT volumes into the NPZ files the reference's
pix2pix-3D code trains on.

Restates preprocess/prep_test_data_pseudo_cbct_3D.py:23-92 of the reference: every volume of a case is resized
to 128^3 with ``skimage.transform.resize`` (order 0 for the structure masks, order 1 for CT / sCBCT,
``preserve_range``, no anti-aliasing; :23-36), the four structure masks are merged into one label volume
(lungs 1, heart 2, cord 3, oesophagus 4, later labels win in that order; :38-47), CT and CBCT are divided by
their maximum when it exceeds 1 (:54-60) and the three arrays are stored as ``np.savez(CT=, CBCT=,
RTSTRUCTS=)`` (:64), which data/cbct2ct3d_dataset.py:53-58 reads back.  The resize runs on the device
(csrc/image_ops.cu: pax_resize); file handling is io_nrrd instead of pynrrd.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from .projector import _ptr, _stream

EXPECTED = (128, 128, 128)


def resize(img, output_shape, order=1):
    """skimage.transform.resize(img, output_shape, order, preserve_range=True, anti_aliasing=False) of a 3-D
    array, on the device (scikit-image 0.17.2 semantics: pixel-centre aligned coordinates, 'mirror' boundary).
    NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out; float32 result."""
    if not torch.cuda.is_available():
        raise _lib.PaxError("no CUDA device: physics_arx_b200 has no CPU fallback")
    was_np = isinstance(img, np.ndarray)
    dev = img.device if (not was_np and img.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    src = torch.as_tensor(np.ascontiguousarray(img) if was_np else img).to(dev, torch.float32).contiguous()
    if src.ndim != 3 or len(output_shape) != 3:
        raise ValueError("resize expects a 3-D array and a 3-D output shape")
    if order not in (0, 1):
        raise ValueError("order must be 0 (masks) or 1 (images), the two the reference uses")
    d = tuple(int(v) for v in output_shape)
    out = torch.empty(d, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.pax_resize(_ptr(src), *(int(v) for v in src.shape), _ptr(out), *d, int(order), _stream()),
                   "pax_resize")
    return out.cpu().numpy() if was_np else out


def fix_size(img_xyz, is_mask=True, expected_shape=EXPECTED):
    """:23-36 -- an (X,Y,Z) array (what pynrrd returns) -> resized if needed, returned as (Z,Y,X)."""
    order = 0 if is_mask else 1
    img = img_xyz
    if tuple(img.shape) != tuple(expected_shape):
        img = resize(img, expected_shape, order=order)
    return np.swapaxes(img, 0, -1) if isinstance(img, np.ndarray) else img.transpose(0, -1)


def combine_rtstructs(heart, lungs, cord, eso):
    """:38-47 -- BG 0, lungs 1, heart 2, cord 3, oesophagus 4; where masks overlap the later assignment wins."""
    rt = np.zeros(tuple(np.asarray(heart).shape), dtype=np.uint8)
    rt[np.where(np.asarray(eso) == 1)] = 4
    rt[np.where(np.asarray(cord) == 1)] = 3
    rt[np.where(np.asarray(heart) == 1)] = 2
    rt[np.where(np.asarray(lungs) == 1)] = 1
    return rt


def save_case_combined_rtstructs(heart, lungs, cord, eso, cbct, ct, subcase, out_folder):
    """:38-64 -- writes <out_folder>/<subcase>.npz with CT, CBCT, RTSTRUCTS; returns the path."""
    rt = combine_rtstructs(heart, lungs, cord, eso)
    cbct, ct = np.array(cbct, copy=True), np.array(ct, copy=True)
    if cbct.max() > 1:
        cbct /= cbct.max()
    if ct.max() > 1:
        ct /= ct.max()
    out_path = os.path.join(out_folder, "{}".format(subcase))
    np.savez(out_path, CT=ct, CBCT=cbct, RTSTRUCTS=rt)
    return out_path + ".npz"


def process_case(base_dir, case, out_dir):
    """:66-92 -- the files of one case folder, recognised by the reference's substrings."""
    from .io_nrrd import read_nrrd
    got = {}
    for file in sorted(os.listdir(os.path.join(base_dir, case))):
        filename = os.path.join(base_dir, case, file)
        if not file.endswith(".nrrd"):
            continue
        xyz = np.ascontiguousarray(np.transpose(read_nrrd(filename).array, (2, 1, 0)))    # pynrrd's index order
        for key, tag, mask in (("heart", "Heart", True), ("lungs", "Lungs", True), ("cord", "Cord", True),
                               ("eso", "Eso", True), ("gtv", "GTV", True), ("dose", "dose", False),
                               ("ct", "CT_plan_50", False), ("cbct", "_pCT_OSSART_", False)):
            if tag in file:
                got[key] = fix_size(xyz, is_mask=mask)
                break
    missing = [k for k in ("heart", "lungs", "cord", "eso", "cbct", "ct") if k not in got]
    if missing:
        raise FileNotFoundError(f"case {case}: no file for {missing}")
    os.makedirs(out_dir, exist_ok=True)
    return save_case_combined_rtstructs(got["heart"], got["lungs"], got["cord"], got["eso"], got["cbct"],
                                        got["ct"], case, out_dir)
