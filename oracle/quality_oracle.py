"""NumPy restatement of TIGRE's Measure_Quality (RMSE / CC / MSSIM / UQI) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: TIGRE is not vendored under /root/reference and the reference leaves ``Quameasopts`` commented
out (pCT_CBCT_Reconstruction.py:99-103), so these formulas are a recall of TIGRE's utilities (SURVEY.md,
TIGRE-recall).  Written directly on the arrays (two-pass moments, float64), independently of the six-sum form
the device path uses.
"""
from __future__ import annotations

import numpy as np


def measure_quality(res_prev, res, opts):
    a = np.asarray(res_prev, dtype=np.float64).ravel()
    b = np.asarray(res, dtype=np.float64).ravel()
    out = []
    for o in ([opts] if isinstance(opts, str) else list(opts)):
        ma, mb = a.mean(), b.mean()
        va, vb = ((a - ma) ** 2).mean(), ((b - mb) ** 2).mean()
        cov = ((a - ma) * (b - mb)).mean()
        if o == "RMSE":
            out.append(np.sqrt(((a - b) ** 2).mean()))
        elif o == "CC":
            out.append(np.corrcoef(a, b)[0, 1])
        elif o == "MSSIM":
            d = max(a.max(), b.max()) - min(a.min(), b.min())
            c1, c2 = (0.01 * d) ** 2, (0.03 * d) ** 2
            out.append(((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma ** 2 + mb ** 2 + c1) * (va + vb + c2)))
        elif o == "UQI":
            out.append(4 * cov * ma * mb / ((va + vb) * (ma ** 2 + mb ** 2)))
        else:
            raise ValueError(o)
    return np.array(out)


def resize_nd(img, shape, order):
    """skimage.transform.resize(img, shape, order, preserve_range=True, anti_aliasing=False) for a 3-D array as
    scikit-image 0.17.2 (the reference's environment.yaml:124) evaluates it: scipy.ndimage.map_coordinates at
    factor*(o + 1/2) - 1/2 with mode 'mirror' (preprocess/prep_test_data_pseudo_cbct_3D.py:23-36)."""
    from scipy import ndimage as ndi
    img = np.asarray(img)
    if img.dtype.char not in "df":
        img = img.astype(np.float64)
    factors = np.divide(img.shape, shape)
    coords = [factors[i] * (np.arange(d) + 0.5) - 0.5 for i, d in enumerate(shape)]
    cmap = np.array(np.meshgrid(*coords, sparse=False, indexing="ij"))
    return ndi.map_coordinates(img, cmap, order=order, mode="mirror", cval=0.0)
