"""Minimal NRRD (NRRD0004/5) reader / writer for the volumes the path consumes and produces.

The reference reads `*_rsc.nrrd` with SimpleITK (pCT_CBCT_Reconstruction.py:60-63:
ReadImage -> GetArrayFromImage (z,y,x), GetSize (x,y,z), GetSpacing) and writes the
reconstruction with the source's Direction / Spacing / Origin copied over (:105-110).  SimpleITK
and pynrrd are not in this image, so the same information is handled here: attached-header NRRD,
raw or gzip encoding, 3-D scalar volumes, `space directions` + `space origin` (or `spacings`).
Host-side file format code; nothing here touches the GPU.
"""
from __future__ import annotations

import gzip
import re

import numpy as np

_TYPES = {
    "signed char": "i1", "int8": "i1", "int8_t": "i1", "uchar": "u1", "unsigned char": "u1", "uint8": "u1",
    "uint8_t": "u1", "short": "i2", "short int": "i2", "signed short": "i2", "int16": "i2", "int16_t": "i2",
    "ushort": "u2", "unsigned short": "u2", "uint16": "u2", "uint16_t": "u2", "int": "i4", "signed int": "i4",
    "int32": "i4", "int32_t": "i4", "uint": "u4", "unsigned int": "u4", "uint32": "u4", "uint32_t": "u4",
    "longlong": "i8", "long long": "i8", "int64": "i8", "int64_t": "i8", "ulonglong": "u8",
    "unsigned long long": "u8", "uint64": "u8", "uint64_t": "u8", "float": "f4", "double": "f8",
}
_NAMES = {"i1": "int8", "u1": "uint8", "i2": "int16", "u2": "uint16", "i4": "int32", "u4": "uint32",
          "i8": "int64", "u8": "uint64", "f4": "float", "f8": "double"}


def _vec(text):
    return [float(v) for v in text.strip().strip("()").split(",")]


class NrrdImage:
    """array is (z,y,x) like sitk.GetArrayFromImage; size/spacing/origin are (x,y,z) like SimpleITK."""

    def __init__(self, array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=None, space=None):
        self.array = np.asarray(array)
        self.spacing = tuple(float(s) for s in spacing)
        self.origin = tuple(float(o) for o in origin)
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
        self.space = space or "left-posterior-superior"

    def GetSize(self):
        return tuple(int(s) for s in self.array.shape[::-1])

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return tuple(self.direction.reshape(-1))


def read_nrrd(path) -> NrrdImage:
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.find(b"\n\n")
    sep = 2
    if end < 0 or (0 <= raw.find(b"\r\n\r\n") < end):
        end, sep = raw.find(b"\r\n\r\n"), 4
    if end < 0 or not raw.startswith(b"NRRD"):
        raise ValueError(f"{path}: not an attached-header NRRD file")
    fields = {}
    for line in raw[:end].decode("ascii", errors="replace").splitlines()[1:]:
        if not line or line.startswith("#"):
            continue
        m = re.match(r"^([^:]+):=?\s*(.*)$", line)
        if m:
            fields[m.group(1).strip().lower()] = m.group(2).strip()
    if int(fields.get("dimension", "0")) != 3:
        raise ValueError(f"{path}: only 3-D scalar volumes are supported")
    if "data file" in fields or "datafile" in fields:
        raise ValueError(f"{path}: detached data files are not supported")
    code = _TYPES.get(fields["type"].lower())
    if code is None:
        raise ValueError(f"{path}: unsupported type {fields['type']!r}")
    endian = "<" if fields.get("endian", "little").lower() == "little" else ">"
    sizes = [int(v) for v in fields["sizes"].split()]            # x y z (fastest first)
    enc = fields.get("encoding", "raw").lower()
    body = raw[end + sep:]
    if enc in ("gzip", "gz"):
        body = gzip.decompress(body)
    elif enc != "raw":
        raise ValueError(f"{path}: unsupported encoding {enc!r}")
    n = sizes[0] * sizes[1] * sizes[2]
    arr = np.frombuffer(body, dtype=np.dtype(endian + code), count=n).reshape(sizes[2], sizes[1], sizes[0])
    direction, spacing = np.eye(3), [1.0, 1.0, 1.0]
    if "space directions" in fields:
        vecs = [_vec(v) for v in re.findall(r"\([^)]*\)", fields["space directions"])]
        if len(vecs) != 3:
            raise ValueError(f"{path}: need three space directions")
        for k, v in enumerate(vecs):
            nv = float(np.linalg.norm(v))
            spacing[k] = nv
            direction[:, k] = np.asarray(v) / nv if nv > 0 else np.eye(3)[:, k]
    elif "spacings" in fields:
        spacing = [float(v) for v in fields["spacings"].split()]
    origin = _vec(fields["space origin"]) if "space origin" in fields else [0.0, 0.0, 0.0]
    return NrrdImage(arr, spacing, origin, direction, fields.get("space"))


def write_nrrd(path, array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=None,
               space="left-posterior-superior", encoding="gzip"):
    """Write a (z,y,x) array; spacing / origin / direction in (x,y,z) order, as SimpleITK holds them."""
    arr = np.ascontiguousarray(array)
    if arr.ndim != 3:
        raise ValueError("write_nrrd expects a 3-D (z,y,x) array")
    code = arr.dtype.str[1:]
    if code not in _NAMES:
        raise ValueError(f"unsupported dtype {arr.dtype}")
    d = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64).reshape(3, 3)
    dirs = " ".join("(" + ",".join(repr(float(d[r, k] * spacing[k])) for r in range(3)) + ")" for k in range(3))
    header = ["NRRD0004", "# written by physics_arx_b200.io_nrrd", f"type: {_NAMES[code]}", "dimension: 3",
              f"space: {space}", f"sizes: {arr.shape[2]} {arr.shape[1]} {arr.shape[0]}",
              f"space directions: {dirs}", "kinds: domain domain domain"]
    if arr.dtype.itemsize > 1:
        header.append("endian: little")
    header += [f"encoding: {encoding}", "space origin: (" + ",".join(repr(float(o)) for o in origin) + ")"]
    body = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes()
    if encoding == "gzip":
        body = gzip.compress(body, compresslevel=1)
    elif encoding != "raw":
        raise ValueError(f"unsupported encoding {encoding!r}")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n\n").encode("ascii"))
        f.write(body)
