"""``tigre.utilities.gpu`` surface (reference call sites pCT_CBCT_Reconstruction.py:24-32)."""
from __future__ import annotations

import torch


class GpuIds:
    """Devices sharing one name; what ``gpuids=`` carries into Ax / Atb / ossart."""

    def __init__(self, name=None, devices=None):
        self.name = name
        self.devices = list(devices or [])

    def __len__(self):
        return len(self.devices)

    def __repr__(self):
        return f"GpuIds(name={self.name!r}, devices={self.devices})"


def getGpuCount():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def getGpuNames():
    return [torch.cuda.get_device_name(i) for i in range(getGpuCount())]


def getGpuIds(name=None):
    names = getGpuNames()
    if name is None:
        return GpuIds(None, range(len(names)))
    return GpuIds(name, [i for i, n in enumerate(names) if n == name])
