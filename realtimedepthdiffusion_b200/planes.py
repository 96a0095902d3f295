"""Pitched device planes (what cv::cuda::GpuMat / cudaMallocPitch give main.cpp).  torch only, no native dependency:
also loaded by path by the oracle-side harness."""
import torch


def pitched_empty(rows, cols, dtype, device, channels=1, align=512, fill=None):
    """A rows x (cols*channels) plane whose row pitch is a multiple of `align` bytes,
    like cv::cuda::GpuMat / cudaMallocPitch.  Returns a strided view; .stride(0)*itemsize is the pitch."""
    item = torch.empty((), dtype=dtype).element_size()
    row_bytes = cols * channels * item
    pitch = (row_bytes + align - 1) // align * align
    if rows == 0 or cols == 0:
        return torch.empty((rows, cols * channels), dtype=dtype, device=device)
    base = torch.empty((rows, pitch // item), dtype=dtype, device=device)
    if fill is not None:
        base.fill_(fill)
    return base[:, : cols * channels]


def to_dev(a, channels=1, device="cuda"):
    """numpy [rows, cols(, channels)] -> pitched device plane [rows, cols * channels] (tools, tests)."""
    import numpy as np
    a = np.ascontiguousarray(a)
    rows, cols = a.shape[:2]
    dt = torch.from_numpy(a.reshape(rows, -1))
    out = pitched_empty(rows, cols, dt.dtype, device, channels=channels)
    out.copy_(dt)
    return out
