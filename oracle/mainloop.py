"""Headless restatement of main.cpp's solve loop over ANY implementation of the
reference's ten-function API (ref: src/main.cpp:116-173, 232-295).

`api` is a dict {name: ctypes callable} -- librtdd.so's reference-named shims
(realtimedepthdiffusion_b200._native.shims) or the reference's own kernels
(oracle.binding.ref_api()).  The OpenCV steps outside the GPU* boundary (gray
pyrDown, depth pyrUp, final convertTo) are done on the host by the oracle's
restatements for BOTH sides, the way main.cpp itself falls back to the CPU for
them, so they cancel out of an A/B comparison.
"""
import ctypes as C

import numpy as np
import torch

from oracle import binding as ob

pitched_empty = ob.pkg_file("planes").pitched_empty


def ptr(t):
    return C.c_void_p(t.data_ptr())


def pitch(t):
    return t.stride(0) * t.element_size()


def to_dev(a, channels=1, device="cuda"):
    """numpy [rows, cols(,3)] -> pitched device plane [rows, cols*channels]."""
    a = np.ascontiguousarray(a)
    rows, cols = a.shape[:2]
    dt = torch.from_numpy(a.reshape(rows, -1))
    out = pitched_empty(rows, cols, dt.dtype, device, channels=channels)
    out.copy_(dt)
    return out


def to_host(t, channels=1):
    a = t.cpu().numpy()
    if channels > 1:
        a = a.reshape(a.shape[0], -1, channels)
    return a


class MainLoop:
    def __init__(self, api, bgr, levels=None, device="cuda"):
        self.api = api
        self.device = device
        bgr = np.ascontiguousarray(bgr, np.uint8)
        self.rows, self.cols = bgr.shape[:2]
        self.levels = ob.pyramid_levels(self.rows, self.cols) if levels is None else levels
        self.sizes = ob.level_sizes(self.rows, self.cols, self.levels)
        gray = [ob.bgr2gray(bgr)]
        for l in range(1, self.levels):
            gray.append(ob.pyrdown_gray(gray[l - 1]))
        self.gray_host = gray
        self.orig = to_dev(bgr, 3, device)
        self.gray = [to_dev(g, 1, device) for g in gray]
        self.scribble = [pitched_empty(r, c, torch.uint8, device, fill=0) for (r, c) in self.sizes]
        self.edited = [pitched_empty(r, c, torch.uint8, device, channels=3, fill=0) for (r, c) in self.sizes]
        self.depth = [pitched_empty(r, c, torch.float32, device, fill=255.0) for (r, c) in self.sizes]
        torch.cuda.synchronize()
        api["GPUAllocateDeviceMemory"](self.rows, self.cols, self.levels)
        api["GPULoadWeights"](0.4)
        self.per_level = {}

    def close(self):
        torch.cuda.synchronize()
        self.api["GPUFreeDeviceMemory"](self.levels)

    def solve_level(self, l, iters):
        r, c = self.sizes[l]
        torch.cuda.synchronize()
        self.api["GPUMatrixFreeSolver"](ptr(self.depth[l]), pitch(self.depth[l]), ptr(self.scribble[l]), pitch(self.scribble[l]),
                                        ptr(self.gray[l]), pitch(self.gray[l]), r, c, 0.4, iters, 1e-5, l)

    def convert(self, l):
        r, c = self.sizes[l]
        self.api["GPUConvertToFloat"](ptr(self.edited[l]), pitch(self.edited[l]), ptr(self.depth[l]), pitch(self.depth[l]),
                                      ptr(self.scribble[l]), pitch(self.scribble[l]), r, c)

    def frame(self, scribble0, edited0, max_iterations=1000, keep_levels=False):
        """One 'd' key press: returns the u8 depth map (host)."""
        L = self.levels
        self.scribble[0].copy_(torch.from_numpy(np.ascontiguousarray(scribble0)))
        self.edited[0].copy_(torch.from_numpy(np.ascontiguousarray(edited0).reshape(self.rows, -1)))
        torch.cuda.synchronize()
        for l in range(1, L):
            pr, pc = self.sizes[l - 1]
            r, c = self.sizes[l]
            self.api["GPUPyrDownAnnotation"](ptr(self.scribble[l - 1]), pitch(self.scribble[l - 1]), ptr(self.edited[l - 1]),
                                             pitch(self.edited[l - 1]), pr, pc, ptr(self.scribble[l]), pitch(self.scribble[l]),
                                             ptr(self.edited[l]), pitch(self.edited[l]), r, c)
        self.convert(L - 1)
        for l in range(L - 1, -1, -1):
            iters = int(np.float32(max_iterations) / np.float32(2.0 ** ((L - 1) - l)))
            if keep_levels:
                torch.cuda.synchronize()
                self.per_level[l] = {"in": to_host(self.depth[l])}
            self.solve_level(l, iters)
            if keep_levels:
                torch.cuda.synchronize()
                self.per_level[l]["out"] = to_host(self.depth[l])
            if l > 0:
                torch.cuda.synchronize()
                r, c = self.sizes[l - 1]
                up = ob.pyrup_f32(to_host(self.depth[l]), r, c)          # main.cpp:275-279 (CPU cv::pyrUp path)
                self.depth[l - 1].copy_(torch.from_numpy(up))
                torch.cuda.synchronize()
                self.convert(l - 1)
        torch.cuda.synchronize()
        self.depth_float = to_host(self.depth[0])
        return ob.quantise_u8(self.depth_float)
