"""Oracle-backed strip engine: lets the CPU tests (gloo, world_size 2) exercise realtimedepthdiffusion_b200.strips --
strip planning, ghost-row staleness, halo exchange, windowed prolongation -- without a GPU.  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np
import torch

from oracle import binding as ob


class CpuStripEngine:
    def __init__(self, bgr, scribble, edited):
        st = ob.FrameState(bgr)
        self.st = st
        self.levels, self.sizes = st.levels, st.sizes
        st.scribble[0] = np.ascontiguousarray(scribble).copy()
        st.edited[0] = np.ascontiguousarray(edited).copy()
        self.depth_u8 = np.zeros(self.sizes[0], np.uint8)
        self.win = {}

    def annotation_pyramid(self):
        st = self.st
        for l in range(1, self.levels):
            st.scribble[l], st.edited[l] = ob.pyrdown_annotation(st.scribble[l - 1], st.edited[l - 1], st.scribble[l], st.edited[l])

    def convert_rows(self, l, r0, r1):
        st = self.st
        st.depth[l][r0:r1] = ob.convert_to_float(st.edited[l][r0:r1], st.depth[l][r0:r1], st.scribble[l][r0:r1])

    def solve_full(self, l, iters):
        st = self.st
        st.depth[l] = ob.solve_level(st.depth[l], st.scribble[l], st.gray[l], iters, l, self.levels - 1, st.lut)

    def pyrup_rows(self, l, r0, r1):
        st = self.st
        src = st.depth[l]
        r, c = self.sizes[l - 1]
        # prolong only from the coarse rows this rank really holds: rows outside are poisoned to prove they are unused
        lo, hi = max(0, (r0 >> 1) - 1), min(src.shape[0], ((r1 - 1) >> 1) + 2)
        poisoned = np.full_like(src, np.nan)
        poisoned[lo:hi] = src[lo:hi]
        up = ob.pyrup_f32(poisoned, r, c)
        assert np.isfinite(up[r0:r1]).all()
        st.depth[l - 1][r0:r1] = up[r0:r1]

    def strip_init(self, l, w0, w1):
        st = self.st
        d = np.ascontiguousarray(st.depth[l][w0:w1])
        g = np.ascontiguousarray(st.gray[l][w0:w1 + 1])        # the row below the window is never used for a kept link
        idx = ob.index_to_weight(g[: w1 - w0], d, l, self.levels - 1)
        self.win[l] = {"w0": w0, "w1": w1, "idx": idx, "scr": np.ascontiguousarray(st.scribble[l][w0:w1]),
                       "xk": torch.from_numpy(d.copy()), "xkm1": torch.zeros(d.shape, dtype=torch.float32)}

    def strip_pass(self, l, k0, n, halo):
        w = self.win[l]
        om = ob.omega_schedule(k0 + n)
        rows, cols = w["xk"].shape
        lib = ob.lib()
        for i in range(n):
            a = w["xk"].numpy()
            prev = w["xkm1"].numpy()
            out = a.copy()                                      # Dirichlet pixels keep their value
            lib.oracle_sweep(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), prev.ctypes.data_as(C.c_void_p),
                             w["idx"].ctypes.data_as(C.c_void_p), w["scr"].ctypes.data_as(C.c_void_p), w["scr"].strides[0],
                             self.st.lut.ctypes.data_as(C.c_void_p), rows, cols, float(om[k0 + i]), np.float32(0.99))
            # oracle_sweep left prev = x_k on free pixels; make it x_k everywhere (Dirichlet: value never used)
            w["xkm1"] = torch.from_numpy(a.copy())
            w["xk"] = torch.from_numpy(out)

    def strip_planes(self, l):
        w = self.win[l]
        return [w["xk"], w["xkm1"]]

    def strip_finish(self, l, r0, r1):
        w = self.win[l]
        self.st.depth[l][r0:r1] = w["xk"].numpy()[r0 - w["w0"]: r1 - w["w0"]]

    def quantise_rows(self, r0, r1):
        self.depth_u8[r0:r1] = ob.quantise_u8(self.st.depth[0][r0:r1])
