"""Python binding of the native row-strip frame (rtdd_strip_frame_*, csrc/rtdd_stripframe.cu): one rank = one process = one GPU.

Everything that computes or schedules is inside librtdd.so; this file only (1) moves the CUDA IPC handles of the ranks' arenas
between the processes at set-up (torch.distributed all_gather_object -- plumbing, not on the data path) and (2) exposes the
entry points.  The single-process form (one host thread per GPU inside the library) is rtdd_mgpu_* and needs no Python at all
(tests/cpp/mgpu_host.cpp).
"""
import ctypes as C

import torch

from ._native import lib
from .api import DepthDiffusion, _pitch, _ptr


class StripFrameRank:
    def __init__(self, rows, cols, rank, world, halo=16, pass_sweeps=8, min_strip_pixels=1 << 22, levels=None, device=None):
        # window-sized scratch planes for the split levels (0.9 GB instead of 6.8 GB per rank at 16384^2 on 8 GPUs)
        self.ctx = DepthDiffusion(rows, cols, levels=levels, device=device, strip=(world, halo, min_strip_pixels) if world > 1 else None)
        self.rank, self.world = int(rank), int(world)
        self.rows, self.cols = int(rows), int(cols)
        self.ctx._ck(lib.rtdd_strip_frame_setup(self.ctx._h, self.rank, self.world, int(halo), int(pass_sweeps), int(min_strip_pixels)))

    # -- set-up ---------------------------------------------------------------------------------------------------------
    def connect(self, dist):
        """Map the neighbouring ranks' arenas (CUDA IPC) so that halo rows and flags travel as peer-memory stores."""
        if self.world <= 1:
            return
        h = C.create_string_buffer(64)
        self.ctx._ck(lib.rtdd_ipc_export(self.ctx._h, h))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(h.raw))
        ptrs = {}
        for nb in (self.rank - 1, self.rank + 1):
            if 0 <= nb < self.world:
                p = C.c_void_p()
                self.ctx._ck(lib.rtdd_ipc_import(self.ctx._h, C.create_string_buffer(handles[nb], 64), C.byref(p)))
                ptrs[nb] = p.value
        self.ctx._ck(lib.rtdd_strip_set_peers(self.ctx._h, C.c_void_p(ptrs.get(self.rank - 1) or 0), C.c_void_p(ptrs.get(self.rank + 1) or 0)))

    def set_image_device(self, bgr):
        """bgr: pitched device plane [rows, 3*cols] u8 (identical on every rank)."""
        self.ctx._ck(lib.rtdd_frame_set_image_device(self.ctx._h, _ptr(bgr), _pitch(bgr)))

    def plane(self, which, level=0):
        """A torch VIEW (no copy) of a context-owned frame plane: [rows, pitch elements]."""
        p, pitch, r, c = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self.ctx._ck(lib.rtdd_frame_plane(self.ctx._h, which, level, C.byref(p), C.byref(pitch), C.byref(r), C.byref(c)))
        f32 = which == self.ctx.PLANE_DEPTH
        ch = 3 if which in (self.ctx.PLANE_EDITED, self.ctx.PLANE_BGR) else 1
        item = 4 if f32 else 1

        class _H:
            pass
        h = _H()
        h.__cuda_array_interface__ = {"shape": (r.value, pitch.value // item), "typestr": "<f4" if f32 else "|u1", "data": (p.value, False), "version": 2}
        return torch.as_tensor(h, device=self.ctx.device)[:, : c.value * ch]

    def set_annotation_device(self, scribble, edited):
        """Level-0 annotation planes (device) -> the context's own planes (what main.cpp uploads, ref: src/main.cpp:236-237)."""
        self.plane(self.ctx.PLANE_SCRIBBLE).copy_(scribble)
        self.plane(self.ctx.PLANE_EDITED).copy_(edited)

    def reset_first_frame_guess(self):
        """depth planes = 255 (ref: src/main.cpp:136) with the level-0 Dirichlet values injected (main.cpp:281): a guess every
        rank can form for ALL rows on its own -- the starting point of the level-0-only measurement."""
        for l in range(self.ctx.levels):
            self.plane(self.ctx.PLANE_DEPTH, l).fill_(255.0)
        d, e, s = self.plane(self.ctx.PLANE_DEPTH), self.plane(self.ctx.PLANE_EDITED), self.plane(self.ctx.PLANE_SCRIBBLE)
        self.ctx.convert_to_float(e, d, s)

    # -- frames ---------------------------------------------------------------------------------------------------------
    def solve(self, max_iterations=1000):
        self.ctx._ck(lib.rtdd_strip_frame_solve(self.ctx._h, int(max_iterations)))

    def level0(self, sweeps):
        self.ctx._ck(lib.rtdd_strip_frame_level0(self.ctx._h, int(sweeps)))

    def rows_of(self, level=0):
        """(split?, own begin, own end, window begin, window end)"""
        v = [C.c_int() for _ in range(5)]
        self.ctx._ck(lib.rtdd_strip_frame_rows(self.ctx._h, int(level), *[C.byref(x) for x in v]))
        return bool(v[0].value), v[1].value, v[2].value, v[3].value, v[4].value

    def effects(self, desat, haze, defocus):
        self.ctx._ck(lib.rtdd_strip_frame_effects(self.ctx._h, _ptr(desat), _pitch(desat), _ptr(haze), _pitch(haze), _ptr(defocus), _pitch(defocus)))

    def close(self):
        self.ctx.close()
