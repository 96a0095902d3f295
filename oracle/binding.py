"""ctypes binding of oracle/liboracle.so (CPU restatement) and oracle/_ref/libref.so
(the reference's own CUDA kernels, compiled unmodified).

TEST INFRASTRUCTURE.  Import only from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke().  The product never imports this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "realtimedepthdiffusion_b200")


def pkg_file(name):
    """A native-free module of the product package (refnames, planes, synth) loaded BY PATH: importing the package itself
    dlopens librtdd.so, which the reference arm of bench.py must not have in its process."""
    import importlib.util
    import sys
    key = "_rtdd_byPath_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(_PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod

LIBORACLE = os.path.join(_HERE, "liboracle.so")
LIBREF = os.path.join(_HERE, "_ref", "libref.so")

vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBORACLE):
            raise ImportError("oracle/liboracle.so missing: run python realtimedepthdiffusion_b200/build.py")
        L = C.CDLL(LIBORACLE)
        sigs = {
            "oracle_num_threads": (i32, []),
            "oracle_set_num_threads": (None, [i32]),
            "oracle_load_weights": (None, [f32, vp]),
            "oracle_omega_schedule": (None, [i32, vp]),
            "oracle_index_to_weight": (None, [vp, sz, vp, sz, i32, i32, i32, i32, vp]),
            "oracle_pack_int2": (None, [vp, i32, vp]),
            "oracle_sweep": (None, [vp, vp, vp, vp, vp, sz, vp, i32, i32, f32, f32]),
            "oracle_solve_level": (i32, [vp, sz, vp, sz, vp, sz, i32, i32, i32, i32, i32, vp]),
            "oracle_convert_to_float": (None, [vp, sz, vp, sz, vp, sz, i32, i32]),
            "oracle_pyrdown_annotation": (None, [vp, sz, vp, sz, i32, i32, vp, sz, vp, sz, i32, i32]),
            "oracle_paint": (None, [i32, i32, i32, i32, vp, sz, vp, sz, i32, i32]),
            "oracle_desaturate": (None, [vp, sz, vp, sz, vp, sz, vp, sz, i32, i32]),
            "oracle_haze": (None, [vp, sz, vp, sz, vp, sz, i32, i32]),
            "oracle_defocus_kernel_size": (i32, [i32, i32]),
            "oracle_defocus": (None, [vp, sz, vp, sz, vp, sz, i32, i32]),
            "oracle_bgr2gray": (None, [vp, sz, vp, sz, i32, i32]),
            "oracle_pyrdown_gray": (None, [vp, sz, i32, i32, vp, sz]),
            "oracle_pyrup_f32": (i32, [vp, sz, i32, i32, vp, sz, i32, i32]),
            "oracle_quantise_u8": (None, [vp, sz, vp, sz, i32, i32]),
        }
        for n, (r, a) in sigs.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(vp)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(n)


def load_weights(beta=0.4):
    lut = np.zeros(257, np.float32)
    lib().oracle_load_weights(beta, _p(lut))
    return lut


def omega_schedule(n):
    om = np.zeros(max(n, 1), np.float32)
    lib().oracle_omega_schedule(n, _p(om))
    return om[:n]


def index_to_weight(gray, depth, level, max_level):
    """-> int32 [rows, cols, 4] = (left, right, up, down) LUT indices, 256 = outside.
    gray may be larger than depth (ceil vs floor sizes); only the top-left window is read."""
    depth = _c(depth, np.float32)
    gray = _c(gray, np.uint8)
    rows, cols = depth.shape
    idx = np.zeros((rows, cols, 4), np.int32)
    lib().oracle_index_to_weight(_p(gray), gray.strides[0], _p(depth), depth.strides[0], rows, cols, level, max_level, _p(idx))
    return idx


def links_from_index(idx):
    """(right, down) u8 link planes in librtdd's layout from the oracle's 4-index plane."""
    right = np.where(idx[..., 1] == 256, 0, idx[..., 1]).astype(np.uint8)
    down = np.where(idx[..., 3] == 256, 0, idx[..., 3]).astype(np.uint8)
    return right, down


def solve_level(depth, scribble, gray, max_iterations, level, max_level, lut=None):
    """GPUMatrixFreeSolver restated; returns the new depth plane (input untouched)."""
    d = np.array(depth, dtype=np.float32, order="C", copy=True)
    s = _c(scribble, np.uint8)
    g = _c(gray, np.uint8)
    if lut is None:
        lut = load_weights(0.4)
    rows, cols = d.shape
    rc = lib().oracle_solve_level(_p(d), d.strides[0], _p(s), s.strides[0], _p(g), g.strides[0], rows, cols, int(max_iterations),
                                  int(level), int(max_level), _p(lut))
    if rc != 0:
        raise MemoryError("oracle_solve_level")
    return d


def convert_to_float(src3, dst, mask):
    d = np.array(dst, dtype=np.float32, order="C", copy=True)
    s = _c(src3, np.uint8).reshape(d.shape[0], -1)
    m = _c(mask, np.uint8)
    lib().oracle_convert_to_float(_p(s), s.strides[0], _p(d), d.strides[0], _p(m), m.strides[0], d.shape[0], d.shape[1])
    return d


def pyrdown_annotation(prev_scribble, prev_edited, curr_scribble, curr_edited):
    ps = _c(prev_scribble, np.uint8)
    pe = _c(prev_edited, np.uint8).reshape(ps.shape[0], -1)
    cs = np.array(curr_scribble, dtype=np.uint8, order="C", copy=True)
    ce = np.array(curr_edited, dtype=np.uint8, order="C", copy=True).reshape(cs.shape[0], -1)
    lib().oracle_pyrdown_annotation(_p(ps), ps.strides[0], _p(pe), pe.strides[0], ps.shape[0], ps.shape[1],
                                    _p(cs), cs.strides[0], _p(ce), ce.strides[0], cs.shape[0], cs.shape[1])
    return cs, ce.reshape(cs.shape[0], cs.shape[1], 3)


def paint(x, y, color, radius, edited, scribble):
    s = np.array(scribble, dtype=np.uint8, order="C", copy=True)
    e = np.array(edited, dtype=np.uint8, order="C", copy=True).reshape(s.shape[0], -1)
    lib().oracle_paint(x, y, color, radius, _p(e), e.strides[0], _p(s), s.strides[0], s.shape[0], s.shape[1])
    return e.reshape(s.shape[0], s.shape[1], 3), s


def desaturate(orig, gray, depth):
    d = _c(depth, np.float32)
    o = _c(orig, np.uint8).reshape(d.shape[0], -1)
    g = _c(gray, np.uint8)
    out = np.zeros_like(o)
    lib().oracle_desaturate(_p(o), o.strides[0], _p(g), g.strides[0], _p(d), d.strides[0], _p(out), out.strides[0], d.shape[0], d.shape[1])
    return out.reshape(d.shape[0], d.shape[1], 3)


def haze(orig, depth):
    d = _c(depth, np.float32)
    o = _c(orig, np.uint8).reshape(d.shape[0], -1)
    out = np.zeros_like(o)
    lib().oracle_haze(_p(o), o.strides[0], _p(d), d.strides[0], _p(out), out.strides[0], d.shape[0], d.shape[1])
    return out.reshape(d.shape[0], d.shape[1], 3)


def defocus_kernel_size(rows, cols):
    return lib().oracle_defocus_kernel_size(rows, cols)


def defocus(orig, depth):
    d = _c(depth, np.float32)
    o = _c(orig, np.uint8).reshape(d.shape[0], -1)
    out = np.zeros_like(o)
    lib().oracle_defocus(_p(o), o.strides[0], _p(d), d.strides[0], _p(out), out.strides[0], d.shape[0], d.shape[1])
    return out.reshape(d.shape[0], d.shape[1], 3)


def bgr2gray(bgr):
    b = _c(bgr, np.uint8)
    rows, cols = b.shape[:2]
    b2 = b.reshape(rows, -1)
    g = np.zeros((rows, cols), np.uint8)
    lib().oracle_bgr2gray(_p(b2), b2.strides[0], _p(g), g.strides[0], rows, cols)
    return g


def pyrdown_gray(src):
    s = _c(src, np.uint8)
    d = np.zeros(((s.shape[0] + 1) // 2, (s.shape[1] + 1) // 2), np.uint8)
    lib().oracle_pyrdown_gray(_p(s), s.strides[0], s.shape[0], s.shape[1], _p(d), d.strides[0])
    return d


def pyrup_f32(src, drows, dcols):
    s = _c(src, np.float32)
    d = np.zeros((drows, dcols), np.float32)
    rc = lib().oracle_pyrup_f32(_p(s), s.strides[0], s.shape[0], s.shape[1], _p(d), d.strides[0], drows, dcols)
    if rc != 0:
        raise ValueError("pyrup size")
    return d


def quantise_u8(src):
    s = _c(src, np.float32)
    d = np.zeros(s.shape, np.uint8)
    lib().oracle_quantise_u8(_p(s), s.strides[0], _p(d), d.strides[0], s.shape[0], s.shape[1])
    return d


# ---- headless restatement of the frame loop (ref: src/main.cpp:95-113, 232-295) --------

def pyramid_levels(rows, cols):
    import math
    return int(math.log2(max(min(cols, rows) // 45, 1))) + 1


def level_sizes(rows, cols, levels):
    return [(int(rows / 2.0 ** l), int(cols / 2.0 ** l)) for l in range(levels)]


class FrameState:
    """What main.cpp keeps between frames: per-level gray (ceil sizes), scribble, edited, depth."""

    def __init__(self, bgr, levels=None):
        bgr = _c(bgr, np.uint8)
        self.rows, self.cols = bgr.shape[:2]
        self.levels = pyramid_levels(self.rows, self.cols) if levels is None else levels
        self.sizes = level_sizes(self.rows, self.cols, self.levels)
        self.bgr = bgr
        self.gray = [bgr2gray(bgr)]
        for l in range(1, self.levels):
            self.gray.append(pyrdown_gray(self.gray[l - 1]))
        self.scribble = [np.zeros(s, np.uint8) for s in self.sizes]
        self.edited = [np.zeros(s + (3,), np.uint8) for s in self.sizes]
        self.depth = [np.full(s, 255.0, np.float32) for s in self.sizes]
        self.lut = load_weights(0.4)
        self.per_level = {}

    def solve(self, scribble0, edited0, max_iterations=1000, keep_levels=False):
        L = self.levels
        self.scribble[0] = _c(scribble0, np.uint8).copy()
        self.edited[0] = _c(edited0, np.uint8).copy()
        for l in range(1, L):
            self.scribble[l], self.edited[l] = pyrdown_annotation(self.scribble[l - 1], self.edited[l - 1], self.scribble[l], self.edited[l])
        self.depth[L - 1] = convert_to_float(self.edited[L - 1], self.depth[L - 1], self.scribble[L - 1])
        for l in range(L - 1, -1, -1):
            iters = int(np.float32(max_iterations) / np.float32(2.0 ** ((L - 1) - l)))
            if keep_levels:
                self.per_level[l] = {"in": self.depth[l].copy()}
            self.depth[l] = solve_level(self.depth[l], self.scribble[l], self.gray[l], iters, l, L - 1, self.lut)
            if keep_levels:
                self.per_level[l]["out"] = self.depth[l].copy()
            if l > 0:
                r, c = self.sizes[l - 1]
                up = pyrup_f32(self.depth[l], r, c)
                self.depth[l - 1] = convert_to_float(self.edited[l - 1], up, self.scribble[l - 1])
        return quantise_u8(self.depth[0])


# ---- the reference's own CUDA kernels (needs a GPU) -----------------------------------------

_ref = None


def ref_api():
    """{name: callable} for the ten GPU* functions of oracle/_ref/libref.so (device pointers!)."""
    global _ref
    if _ref is None:
        if not os.path.exists(LIBREF):
            raise ImportError("oracle/_ref/libref.so missing (built from /root/reference by realtimedepthdiffusion_b200/build.py)")
        _ref = pkg_file("refnames").bind_reference_api(C.CDLL(LIBREF, mode=os.RTLD_LOCAL | os.RTLD_NOW))
    return _ref
