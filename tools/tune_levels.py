"""GPU-box scan: per-level sweep time of the blocked forms on a whole frame, for (form, cluster size, sweeps per pass)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

sizes = {"4k": (2160, 3840, 1003), "1080p": (1080, 1920, 1002), "8k": (4320, 7680, 1004)}
for name in (sys.argv[1:] or ["4k"]):
    rows, cols, seed = sizes[name]
    bgr, scribble, edited = synth.synth_case(rows, cols, seed)
    best = {}
    for mode, cl in ((1, 1), (3, 1), (3, 2)):
        for T in (6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16):
            ctx = rtdd.DepthDiffusion(rows, cols)
            ctx.set_tuning("blocked_tma", mode)
            ctx.set_tuning("blocked_cluster", cl)
            ctx.set_tuning("blocked_tile", 64)
            ctx.set_sweep_variant(2, T)
            ctx.frame_set_image(bgr)
            ctx.frame_solve_host(scribble, edited, 1000, np.zeros((rows, cols), np.uint8))
            for _ in range(6):
                ctx.frame_solve(1000)
            ctx.sync()
            for l in range(ctx.levels):
                ms, it, k = ctx.level_sweep_ms(l)
                best.setdefault(l, []).append((ms, mode, cl, T, k))
            ctx.set_tuning("blocked_tma", 2)
            ctx.set_tuning("blocked_cluster", 2)
            ctx.set_tuning("blocked_tile", 0)
            ctx.close()
    for l, lst in sorted(best.items()):
        lst.sort()
        print(name, "level", l, " ".join("%.4f(m%d c%d T%d k%d)" % x for x in lst[:6]), flush=True)
