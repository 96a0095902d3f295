"""GPU-box tuning on the real 4K / 1080p frame: per-level sweep time for every (tile, T) of the blocked kernel.
python tools/tune_frame.py [4k|1080p] > gpurun_out/tune_frame.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from bench import WORKLOADS                         # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "4k"
rows, cols, seed = WORKLOADS[name]
bgr, scribble, edited = synth.synth_case(rows, cols, seed)
out = np.zeros((rows, cols), np.uint8)
table = {}
for tile in (64, 32, 34):
    for T in (4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16):
        if tile in (32, 34) and 2 * T >= 32:
            continue
        ctx = rtdd.DepthDiffusion(rows, cols)
        ctx.set_tuning("blocked_tile", tile)
        ctx.set_sweep_variant(2, T)
        ctx.frame_set_image(bgr)
        ctx.frame_solve_host(scribble, edited, 1000, out)
        for _ in range(6):
            ctx.frame_solve(1000)
        ctx.sync()
        for l in range(ctx.levels):
            ms, it, k = ctx.level_sweep_ms(l)
            table.setdefault(l, []).append((ms, tile, T, k))
        ctx.set_tuning("blocked_tile", 0)
        ctx.close()
for l, lst in sorted(table.items()):
    lst.sort()
    print("level", l, " best:", ["%.4f ms tile %d T %d (%d launches)" % x for x in lst[:4]], " worst: %.4f" % lst[-1][0], flush=True)
