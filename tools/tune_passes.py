"""GPU-box scan: pass plans of the finest levels of a whole frame (rtdd_set_pass_plan) -- level time on the device and the
end-to-end call (annotation plane up, 8-bit map down) with the map stored by the last pass (zero copy) or copied afterwards."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

sizes = {"4k": (2160, 3840, 1003), "1080p": (1080, 1920, 1002), "8k": (4320, 7680, 1004)}
plans = {
    "4k": {0: [None, [7, 7, 7, 7, 3], [7, 7, 7, 10], [10, 7, 7, 7], [10, 10, 11], [11, 10, 10], [8, 8, 15], [6, 6, 6, 6, 7], [15, 16], [5, 5, 5, 5, 11],
               [7, 8, 16], [7, 7, 7, 6, 4], [9, 9, 13], [7, 12, 12]],
           1: [None, [16, 16, 16, 14], [15, 15, 16, 16], [12, 12, 12, 13, 13], [10, 10, 10, 10, 11, 11], [14, 16, 16, 16]]},
    "1080p": {0: [None, [16, 16, 16, 14], [15, 15, 16, 16], [12, 12, 12, 13, 13], [14, 16, 16, 16], [10, 10, 10, 10, 11, 11]]},
    "8k": {0: [None, [8, 7], [7, 8], [4, 11], [15], [5, 10], [5, 5, 5]]},
}


def e2e_ms(ctx, annot, out, n=15):
    for _ in range(3):
        ctx.frame_solve_host_annotation(annot, 1000, out)
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.frame_solve_host_annotation(annot, 1000, out)
    return (time.perf_counter() - t0) / n * 1e3


for name in (sys.argv[1:] or ["4k"]):
    rows, cols, seed = sizes[name]
    bgr, scribble, edited = synth.synth_case(rows, cols, seed)
    annot = torch.from_numpy(synth.annotation_plane(scribble, edited)).pin_memory()
    out = torch.empty((rows, cols), dtype=torch.uint8).pin_memory()
    for level, lst in plans[name].items():
        for plan in lst:
            ctx = rtdd.DepthDiffusion(rows, cols)
            ctx.frame_set_image(bgr)
            if plan is not None:
                ctx.set_pass_plan(level, plan)
            res = []
            for zc in (1, 0):
                ctx.set_tuning("zero_copy_out", zc)
                res.append(e2e_ms(ctx, annot, out))
            ctx.set_tuning("zero_copy_out", 1)
            for _ in range(6):
                ctx.frame_solve(1000)
            ctx.sync()
            ms = [ctx.level_sweep_ms(l) for l in range(min(2, ctx.levels))]
            print(name, "level", level, "plan", plan, " ".join("L%d %.4f ms (%d launches)" % (l, m[0], m[2]) for l, m in enumerate(ms)),
                  "e2e zero-copy %.3f ms, staged %.3f ms" % (res[0], res[1]), flush=True)
            ctx.close()
