"""GPU-box scan: level-0 pass plans of frames between 2^18 and 2^21 pixels (dataset-sized images), both tiling forms -- a check
of the pass planner's cost model where nobody fitted it.  python tools/tune_passes_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402


def rep(t, total):
    p = [t] * (total // t)
    if total % t:
        p = [total % t] + p
    return p


for rows, cols in ((720, 1280), (853, 1280), (910, 910), (624, 672)):
    bgr, scribble, edited = synth.synth_case(rows, cols, 5 + rows)
    out = np.zeros((rows, cols), np.uint8)
    probe = rtdd.DepthDiffusion(rows, cols)
    iters0 = 1000 >> (probe.levels - 1)
    planned = rtdd.DepthDiffusion.plan_passes(rows, cols, iters0)
    probe.close()
    res = []
    for tma, nm in ((2, "planner"), (1, "single"), (3, "cluster")):
        for t in ((None,) if tma == 2 else (4, 6, 8, 11, 13, 16)):
            ctx = rtdd.DepthDiffusion(rows, cols)
            ctx.set_tuning("blocked_tma", tma)
            if t is not None:
                ctx.set_pass_plan(0, rep(t, iters0))
            ctx.frame_set_image(bgr)
            ctx.frame_solve_host(scribble, edited, 1000, out)
            ms = []
            for _ in range(8):
                ctx.frame_solve(1000)
                ctx.sync()
                ms.append(ctx.level_sweep_ms(0)[0])
            res.append((float(np.median(ms)), nm, t))
            ctx.set_tuning("blocked_tma", 2)
            ctx.close()
    res.sort()
    print("%dx%d level 0 x %d sweeps, planner %s: " % (cols, rows, iters0, planned) + "  ".join("%.4f(%s %s)" % r for r in res), flush=True)
