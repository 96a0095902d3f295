"""GPU-box scan: one level between 48 K and 2^18 pixels (the mid levels of a pyramid), every form x sweeps per pass, against the
pass planner's own plan.  python tools/tune_mid_levels.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402


TS = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (4, 6, 8, 9, 11, 13, 16)
TOP = int(sys.argv[2]) if len(sys.argv) > 2 else 8


def rep(t, total):
    p = [t] * (total // t)
    if total % t:
        p = [total % t] + p
    return p


for rows, cols, iters in ((270, 480, 250), (360, 640, 125), (426, 640, 125), (455, 455, 125), (312, 336, 250), (300, 700, 125), (227, 227, 250)):
    rng = np.random.default_rng(1)
    gray = synth.synth_image(rows, cols, 3)[..., 0].copy()
    depth = (rng.integers(0, 5, (rows, cols)) * 60 + rng.uniform(0, 14, (rows, cols))).astype(np.float32)
    scribble = np.where(rng.random((rows, cols)) < 0.1, 255, 0).astype(np.uint8)
    d0, s, g = to_dev(depth), to_dev(scribble), to_dev(gray)
    res = []
    for name, tile, tma in (("planner", 0, 2), ("flat", 34, 0), ("single", 64, 1), ("cluster", 64, 3)):
        for t in ((None,) if name == "planner" else TS):
            if name == "flat" and t > 15:
                continue
            ctx = rtdd.DepthDiffusion(rows, cols, 1)
            ctx.set_tuning("blocked_tile", tile)
            ctx.set_tuning("blocked_tma", tma)
            ctx.set_sweep_variant(2, 0)
            if os.environ.get("RTDD_PDL") is not None:
                ctx.set_tuning("pdl", int(os.environ["RTDD_PDL"]))
            if t is not None:
                ctx.set_pass_plan(0, rep(t, iters))
            d = d0.clone()
            ms = []
            for _ in range(6):
                ctx.matrix_free_solver(d, s, g, iters, 0)
                ctx.sync()
                ms.append(ctx.level_sweep_ms(0)[0])
            res.append((float(np.median(ms)), name, t))
            ctx.set_tuning("blocked_tile", 0)
            ctx.set_tuning("blocked_tma", 2)
            ctx.set_tuning("pdl", 1)
            ctx.close()
    res.sort()
    print("%dx%d x %d sweeps, planner %s: " % (cols, rows, iters, rtdd.DepthDiffusion.plan_passes(rows, cols, iters)) +
          "  ".join("%.4f(%s %s)" % r for r in res[:TOP]), flush=True)
