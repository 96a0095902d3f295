"""GPU-box measurement: a 4K frame whose image carries isolated salt-and-pepper pixels (weight sums below 2^-100: the pixels that
need the exact-division path) against the clean frame -- what the rare path costs when it is not rare.
    python tools/salt_pepper_frame.py [fraction of pixels, default 0.0005]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

rows, cols, seed = 2160, 3840, 1003
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0005
bgr, scribble, edited = synth.synth_case(rows, cols, seed)
rng = np.random.default_rng(7)
noisy = bgr.copy()
n = int(rows * cols * frac)
ys, xs = rng.integers(1, rows - 1, n), rng.integers(1, cols - 1, n)
noisy[ys, xs, :] = np.where(bgr[ys, xs, :1].astype(np.int32) < 128, 255, 0).astype(np.uint8)
out = np.zeros((rows, cols), np.uint8)
for name, img in (("clean", bgr), ("salt-and-pepper %.3f %% of the pixels" % (100 * frac), noisy)):
    ctx = rtdd.DepthDiffusion(rows, cols)
    ctx.frame_set_image(img)
    ctx.frame_solve_host(scribble, edited, 1000, out)
    ms = []
    for _ in range(8):
        ctx.frame_solve(1000)
        ctx.sync()
        ms.append([ctx.level_sweep_ms(l)[0] for l in range(ctx.levels)])
    med = np.median(np.array(ms), axis=0)
    print("%-44s levels %s  sum %.3f ms" % (name, " ".join("%.3f" % v for v in med), med.sum()), flush=True)
    ctx.close()
