"""GPU-box A/B of the blocked kernel forms on whole frames: per-level sweep ms and the frame ms for
blocked_tma in {1 (single CTAs), 2 (clusters)} x blocked_cluster in {1, 2, 4} x sweeps per pass."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

sizes = {"4k": (2160, 3840, 1003), "1080p": (1080, 1920, 1002)}
which = sys.argv[1:] or ["4k"]
for name in which:
    rows, cols, seed = sizes[name]
    bgr, scribble, edited = synth.synth_case(rows, cols, seed)
    ref = None
    combos = ((1, 1, 0), (3, 1, 0), (3, 2, 0), (3, 4, 0), (3, 2, 6), (3, 2, 7), (3, 2, 10), (3, 4, 11), (3, 2, 16), (3, 4, 16), (2, 2, 0))
    if os.environ.get("RTDD_QUICK"):
        combos = ((1, 1, 0), (3, 1, 0), (3, 2, 0), (2, 2, 0))
    for mode, cl, T in combos:
        ctx = rtdd.DepthDiffusion(rows, cols)
        ctx.set_tuning("blocked_tma", mode)
        ctx.set_tuning("blocked_cluster", cl)
        ctx.set_sweep_variant(0, T)
        ctx.frame_set_image(bgr)
        u8 = ctx.frame_solve_host(scribble, edited, 1000, np.zeros((rows, cols), np.uint8)).numpy().copy()
        if ref is None:
            ref = u8
        same = bool(np.array_equal(ref, u8))
        for _ in range(5):
            ctx.frame_solve(1000)
        ctx.sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(20):
            ctx.frame_solve(1000)
        ev1.record()
        torch.cuda.synchronize()
        lv = [ctx.level_sweep_ms(l) for l in range(ctx.levels)]
        print("%s mode %d cluster %d T %2d: frame %.4f ms  same=%s  levels %s" % (name, mode, cl, T, ev0.elapsed_time(ev1) / 20, same,
              " ".join("%.3f(%d)" % (m, k) for m, _, k in lv)), flush=True)
        ctx.set_tuning("blocked_tma", 2)
        ctx.set_tuning("blocked_cluster", 2)
        ctx.close()
