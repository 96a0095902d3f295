"""GPU-box tuning sweep for the temporally blocked kernel: per level size, sweeps-per-pass T and tile shape.
python tools/tune_blocked.py > gpurun_out/tune.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402

LEVELS = [(2160, 3840, 31), (1080, 1920, 62), (540, 960, 125), (270, 480, 250), (135, 240, 500)]
if len(sys.argv) > 1:
    LEVELS = [LEVELS[int(a)] for a in sys.argv[1:]]
for rows, cols, iters in LEVELS:
    rng = np.random.default_rng(1)
    gray = synth.synth_image(rows, cols, 3)[..., 0].copy()
    depth = (rng.integers(0, 5, (rows, cols)) * 60 + rng.uniform(0, 14, (rows, cols))).astype(np.float32)
    scribble = np.where(rng.random((rows, cols)) < 0.1, 255, 0).astype(np.uint8)
    d0, s, g = to_dev(depth), to_dev(scribble), to_dev(gray)
    best = None
    for tile in (64, 32):
        for T in (3, 4, 5, 6, 7, 8, 10, 12, 14, 16):
            if tile == 32 and 2 * T >= 32:
                continue
            ctx = rtdd.DepthDiffusion(rows * 2, cols * 2, 2)
            ctx.set_tuning("blocked_tile", tile)
            ctx.set_sweep_variant(2, T)
            # keep the GPU busy (steady clocks): enqueue many solves back to back, time the last ones
            ms = []
            d = d0.clone()
            for rep in range(3):
                for _ in range(25):
                    ctx.matrix_free_solver(d, s, g, iters, 1)
                ctx.sync()
                ms.append(ctx.level_sweep_ms(1)[0])
            k = ctx.level_sweep_ms(1)[2]
            m = float(np.median(ms))
            print("%dx%d iters %d tile %d T %2d : %.4f ms  (%d launches, %.1f Gpx-sweeps/s)" % (cols, rows, iters, tile, T, m, k, rows * cols * iters / m / 1e6), flush=True)
            if best is None or m < best[0]:
                best = (m, tile, T)
            ctx.set_tuning("blocked_tile", 0)
            ctx.close()
    print("BEST %dx%d: %.4f ms tile %d T %d" % (cols, rows, best[0], best[1], best[2]), flush=True)
