"""GPU-box experiment: cluster-resident kernel, one row per warp vs two rows per warp, warps per CTA, for the coarse levels of
the 4K / 16K pyramids (sizes x sweeps as in the frames)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402

for rows, cols, iters in ((67, 120, 1000), (135, 240, 500), (64, 64, 1000), (128, 128, 500)):
    rng = np.random.default_rng(1)
    gray = synth.synth_image(rows, cols, 3)[..., 0].copy()
    depth = (rng.integers(0, 5, (rows, cols)) * 60 + rng.uniform(0, 14, (rows, cols))).astype(np.float32)
    scribble = np.where(rng.random((rows, cols)) < 0.1, 255, 0).astype(np.uint8)
    d0, s, g = to_dev(depth), to_dev(scribble), to_dev(gray)
    line = []
    for r1max in (32, 2):
        for w in (2, 4, 6, 8, 12, 16):
            ctx = rtdd.DepthDiffusion(rows, cols, 1)
            ctx.set_tuning("resident_r1_max_warps", r1max)
            ctx.set_tuning("resident_warps", w)
            ctx.set_sweep_variant(3, 0)
            d = d0.clone()
            ms = []
            try:
                for rep in range(3):
                    for _ in range(20):
                        ctx.matrix_free_solver(d, s, g, iters, 0)
                    ctx.sync()
                    ms.append(ctx.level_sweep_ms(0)[0])
                line.append("r1max%d w%d %.4f" % (r1max, w, float(np.median(ms))))
            except Exception as e:
                line.append("r1max%d w%d ERR" % (r1max, w))
            ctx.set_tuning("resident_warps", 8)
            ctx.set_tuning("resident_r1_max_warps", 32)
            ctx.close()
    print("%dx%d x%d: " % (cols, rows, iters) + "  ".join(line), flush=True)
