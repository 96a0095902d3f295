"""GPU-box tuning of the resident kernel: warps per CTA (cluster width) per coarse level of the 4K / 16K pyramids."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402

for rows, cols, iters in ((67, 120, 1000), (135, 240, 500), (64, 64, 1000), (128, 128, 500), (256, 256, 250)):
    rng = np.random.default_rng(1)
    gray = synth.synth_image(rows, cols, 3)[..., 0].copy()
    depth = (rng.integers(0, 5, (rows, cols)) * 60 + rng.uniform(0, 14, (rows, cols))).astype(np.float32)
    scribble = np.where(rng.random((rows, cols)) < 0.1, 255, 0).astype(np.uint8)
    d0, s, g = to_dev(depth), to_dev(scribble), to_dev(gray)
    line = []
    for w in (4, 6, 8, 12, 16):
        ctx = rtdd.DepthDiffusion(rows, cols, 1)
        ctx.set_tuning("resident_warps", w)
        ctx.set_sweep_variant(3, 0)
        d = d0.clone()
        ms = []
        for rep in range(3):
            for _ in range(20):
                ctx.matrix_free_solver(d, s, g, iters, 0)
            ctx.sync()
            ms.append(ctx.level_sweep_ms(0)[0])
        line.append("w%d %.4f" % (w, float(np.median(ms))))
        ctx.set_tuning("resident_warps", 8)
        ctx.close()
    print("%dx%d x%d: " % (cols, rows, iters) + "  ".join(line), flush=True)
