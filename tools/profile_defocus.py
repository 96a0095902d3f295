"""One stand-alone defocus call at 4K (summed-area table build + lookup) for an ncu launch list."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402

rows, cols = 2160, 3840
bgr = synth.synth_image(rows, cols, 5)
rng = np.random.default_rng(1)
depth = (rng.uniform(0, 255, (rows // 8, cols // 8)).astype(np.float32)).repeat(8, 0).repeat(8, 1)
ctx = rtdd.DepthDiffusion(rows, cols)
o, d = to_dev(bgr, 3), to_dev(depth)
out = to_dev(np.zeros_like(bgr), 3)
for _ in range(3):
    ctx.simulate_defocus(o, d, out)
ctx.sync()
print("ok")
