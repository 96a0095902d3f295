"""One solve frame (plus one of each effect) for ncu: python tools/profile_frame.py [workload] [frames]."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from bench import WORKLOADS                         # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows, cols, seed = WORKLOADS[name]
bgr, scribble, edited = synth.synth_case(rows, cols, seed)
ctx = rtdd.DepthDiffusion(rows, cols)
if "RTDD_TMA" in os.environ:
    ctx.set_tuning("blocked_tma", int(os.environ["RTDD_TMA"]))
if "RTDD_CLUSTER" in os.environ:
    ctx.set_tuning("blocked_cluster", int(os.environ["RTDD_CLUSTER"]))
ctx.frame_set_image(bgr)
out = np.zeros((rows, cols), np.uint8)
for _ in range(frames):
    ctx.frame_solve_host(scribble, edited, 1000, out)
print("levels", ctx.levels, "launches", ctx.launch_count, "mean depth", float(out.mean()))
for l in range(ctx.levels):
    print(l, ctx.level_sweep_ms(l))
ctx.close()
