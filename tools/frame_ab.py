"""GPU-box A/B of one library tuning key on the whole solve frame: ms per frame (device-resident, CUDA events, median).
python tools/frame_ab.py fused_prolong 0 1 [4k 1080p 720p]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from bench import WORKLOADS                         # noqa: E402

key = sys.argv[1]
values = [int(sys.argv[2]), int(sys.argv[3])]
names = sys.argv[4:] or ["4k", "1080p"]
for name in names:
    rows, cols, seed = WORKLOADS[name]
    bgr, scribble, edited = synth.synth_case(rows, cols, seed)
    out = np.zeros((rows, cols), np.uint8)
    line = []
    for rep in range(2):
        for v in values:
            ctx = rtdd.DepthDiffusion(rows, cols)
            ctx.set_tuning(key, v)
            ctx.frame_set_image(bgr)
            ctx.frame_solve_host(scribble, edited, 1000, out)
            for _ in range(5):
                ctx.frame_solve(1000)
            ctx.sync()
            ms = []
            for _ in range(30):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.frame_solve(1000)
                e1.record()
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            line.append("%s=%d %.4f ms" % (key, v, float(np.median(ms))))
            ctx.close()
    print(name, " | ".join(line), flush=True)
