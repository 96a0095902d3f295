"""GPU-box experiment: aggregate throughput of K independent solve contexts on K streams of ONE GPU (configs[3]: batch of
images).  python tools/concurrent_frames.py [workload] [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from bench import WORKLOADS, pixel_sweeps          # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rows, cols, seed = WORKLOADS[name]
out = np.zeros((rows, cols), np.uint8)
for K in (1, 2, 3, 4, 6):
    ctxs, streams = [], []
    for k in range(K):
        bgr, scribble, edited = synth.synth_case(rows, cols, seed + k)
        ctx = rtdd.DepthDiffusion(rows, cols)
        st = torch.cuda.Stream()
        ctx.set_stream(st)
        ctx.frame_set_image(bgr)
        ctx.frame_solve_host(scribble, edited, 1000, out)
        ctxs.append(ctx)
        streams.append(st)
    total_ps, _ = pixel_sweeps(rows, cols, ctxs[0].levels)
    for ctx in ctxs:
        ctx.frame_solve(1000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    for f in range(frames):
        ctxs[f % K].frame_solve(1000)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("K=%d: %d frames in %.3f ms -> %.3f ms/frame, %.1f Gpixel-sweeps/s" % (K, frames, ms, ms / frames, total_ps * frames / ms / 1e6), flush=True)
    for ctx in ctxs:
        ctx.close()
