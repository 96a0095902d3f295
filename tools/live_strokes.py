"""BASELINE configs[1]: 1920x1080 synthetic image, a sequence of brush strokes, one solve frame per stroke.
Reports ms/frame for the parity path (full fixed-schedule solve, as main.cpp --live does) and for the opt-in extensions --
the warm start that skips coarse levels (rtdd_frame_solve_incremental) and the band re-solve around the stroke
(rtdd_frame_solve_band, dilation D rows per level) -- each with its deviation from the parity frame of the same stroke sequence
(every mode carries its OWN state from frame to frame, so deviations accumulate over the 32 frames).
python tools/live_strokes.py [strokes] > gpurun_out/live.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

rows, cols, seed = 1080, 1920, 1002
nstrokes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bgr, scribble, edited = synth.synth_case(rows, cols, seed, strokes=8)
events = synth.brush_events(rows, cols, seed + 1, nstrokes, 1)          # one brush event per stroke frame
out = np.zeros((rows, cols), np.uint8)
res = {"workload": "configs[1]: 1920x1080 synthetic image, %d live brush events, one solve frame per event" % nstrokes}
ctxs = {}
pinned = torch.zeros((rows, cols), dtype=torch.uint8).pin_memory()
for mode in ("parity", "parity_with_download", "incremental_L2", "incremental_L1", "band_D24", "band_D48", "band_D96"):
    ctx = rtdd.DepthDiffusion(rows, cols)
    ctx.frame_set_image(bgr)
    ctx.frame_solve_host(scribble, edited, 1000, out)
    ctxs[mode] = ctx
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
times = {m: [] for m in ctxs}
delta = {m: [] for m in ctxs}
ident = {m: [] for m in ctxs}
for (x, y, colour, radius) in events:
    for mode, ctx in ctxs.items():
        ev0.record()
        ctx.frame_paint(x, y, colour, radius)
        if mode == "parity":
            ctx.frame_solve(1000)
        elif mode == "parity_with_download":
            ctx.frame_solve_download(pinned, 1000)             # main.cpp:291 included: the map lands in pinned host memory
        elif mode.startswith("band"):
            h = radius // 2
            ctx.frame_solve_band(1000, max(y - h, 0), min(y + h + 1, rows), int(mode.split("_D")[1]))
        else:
            ctx.frame_solve_incremental(1000, int(mode[-1]))
        ev1.record()
        ev1.synchronize()
        times[mode].append(ev0.elapsed_time(ev1))
    ref = ctxs["parity"].frame_plane(0, 0)
    refq = ctxs["parity"].frame_plane(5, 0)
    for mode, ctx in ctxs.items():
        d = ctx.frame_plane(0, 0)
        delta[mode].append(float((d - ref).abs().mean()))
        ident[mode].append(float((ctx.frame_plane(5, 0) == refq).float().mean()))
for mode in ctxs:
    res[mode] = {"ms_per_frame_median": float(np.median(times[mode])), "mean_abs_depth_delta_vs_parity": float(np.mean(delta[mode])),
                 "identical_8bit_fraction_vs_parity": float(np.mean(ident[mode])),
                 "identical_8bit_fraction_last_frame": float(ident[mode][-1]), "mean_abs_depth_delta_last_frame": float(delta[mode][-1])}
print(json.dumps(res))
