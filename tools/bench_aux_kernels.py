"""GPU-box micro-benchmark of the memory-bound helper kernels (GB/s of algorithmic bytes).
python tools/bench_aux_kernels.py [rows cols]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd                      # noqa: E402
from realtimedepthdiffusion_b200.api import pitched_empty     # noqa: E402

rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
dev = "cuda"
ctx = rtdd.DepthDiffusion(rows, cols)
g = torch.Generator(device=dev).manual_seed(1)
bgr = pitched_empty(rows, cols, torch.uint8, dev, channels=3)
bgr.copy_(torch.randint(0, 256, (rows, cols * 3), generator=g, device=dev, dtype=torch.uint8))
gray = pitched_empty(rows, cols, torch.uint8, dev)
ctx.bgr2gray(bgr, gray)
depth = pitched_empty(rows, cols, torch.float32, dev)
depth.copy_(torch.rand((rows, cols), generator=g, device=dev) * 255)
scribble = pitched_empty(rows, cols, torch.uint8, dev, fill=0)
scribble.copy_((torch.rand((rows, cols), generator=g, device=dev) < 0.1).to(torch.uint8) * 255)
coarse = pitched_empty(rows // 2, cols // 2, torch.float32, dev)
coarse.copy_(torch.rand((rows // 2, cols // 2), generator=g, device=dev) * 255)
half_s = pitched_empty(rows // 2, cols // 2, torch.uint8, dev, fill=0)
half_e = pitched_empty(rows // 2, cols // 2, torch.uint8, dev, channels=3, fill=0)
gray2 = pitched_empty((rows + 1) // 2, (cols + 1) // 2, torch.uint8, dev)
u8 = pitched_empty(rows, cols, torch.uint8, dev)
outs = [pitched_empty(rows, cols, torch.uint8, dev, channels=3) for _ in range(3)]
px = rows * cols


def timeit(fn, reps=10):
    fn()
    ctx.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [
    ("pyrup_depth (4/4 B in, 4 B out per px)", lambda: ctx.pyrup_depth(coarse, depth), 5.0),
    ("convert_to_float (mask 1 + ~10% x (1 + 4))", lambda: ctx.convert_to_float(bgr, depth, scribble), 1.5),
    ("quantise_u8 (4 in, 1 out)", lambda: ctx.quantise_u8(depth, u8), 5.0),
    ("pyrdown_annotation (per coarse px: 4 mask)", lambda: ctx.pyrdown_annotation(scribble, bgr, half_s, half_e), 1.25),
    ("bgr2gray (3 in, 1 out)", lambda: ctx.bgr2gray(bgr, gray), 4.0),
    ("pyrdown_gray (1 in, 0.25 out)", lambda: ctx.pyrdown_gray(gray, gray2), 1.25),
    ("edge_weights/level_init (6 in, 7 out)", lambda: ctx.edge_weights_only(depth, gray, 0), 13.0),
    ("desaturation (8 in, 3 out)", lambda: ctx.simulate_desaturation(bgr, gray, depth, outs[0]), 11.0),
    ("haze (7 in, 3 out)", lambda: ctx.simulate_haze(bgr, depth, outs[1]), 10.0),
    ("defocus (7 in, 3 out + SAT)", lambda: ctx.simulate_defocus(bgr, depth, outs[2]), 10.0),
    ("effects fused (8 in, 9 out + SAT)", lambda: ctx.effects_fused(bgr, gray, depth, outs[0], outs[1], outs[2]), 17.0),
]
print("%dx%d" % (cols, rows))
# beyond a ~10K-pixel diagonal the defocus windows exceed 65 793 pixels and replay the reference's tap-by-tap fp32
# accumulation (bit-exactness): minutes per call at 16K -- never time that by accident
if (rows * rows + cols * cols) ** 0.5 * 0.025 > 256:
    cases = [c for c in cases if "SAT" not in c[0]]
    print("(defocus skipped: window sizes beyond the exact-integer range at this image size)")
for name, fn, bpp in cases:
    ms = timeit(fn)
    print("%-48s %8.4f ms  %7.0f GB/s" % (name, ms, bpp * px / ms / 1e6), flush=True)
ctx.close()
