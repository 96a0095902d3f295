"""GPU-box timing of coarse levels (64x64 ... 512x512) on two kinds of data:
  ordinary   random image, 10 % scribbles
  pockets    free pockets enclosed by depth-0 scribbles: they decay to zero and some end in a denormal limit cycle, so
             numerators stay below 2^-100 for the rest of the level (the resident kernel's div_tiny path)
and, per size, the cluster-resident kernel against the temporally blocked one.
python tools/coarse_level_cases.py > gpurun_out/coarse_cases.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402
from realtimedepthdiffusion_b200.api import pitched_empty   # noqa: E402


def to_dev(a):
    t = pitched_empty(a.shape[0], a.shape[1], torch.from_numpy(a[:1, :1]).dtype, torch.device("cuda"))
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t


def ordinary(rows, cols):
    rng = np.random.default_rng(1)
    gray = synth.synth_image(rows, cols, 3)[..., 0].copy()
    depth = np.full((rows, cols), 255.0, np.float32)
    scribble = np.where(rng.random((rows, cols)) < 0.1, 255, 0).astype(np.uint8)
    depth[scribble == 255] = rng.choice(np.array([0, 64, 128, 192, 254], np.float32), int((scribble == 255).sum()))
    return gray, depth, scribble


def pockets(rows, cols):
    rng = np.random.default_rng(2)
    gray, depth, scribble = ordinary(rows, cols)
    half = cols // 2
    scribble[:, :half] = 255
    depth[:, :half] = 0.0
    for _ in range(max(6, rows * half // 60)):
        h, w = rng.integers(1, 6), rng.integers(1, 6)
        y, x = rng.integers(1, max(2, rows - h - 1)), rng.integers(1, max(2, half - w - 1))
        scribble[y:y + h, x:x + w] = 0
        depth[y:y + h, x:x + w] = 255.0
    return gray, depth, scribble


for rows, cols, iters in ((64, 64, 1000), (67, 120, 1000), (128, 128, 500), (135, 240, 500), (256, 256, 250), (270, 480, 250), (512, 512, 125)):
    for name, make in (("ordinary", ordinary), ("pockets", pockets)):
        gray, depth, scribble = make(rows, cols)
        d0, s, g = to_dev(depth), to_dev(scribble), to_dev(gray)
        line = []
        for variant, T, label in ((3, 0, "resident"), (2, 11, "blocked T11"), (2, 12, "blocked T12")):
            ctx = rtdd.DepthDiffusion(rows, cols, 1)
            if T < 0:
                ctx.set_tuning("resident_r1_max_warps", -T)
                T = 0
            ctx.set_sweep_variant(variant, T)
            ms = []
            d = d0.clone()
            for rep in range(3):
                for _ in range(10):
                    d.copy_(d0)                      # same plane every time (one cached graph); legacy default stream orders it
                    ctx.matrix_free_solver(d, s, g, iters, 0)
                ctx.sync()
                ms.append(ctx.level_sweep_ms(0)[0])
            k = ctx.level_sweep_ms(0)[2]
            ctx.set_tuning("resident_r1_max_warps", 32)
            line.append("%s %.4f ms (%d launches, %.3f us/sweep)" % (label, float(np.median(ms)), k, 1e3 * float(np.median(ms)) / iters))
            ctx.close()
        print("%4dx%-4d x%-4d %-8s: " % (cols, rows, iters, name) + "  |  ".join(line), flush=True)
