"""Seeded synthetic inputs generated ON the device (torch), for workloads too large or too many to synthesise with numpy
inside a bench run: the 16384 x 16384 image of BASELINE configs[4] and the 256 images of configs[3].
Same recipe as synth.py (random rectangles / ellipses + Gaussian noise; random-walk brush strokes painted with the
reference's square brush through the library's own GPUPaintImage replacement), different random streams.
"""
import numpy as np
import torch

from . import synth
from .planes import pitched_empty


def synth_image_device(rows, cols, seed, dev, shapes=64, noise_sigma=4.0):
    """-> pitched BGR plane [rows, 3*cols] u8 on `dev` (identical on every rank for the same seed)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    plane = torch.empty((rows, cols, 3), dtype=torch.uint8, device=dev)
    plane[:] = torch.tensor(rng.integers(0, 256, 3), dtype=torch.uint8, device=dev)
    for _ in range(shapes):
        cy, cx = rng.uniform(0, rows), rng.uniform(0, cols)
        hy, hx = rng.uniform(0.03, 0.25) * rows, rng.uniform(0.03, 0.25) * cols
        colour = torch.tensor(rng.integers(0, 256, 3), dtype=torch.uint8, device=dev)
        y0, y1 = max(int(cy - hy), 0), min(int(cy + hy) + 1, rows)
        x0, x1 = max(int(cx - hx), 0), min(int(cx + hx) + 1, cols)
        if y1 <= y0 or x1 <= x0:
            continue
        if rng.random() < 0.5:
            plane[y0:y1, x0:x1] = colour
        else:
            yy = (torch.arange(y0, y1, device=dev, dtype=torch.float32)[:, None] - cy) / hy
            xx = (torch.arange(x0, x1, device=dev, dtype=torch.float32)[None, :] - cx) / hx
            sub = plane[y0:y1, x0:x1]
            sub[(yy * yy + xx * xx) <= 1.0] = colour
    step = max(1, (1 << 24) // cols)
    for r0 in range(0, rows, step):                             # noise in row chunks (bounded temporaries)
        r1 = min(rows, r0 + step)
        n = torch.randn((r1 - r0, cols, 1), generator=g, device=dev) * noise_sigma
        plane[r0:r1] = (plane[r0:r1].float() + n).round_().clamp_(0, 255).to(torch.uint8)
    bgr = pitched_empty(rows, cols, torch.uint8, dev, channels=3)
    bgr.copy_(plane.view(rows, cols * 3))
    return bgr


def synth_case_device(rows, cols, seed, ctx, coverage=0.10):
    """-> (bgr, scribble, edited) pitched device planes; strokes are painted by ctx.paint_image (the reference's brush)."""
    dev = ctx.device
    bgr = synth_image_device(rows, cols, seed, dev)
    scribble = pitched_empty(rows, cols, torch.uint8, dev, fill=0)
    edited = pitched_empty(rows, cols, torch.uint8, dev, channels=3)
    edited.copy_(bgr)
    radius = int(min(rows, cols) * 0.02)
    side = 2 * (radius // 2) + 1
    per_stroke = side * side + 23 * side * max(radius * 0.6, 1.0)
    nstrokes = max(int(coverage * rows * cols / per_stroke), 2)
    for (x, y, colour, rad) in synth.brush_events(rows, cols, seed, nstrokes, 24):
        ctx.paint_image(x, y, colour, rad, edited, scribble)
    ctx.sync()
    return bgr, scribble, edited


def annotation_plane_device(scribble, edited):
    """The single-plane annotation (32 = not annotated, ref: src/main.cpp:160-170) equivalent to (scribble, edited)."""
    rows, cols = scribble.shape
    e0 = edited.view(torch.uint8)[:, 0:3 * cols:3]
    return torch.where(scribble == 255, e0, torch.full_like(e0, 32)).contiguous()
