"""Seeded synthetic inputs for the bench and the parity tests.

Images: piecewise-constant random rectangles / ellipses plus Gaussian noise, BGR u8.
Scribbles: random-walk brush strokes (the reference's square brush, src/GPUImageProcessing.cu:51-70)
with depths drawn from the values main.cpp can paint ({0, 64, 128, 192, 254}, src/main.cpp:41-42).
"""
import numpy as np

DEPTH_VALUES = np.array([0, 64, 128, 192, 254], np.uint8)


def synth_image(rows, cols, seed, shapes=64, noise_sigma=4.0):
    rng = np.random.default_rng(seed)
    img = np.empty((rows, cols, 3), np.float32)
    img[:] = rng.integers(0, 256, 3).astype(np.float32)
    yy = np.arange(rows, dtype=np.float32)[:, None]
    xx = np.arange(cols, dtype=np.float32)[None, :]
    for _ in range(shapes):
        cy, cx = rng.uniform(0, rows), rng.uniform(0, cols)
        hy, hx = rng.uniform(0.03, 0.25) * rows, rng.uniform(0.03, 0.25) * cols
        colour = rng.integers(0, 256, 3).astype(np.float32)
        y0, y1 = max(int(cy - hy), 0), min(int(cy + hy) + 1, rows)
        x0, x1 = max(int(cx - hx), 0), min(int(cx + hx) + 1, cols)
        if y1 <= y0 or x1 <= x0:
            continue
        if rng.random() < 0.5:
            img[y0:y1, x0:x1] = colour
        else:
            m = ((yy[y0:y1] - cy) / hy) ** 2 + ((xx[:, x0:x1] - cx) / hx) ** 2 <= 1.0
            img[y0:y1, x0:x1][m] = colour
    noise = rng.standard_normal((rows, cols), dtype=np.float32) * noise_sigma
    img += noise[..., None]
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def brush_events(rows, cols, seed, strokes, steps_per_stroke=24):
    """-> list of (x, y, colour, radius) brush events, radius = int(0.02*min(rows, cols)) (src/main.cpp:154)."""
    rng = np.random.default_rng(seed + 7919)
    radius = int(min(rows, cols) * 0.02)
    ev = []
    for _ in range(strokes):
        x, y = rng.uniform(0, cols), rng.uniform(0, rows)
        ang = rng.uniform(0, 2 * np.pi)
        colour = int(DEPTH_VALUES[rng.integers(0, len(DEPTH_VALUES))])
        step = max(radius * 0.6, 1.0)
        for _ in range(steps_per_stroke):
            ev.append((int(x), int(y), colour, radius))
            ang += rng.normal(0, 0.35)
            x = min(max(x + step * np.cos(ang), 0), cols - 1)
            y = min(max(y + step * np.sin(ang), 0), rows - 1)
    return ev


def paint_events(bgr, events, scribble=None, edited=None):
    """Apply brush events on the host exactly like paintImage: square side 2*(r//2)+1."""
    rows, cols = bgr.shape[:2]
    if scribble is None:
        scribble = np.zeros((rows, cols), np.uint8)
    if edited is None:
        edited = bgr.copy()      # main.cpp:158 -- editedImage[0] = imread(input)
    for (x, y, colour, radius) in events:
        h = radius // 2
        y0, y1 = max(y - h, 0), min(y + h, rows - 1)
        x0, x1 = max(x - h, 0), min(x + h, cols - 1)
        if y1 < y0 or x1 < x0:
            continue
        edited[y0:y1 + 1, x0:x1 + 1] = colour
        scribble[y0:y1 + 1, x0:x1 + 1] = 255
    return scribble, edited


def synth_case(rows, cols, seed, strokes=None, coverage=0.10):
    """Image + scribbles covering roughly `coverage` of the pixels."""
    bgr = synth_image(rows, cols, seed)
    radius = int(min(rows, cols) * 0.02)
    side = 2 * (radius // 2) + 1
    steps = 24
    if strokes is None:
        per_stroke = side * side + (steps - 1) * side * max(radius * 0.6, 1.0)
        strokes = max(int(coverage * rows * cols / per_stroke), 2)
    ev = brush_events(rows, cols, seed, strokes, steps)
    scribble, edited = paint_events(bgr, ev)
    return bgr, scribble, edited


def annotation_plane(scribble, edited):
    """The single-plane annotation main.cpp would have loaded to arrive at (scribble, edited): the painted value where
    scribble == 255, 32 elsewhere (ref: src/main.cpp:160-170; none of the paintable depths is 32)."""
    return np.where(scribble == 255, edited[..., 0], 32).astype(np.uint8)
