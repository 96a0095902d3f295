"""Small, fast pass through every kernel of librtdd.so for compute-sanitizer (one tool per gpurun call):
compute-sanitizer --tool memcheck python tools/sanitizer_case.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import strips, synth       # noqa: E402
from realtimedepthdiffusion_b200.api import to_dev   # noqa: E402

rows, cols = 203, 317
bgr, scribble, edited = synth.synth_case(rows, cols, 5)
out = np.zeros((rows, cols), np.uint8)
for variant, T, tile, tma in ((1, 0, 0, 1), (2, 5, 64, 1), (2, 7, 64, 3), (2, 8, 64, 0), (2, 7, 34, 0), (2, 4, 32, 0), (3, 0, 0, 1), (0, 0, 0, 2)):
    ctx = rtdd.DepthDiffusion(rows, cols)
    ctx.set_tuning("blocked_tile", tile)
    ctx.set_tuning("blocked_tma", tma)
    ctx.set_sweep_variant(variant, T)
    ctx.frame_set_image(bgr)
    ctx.frame_solve_host(scribble, edited, 40, out)
    ctx.frame_paint(100, 50, 128, 9)
    ctx.frame_solve(40)
    ctx.frame_solve_incremental(40, 1)
    ctx.frame_solve_band(40, 40, 60, 8)
    ctx.sync()
    print(variant, T, tile, tma, float(out.mean()), ctx.level_residual(0))
    ctx.set_tuning("blocked_tile", 0)
    ctx.set_tuning("blocked_tma", 2)
    ctx.close()
ctx = rtdd.DepthDiffusion(rows, cols)
o, g, d = to_dev(bgr, 3), to_dev(np.ascontiguousarray(bgr[..., 1])), to_dev(np.full((rows, cols), 120.0, np.float32))
outs = [to_dev(np.zeros_like(bgr), 3) for _ in range(3)]
ctx.simulate_desaturation(o, g, d, outs[0])
ctx.simulate_haze(o, d, outs[1])
ctx.simulate_defocus(o, d, outs[2])
ctx.effects_fused(o, g, d, outs[0], outs[1], outs[2])
n, res = ctx.matrix_free_solver_converge(d, to_dev(scribble), g, 40, 0.5, 0, check_every=8)
print("converge", n, res, ctx.selftest_division(1 << 16))
ctx.close()
engines = [strips.GpuStripEngine(rtdd.DepthDiffusion(rows, cols), to_dev(bgr, 3), to_dev(scribble), to_dev(edited, 3)) for _ in range(2)]
res, ex = strips.run_local(engines, 40, halo=4, min_strip_pixels=1)
torch.cuda.synchronize()
print("strips", ex)
for e in engines:
    e.ctx.close()
print("SANITIZER_CASE_DONE")
