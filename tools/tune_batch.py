"""GPU-box scan for BASELINE configs[3] on one GPU: N synthetic 1080p images through K contexts in flight, for pass plans of the two
finest levels (rtdd_set_pass_plan) and tiling forms -- what the "plan_throughput" objective of the pass planner should reproduce."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd          # noqa: E402
from realtimedepthdiffusion_b200 import synth       # noqa: E402

rows, cols, N, K = 1080, 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 6
ONLY_PLANNERS = len(sys.argv) > 3 and sys.argv[3] == "planners"
cases = []
for i in range(8):
    bgr, scribble, edited = synth.synth_case(rows, cols, 2000 + i)
    cases.append((torch.from_numpy(bgr).pin_memory(), torch.from_numpy(synth.annotation_plane(scribble, edited)).pin_memory()))


def run(name, setup):
    ctxs = []
    for k in range(K):
        c = rtdd.DepthDiffusion(rows, cols)
        st = torch.cuda.Stream()
        c.set_stream(st)
        setup(c)
        ctxs.append((c, st, torch.empty((rows, cols), dtype=torch.uint8).pin_memory()))
    for k, (c, st, ho) in enumerate(ctxs):
        c.frame_set_image(cases[k % 8][0])
        c.frame_solve_host_annotation(cases[k % 8][1], 1000, None)
        c.frame_read_depth_u8(ho, sync=True)
    best = 1e9
    for rep in range(2):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _, st, _ in ctxs:
            st.wait_event(ev0)
        for j in range(N):
            c, st, ho = ctxs[j % K]
            c.frame_set_image(cases[j % 8][0], sync=False)
            c.frame_solve_host_annotation(cases[j % 8][1], 1000, None)
            c.frame_read_depth_u8(ho, sync=False)
        for _, st, _ in ctxs:
            torch.cuda.current_stream().wait_stream(st)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    for c, _, _ in ctxs:
        c.set_tuning("blocked_tma", 2)
        c.close()
    print("%-58s %.3f ms per image" % (name, best / N), flush=True)


def plans(l0, l1, tma=2):
    def f(c):
        c.set_tuning("blocked_tma", tma)
        if l0:
            c.set_pass_plan(0, l0)
        if l1:
            c.set_pass_plan(1, l1)
    return f


def rep(t, total):
    p = [t] * (total // t)
    if total % t:
        p = [total % t] + p
    return p


run("latency planner", lambda c: None)
run("throughput planner", lambda c: c.set_tuning("plan_throughput", 1))
if ONLY_PLANNERS:
    sys.exit(0)


def level2(plan, tma):
    def f(c):
        c.set_tuning("plan_throughput", 1)
        c.set_tuning("blocked_tma", tma)
        c.set_pass_plan(2, plan)
    return f


if len(sys.argv) > 3 and sys.argv[3] == "level2":
    for tma, nm in ((3, "clusters"), (1, "single CTAs")):
        for t2 in (6, 8, 10, 12, 16):
            run("throughput plans; level 2 (480x270) %s, passes of %d" % (nm, t2), level2(rep(t2, 250), tma))
    sys.exit(0)
for tma, nm in ((3, "clusters"), (1, "single CTAs")):
    for t0 in (6, 8, 10, 12):
        run("%s, level 0 in passes of %d" % (nm, t0), plans(rep(t0, 62), None, tma))
for tma, nm in ((3, "clusters"), (1, "single CTAs")):
    for t1 in (8, 10, 12, 13, 16):
        run("level 0 throughput plan; level 1 %s, passes of %d" % (nm, t1), plans([8, 8, 8, 8, 8, 11, 11] if tma == 3 else rep(8, 62), rep(t1, 125), tma))
