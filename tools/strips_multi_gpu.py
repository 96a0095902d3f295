"""Row-strip solve of one large image on N GPUs (torchrun, NCCL) -- BASELINE configs[4].

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/strips_multi_gpu.py --size 16384 --steps 5 [--check]

Prints one JSON line from rank 0: ms per full-pyramid solve (max over ranks, CUDA events), ms of the split levels alone,
and with --check (sizes up to 4096) the bit-for-bit comparison with the single-GPU solve."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd                      # noqa: E402
from realtimedepthdiffusion_b200 import strips, synth          # noqa: E402
from realtimedepthdiffusion_b200.api import pitched_empty     # noqa: E402


def synth_on_device(rows, cols, seed, ctx, coverage=0.10):
    from realtimedepthdiffusion_b200 import synth_device
    return synth_device.synth_case_device(rows, cols, seed, ctx, coverage)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--halo", type=int, default=8)
    ap.add_argument("--pass-sweeps", type=int, default=0, help="sweeps per pass when smaller than --halo: several passes between two exchanges (NCCL mode)")
    ap.add_argument("--min-strip-pixels", type=int, default=1 << 22)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused", action="store_true", help="halo rows pushed by the sweep kernels over peer memory instead of NCCL send/recv")
    ap.add_argument("--staged", action="store_true", help="peer memory through staging rows: plain sweep kernels + small push / pull kernels (no NCCL on the data path)")
    ap.add_argument("--force-strip-path", action="store_true", help="one GPU, --level0-sweeps: go through rtdd_strip_* with the whole image as the window (isolates the strip path's own overhead)")
    ap.add_argument("--level0-sweeps", type=int, default=0,
                    help="SURVEY.md 8d config 5 (i): time only this many finest-level sweeps (strip-decomposed) instead of the whole pyramid")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    rows, cols = (args.rows or args.size), args.size
    ctx = rtdd.DepthDiffusion(rows, cols)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream)
    with torch.cuda.stream(stream):
        bgr, scribble, edited = synth_on_device(rows, cols, 1005, ctx)
        eng = strips.GpuStripEngine(ctx, bgr, scribble, edited)
        if (args.fused or args.staged) and world > 1:
            eng.enable_fused_halo_distributed(dist)
            if args.staged:
                eng.enable_staged_halo()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = args.level0_sweeps
        ps = args.pass_sweeps or None
        guess = None
        if l0 > 0:
            # the guess the finest level really starts from: one untimed whole-pyramid frame up to the prolongation
            if world > 1:
                strips.run_distributed(eng, dist, 1000, halo=args.halo, min_strip_pixels=args.min_strip_pixels, pass_sweeps=ps)
            else:
                strips.run_local([eng], 1000, halo=args.halo, min_strip_pixels=args.min_strip_pixels, pass_sweeps=ps)
            if world > 1:                                        # every rank needs all rows of the guess: take rank order
                rows0 = eng.depth[0].shape[0]
                b = [(r * rows0) // world for r in range(world)] + [rows0]
                for r in range(world):
                    part = eng.depth[0][b[r]:b[r + 1]]
                    buf = part.contiguous()
                    dist.broadcast(buf, src=r)
                    part.copy_(buf)
            guess = eng.depth[0].clone()

        def frame():
            if l0 > 0:
                eng.depth[0].copy_(guess)                        # outside the timed region
            else:
                for d in eng.depth:
                    d.fill_(255.0)                               # first-frame state (main.cpp:136), outside the timed region
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0.record(stream)
            if world > 1:
                plan, own, exchanges = strips.run_distributed(eng, dist, 1000, halo=args.halo, min_strip_pixels=args.min_strip_pixels, level0_sweeps=l0, pass_sweeps=ps)
            else:
                res, exchanges = strips.run_local([eng], 1000, halo=args.halo, min_strip_pixels=args.min_strip_pixels, level0_sweeps=l0, pass_sweeps=ps,
                                                  force_strip_path=args.force_strip_path)
                plan, own = res[0]
            ev1.record(stream)
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1), plan, own, exchanges

        for _ in range(args.warmup):
            frame()
        times = []
        for _ in range(args.steps):
            ms, plan, own, exchanges = frame()
            times.append(ms)
        ms = float(np.median(times))
        eng.marks = []
        frame()
        phases = eng.phase_ms()
        eng.marks = None
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = None
        if args.check:
            solo = rtdd.DepthDiffusion(rows, cols)
            solo.set_stream(stream)
            e1 = strips.GpuStripEngine(solo, bgr, scribble.clone(), edited.clone())
            if l0 > 0:
                e1.depth[0].copy_(guess)
            strips.run_local([e1], 1000, halo=args.halo, level0_sweeps=l0, pass_sweeps=ps)
            torch.cuda.synchronize()
            mine = eng.depth[0][own[0]:own[1]]
            ref = e1.depth[0][own[0]:own[1]]
            same = torch.equal(mine.view(torch.int32), ref.view(torch.int32))
            f = torch.tensor([1 if same else 0], device="cuda")
            if world > 1:
                dist.all_reduce(f, op=dist.ReduceOp.MIN)
            ok = bool(f.item())
    if rank == 0:
        total = 0
        L = ctx.levels
        per = []
        for l, (r, c) in enumerate(ctx.sizes):
            it = strips.level_iterations(1000, L, l)
            if l0 > 0:
                it = l0 if l == 0 else 0
            total += r * c * it
            per.append({"level": l, "size": "%dx%d" % (c, r), "sweeps": it, "split": plan[l] is not None})
        print(json.dumps({"workload": "configs[4]%s: %dx%d single synthetic image, row strips + NVLink halo exchange (%s), halo %d rows%s"
                                      % (" (i) finest level only, %d sweeps incl. edge-weight pass and result copy" % l0 if l0 > 0 else "",
                                         cols, rows, "peer-memory staging rows + flags" if args.staged else "peer-memory stores + flags" if args.fused else "NCCL send/recv", args.halo, (", FUSED: halo rows pushed by the sweep kernels over peer memory" if args.fused else "")
                                         + (", passes of %d sweeps" % ps if ps else "")),
                          "n_gpus": world, "ms_per_solve": float(t.item()), "Mpixel-sweeps/s": total / (float(t.item()) * 1e-3) / 1e6,
                          "pixel_sweeps": total, "halo_exchanges_per_solve": exchanges, "levels": per, "scaling": "strong",
                          "bit_identical_to_single_gpu": ok, "rank0_phase_ms": [[k, round(v, 4)] for k, v in phases]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
