"""Row-strip solve of one large image on N GPUs through the NATIVE strip frame (rtdd_strip_frame_*: C++ frame loop, staged
peer-memory halo exchange, no NCCL on the data path) -- BASELINE configs[4].

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/strips_native.py --size 16384 --steps 5 [--level0-sweeps 64] [--check] [--effects]

Rank 0 prints one JSON line: ms per solve (max over ranks, CUDA events) and, with --check, whether every rank's own rows equal
the single-GPU frame bit for bit (fp32 depth, u8 map; with --effects also the three effects on the rank's rows)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtimedepthdiffusion_b200 as rtdd                                  # noqa: E402
from realtimedepthdiffusion_b200 import stripframe, synth_device            # noqa: E402
from realtimedepthdiffusion_b200.api import pitched_empty                   # noqa: E402


def run(args, world, rank, local):
    rows, cols = (args.rows or args.size), args.size
    sf = stripframe.StripFrameRank(rows, cols, rank, world, args.halo, args.pass_sweeps, args.min_strip_pixels)
    ctx = sf.ctx
    stream = torch.cuda.Stream()
    ctx.set_stream(stream)
    out = {}
    with torch.cuda.stream(stream):
        bgr, scribble, edited = synth_device.synth_case_device(rows, cols, args.seed, ctx)
        sf.set_image_device(bgr)
        sf.set_annotation_device(scribble, edited)
        if world > 1:
            sf.connect(dist)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = args.level0_sweeps

        def frame():
            sf.reset_first_frame_guess()                      # outside the timed region
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev0.record(stream)
            if l0 > 0:
                sf.level0(l0)
            else:
                sf.solve(1000)
            ev1.record(stream)
            ctx.sync()
            return ev0.elapsed_time(ev1)

        for _ in range(args.warmup):
            frame()
        ms = float(np.median([frame() for _ in range(args.steps)]))
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms"] = float(t.item())
        split, a, b, w0, w1 = sf.rows_of(0)
        ok = None
        if args.check:
            solo = rtdd.DepthDiffusion(rows, cols)
            solo.set_stream(stream)
            solo._ck(rtdd._native.lib.rtdd_frame_set_image_device(solo._h, rtdd.api._ptr(bgr), rtdd.api._pitch(bgr)))
            one = stripframe.StripFrameRank.__new__(stripframe.StripFrameRank)
            one.ctx, one.rank, one.world, one.rows, one.cols = solo, 0, 1, rows, cols
            solo._ck(rtdd._native.lib.rtdd_strip_frame_setup(solo._h, 0, 1, args.halo, args.pass_sweeps, args.min_strip_pixels))
            one.set_annotation_device(scribble, edited)
            one.reset_first_frame_guess()
            if l0 > 0:
                one.level0(l0)
            else:
                solo.frame_solve(1000)
            solo.sync()
            mine = sf.plane(ctx.PLANE_DEPTH)[a:b]
            ref = one.plane(solo.PLANE_DEPTH)[a:b]
            same = torch.equal(mine.contiguous().view(torch.int32), ref.contiguous().view(torch.int32))
            if l0 == 0:
                same = same and torch.equal(sf.plane(ctx.PLANE_DEPTH_U8)[a:b], one.plane(solo.PLANE_DEPTH_U8)[a:b])
            if args.effects and l0 == 0:
                outs = [pitched_empty(rows, cols, torch.uint8, "cuda", channels=3, fill=0) for _ in range(6)]
                sf.effects(*outs[:3])
                solo.frame_effects(*outs[3:])
                ctx.sync()
                solo.sync()
                for k in range(3):
                    same = same and torch.equal(outs[k][a:b], outs[3 + k][a:b])
            f = torch.tensor([1 if same else 0], device="cuda")
            if world > 1:
                dist.all_reduce(f, op=dist.ReduceOp.MIN)
            ok = bool(f.item())
            solo.close()
        out["ok"] = ok
        out["split0"] = split
        out["levels"] = ctx.levels
        out["sizes"] = ctx.sizes
        out["split"] = [sf.rows_of(l)[0] for l in range(ctx.levels)]
    sf.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--halo", type=int, default=16)
    ap.add_argument("--pass-sweeps", type=int, default=8)
    ap.add_argument("--min-strip-pixels", type=int, default=1 << 22)
    ap.add_argument("--seed", type=int, default=1005)
    ap.add_argument("--level0-sweeps", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--effects", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    out = run(args, world, rank, local)
    if rank == 0:
        L = out["levels"]
        total = 0
        for l, (r, c) in enumerate(out["sizes"]):
            it = int(np.float32(1000) / np.float32(2.0 ** ((L - 1) - l)))
            if args.level0_sweeps > 0:
                it = args.level0_sweeps if l == 0 else 0
            total += r * c * it
        print(json.dumps({"workload": "configs[4]%s: %dx%d single synthetic image, row strips through rtdd_strip_frame_* (C++ frame loop, staged "
                                      "peer-memory halo exchange), halo %d rows, passes of %d sweeps"
                                      % (" (i) finest level only, %d sweeps incl. edge-weight pass" % args.level0_sweeps if args.level0_sweeps else "",
                                         args.size, args.rows or args.size, args.halo, args.pass_sweeps),
                          "n_gpus": world, "ms_per_solve": out["ms"], "Mpixel-sweeps/s": total / (out["ms"] * 1e-3) / 1e6, "pixel_sweeps": total,
                          "split_levels": out["split"], "scaling": "strong", "bit_identical_to_single_gpu": out["ok"]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
