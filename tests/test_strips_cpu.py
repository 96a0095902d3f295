"""Row-strip decomposition on the CPU: planning, lockstep emulation, and a real world_size-2 gloo run.
The arithmetic is the oracle's; what is under test is realtimedepthdiffusion_b200/strips.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import binding as ob
from realtimedepthdiffusion_b200 import strips, synth
from tests.strip_cpu_engine import CpuStripEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_strips_geometry():
    sizes = [(16384 >> l, 16384 >> l) for l in range(9)]
    for n in (2, 4, 8):
        plan = strips.plan_strips(sizes, n, 8)
        split = [l for l in range(9) if plan[l] is not None]
        assert split == [0, 1, 2, 3]                      # >= 4 M pixels
        for l in split:
            rows = sizes[l][0]
            assert plan[l][0][0] == 0 and plan[l][-1][1] == rows
            for r in range(n - 1):
                assert plan[l][r][1] == plan[l][r + 1][0]
            if l > 0 and plan[l - 1] is not None:
                for r in range(n):
                    assert plan[l - 1][r][0] == 2 * plan[l][r][0]
    assert all(p is None for p in strips.plan_strips(sizes, 1, 8))
    # odd sizes: the last strip absorbs the extra row
    sizes = [(1081, 700), (540, 350), (270, 175)]
    plan = strips.plan_strips(sizes, 3, 4, min_strip_pixels=1)
    assert plan[0][-1][1] == 1081 and plan[1][-1][1] == 540 and plan[2][-1][1] == 270
    assert plan[0][1][0] == 2 * plan[1][1][0] == 4 * plan[2][1][0]


def _plan_restated(sizes, nranks, halo, min_strip_pixels):
    """The planning rule written out in Python, to pin the native rtdd_plan_strips."""
    levels = len(sizes)
    plan = [None] * levels
    cs = -1
    for l in range(levels):
        rows, cols = sizes[l]
        if rows * cols >= min_strip_pixels and rows // nranks >= max(halo, 2):
            cs = l
        else:
            break
    if nranks <= 1 or cs < 0:
        return plan
    bounds = [(r * sizes[cs][0]) // nranks for r in range(nranks)]
    for l in range(cs, -1, -1):
        if l < cs:
            bounds = [2 * b for b in bounds]
        full = bounds + [sizes[l][0]]
        plan[l] = [(full[r], full[r + 1]) for r in range(nranks)]
    return plan


def _schedule_restated(iters, halo, pass_sweeps, level):
    T = halo if not pass_sweeps or pass_sweeps > halo else pass_sweeps
    out, k, since = [], 0, 0
    while k < iters:
        n = min(T, iters - k, halo - since)
        k += n
        since += n
        ex = (since >= halo or k >= iters) and (k < iters or level > 0)
        out.append((n, ex))
        if ex:
            since = 0
    return out


def test_native_planning_matches_its_restatement():
    rng = np.random.default_rng(11)
    for _ in range(300):
        levels = int(rng.integers(1, 9))
        rows, cols = int(rng.integers(64, 20000)), int(rng.integers(64, 20000))
        sizes = [(rows >> l, cols >> l) for l in range(levels) if (rows >> l) > 0 and (cols >> l) > 0]
        nranks, halo = int(rng.integers(1, 9)), int(rng.integers(1, 33))
        mp = int(rng.choice([1, 1 << 16, 1 << 22]))
        want = _plan_restated(sizes, nranks, halo, mp)
        if any(p is not None and any(e - b < halo for b, e in p) for p in want):
            continue                                       # the native planner refuses strips shorter than the halo
        assert strips.plan_strips(sizes, nranks, halo, mp) == want
    for _ in range(300):
        iters, halo = int(rng.integers(0, 1100)), int(rng.integers(1, 33))
        ps = int(rng.integers(0, 40))
        level = int(rng.integers(0, 3))
        got = strips.strip_schedule(iters, halo, ps, level)
        assert got == _schedule_restated(iters, halo, ps, level)
        assert sum(n for n, _ in got) == iters
    # one exchange per `halo` sweeps; level 0 needs none after its last pass, other levels do
    assert [e for _, e in strips.strip_schedule(64, 16, 8, 0)] == [False, True, False, True, False, True, False, False]
    assert [e for _, e in strips.strip_schedule(31, 8, 0, 3)] == [True, True, True, True]


def test_strip_plane_rows_are_the_largest_window_and_equal_on_every_rank():
    """rtdd_create_strip sizes the scratch planes of split levels by rtdd_plan_strip_planes: one layout for all ranks (peers address
    each other's planes by offset), large enough for every rank's window, full size for the levels every rank solves whole."""
    import ctypes as C
    from realtimedepthdiffusion_b200._native import lib
    rows = cols = 16384
    levels = 9
    sizes = [(int(rows / 2.0 ** l), int(cols / 2.0 ** l)) for l in range(levels)]
    for nranks, halo in ((8, 16), (2, 16), (4, 8), (1, 16)):
        IntArr = C.c_int * levels
        lr, lc = IntArr(*[s[0] for s in sizes]), IntArr(*[s[1] for s in sizes])
        plane = IntArr()
        assert lib.rtdd_plan_strip_planes(lr, lc, levels, nranks, halo, 1 << 22, plane) == 0
        plan = strips.plan_strips(sizes, nranks, halo, 1 << 22)
        for l in range(levels):
            if plan[l] is None:
                assert plane[l] == sizes[l][0]
            else:
                windows = [min(sizes[l][0], b + halo) - max(0, a - halo) for a, b in plan[l]]
                assert plane[l] == max(windows) and plane[l] < sizes[l][0]
    # 16384^2 on 8 GPUs: level 0 keeps 2048 + 2 x 16 rows of 16384
    assert lib.rtdd_plan_strip_planes(lr, lc, levels, 8, 16, 1 << 22, plane) == 0 and plane[0] == 2080


def test_blocked_pass_planner_reproduces_the_measured_optima():
    """rtdd_plan_blocked (host only): sweeps per pass and form of a large level.  The model behind it was fitted to per-level scans
    on a B200 (profiles/r02_tune_levels.txt); these are the measured optima it must keep reproducing."""
    import ctypes as C
    from realtimedepthdiffusion_b200._native import lib
    T, cl = C.c_int(), C.c_int()

    def plan(r, c, it):
        assert lib.rtdd_plan_blocked(r, c, it, 148, C.byref(T), C.byref(cl)) == 0
        return T.value, cl.value
    assert plan(2160, 3840, 31) == (7, 1)            # 4K level 0: clusters of 2, 7 sweeps per pass (0.489 ms; 8: 0.500; single CTAs: 0.536)
    assert plan(1080, 1920, 62) == (16, 1)           # 4K level 1 / 1080p level 0: clusters, 16 (0.262 ms; single CTAs at 8: 0.295)
    assert plan(540, 960, 125) == (13, 0)            # single CTAs, 13 (0.188 ms; clusters at 16: 0.199)
    for r, c, it in ((853, 1280, 62), (4320, 7680, 15), (16384, 16384, 3), (2080, 16384, 64), (512, 512, 125)):
        t, f = plan(r, c, it)
        assert 4 <= t <= 16 and f in (0, 1)
    assert lib.rtdd_plan_blocked(0, 10, 5, 148, C.byref(T), C.byref(cl)) == -1


def test_pass_planner_with_passes_of_their_own_lengths():
    """rtdd_plan_passes (host only): what the level driver runs by default -- every pass with the halo of its own length.  Pinned to
    the plans measured on a B200 (profiles/r02_tune_passes.txt): 3840x2160 x 31 as (7, 7, 7, 10) 0.466 ms against 0.484 for 4 x 7 + 3;
    with the 8-bit map stored into pinned host memory by the last pass, that pass is as long as the tiling allows (end to end
    2.01 ms with (7, 8, 16) against 2.07 with (7, 7, 7, 10))."""
    import ctypes as C
    import random
    from realtimedepthdiffusion_b200 import DepthDiffusion as D
    from realtimedepthdiffusion_b200._native import lib
    assert D.plan_passes(2160, 3840, 31) == ([7, 7, 7, 10], True)
    assert D.plan_passes(2160, 3840, 31, host_map=True) == ([7, 8, 16], True)
    assert D.plan_passes(1080, 1920, 62) == ([14, 16, 16, 16], True)
    assert D.plan_passes(4320, 7680, 15) == ([7, 8], True)
    assert D.plan_passes(4320, 7680, 15, host_map=True) == ([15], True)
    assert D.plan_passes(540, 960, 125)[1] == 0
    assert D.plan_passes(270, 480, 250) == ([8] + [11] * 22, 2)      # 4K level 3 / 1080p level 2: flat tiles, 11 per pass (round 1's measured choice)
    # contexts of a batch (several images in flight): least total SM time, not least latency -- shorter passes, smaller halos
    # (profiles/r02_tune_batch.txt: 0.706 against 0.765 ms per 1080p image)
    assert D.plan_passes(1080, 1920, 62, throughput=True) == ([8, 8, 8, 8, 8, 11, 11], True)
    assert D.plan_passes(540, 960, 125, throughput=True)[1] == 1
    rng = random.Random(5)
    for _ in range(200):
        r, c, it = rng.randint(1, 5000), rng.randint(1, 9000), rng.randint(1, 300)
        for host in (False, True):
            plan, _ = D.plan_passes(r, c, it, sm_count=rng.choice((1, 2, 74, 132, 148)), host_map=host, throughput=rng.random() < 0.3)
            assert sum(plan) == it and all(1 <= m <= 16 for m in plan), (r, c, it, plan)
            if host:
                assert plan[-1] in (min(16, it), min(15, it))        # 15: the flat form of a level below 2^18 pixels (128x32 regions)
            else:
                assert plan == sorted(plan)
    buf, f = (C.c_int * 4)(), C.c_int()
    assert lib.rtdd_plan_passes(2160, 3840, 31, 148, 0, buf, 2, C.byref(f)) < 0          # capacity too small
    assert lib.rtdd_plan_passes(2160, 3840, 0, 148, 0, buf, 4, C.byref(f)) == -1
    assert lib.rtdd_plan_passes(2160, 3840, 31, 148, 4, buf, 4, C.byref(f)) == -1


def test_strip_schedule_errors_are_negative():
    """rtdd.h: errors are negative (RTDD_E_ARG = -1), never confusable with a pass count."""
    import ctypes as C
    from realtimedepthdiffusion_b200._native import lib
    buf = (C.c_int * 4)()
    assert lib.rtdd_strip_schedule(-1, 8, 8, 0, buf, buf, 4) == -1            # negative sweep count
    assert lib.rtdd_strip_schedule(10, 0, 8, 0, buf, buf, 4) == -1            # halo < 1
    assert lib.rtdd_strip_schedule(10, 8, 8, -1, buf, buf, 4) == -1           # negative level
    assert lib.rtdd_strip_schedule(10, 8, 8, 0, None, buf, 4) == -1           # null output
    assert lib.rtdd_strip_schedule(64, 8, 8, 0, buf, buf, 4) == -1            # 8 passes do not fit a capacity of 4
    assert lib.rtdd_strip_schedule(32, 8, 8, 0, buf, buf, 4) == 4             # exactly fits
    assert lib.rtdd_strip_schedule(0, 8, 8, 0, buf, buf, 0) == 0              # no sweeps: no passes


@pytest.mark.parametrize("rows,cols,nranks,halo,iters", [(203, 150, 2, 4, 70), (256, 96, 3, 8, 100), (181, 130, 2, 5, 33)])
def test_lockstep_strips_are_bit_identical_to_the_single_solve(rows, cols, nranks, halo, iters):
    bgr, scribble, edited = synth.synth_case(rows, cols, 321)
    want_state = ob.FrameState(bgr)
    want = want_state.solve(scribble, edited, iters)
    engines = [CpuStripEngine(bgr, scribble, edited) for _ in range(nranks)]
    results, exchanges = strips.run_local(engines, iters, halo=halo, min_strip_pixels=1)
    assert exchanges > 0
    got = np.zeros_like(want)
    gotf = np.zeros_like(want_state.depth[0])
    for r, (plan, own) in enumerate(results):
        assert plan[0] is not None
        got[own[0]:own[1]] = engines[r].depth_u8[own[0]:own[1]]
        gotf[own[0]:own[1]] = engines[r].st.depth[0][own[0]:own[1]]
    assert np.array_equal(gotf.view(np.uint32), want_state.depth[0].view(np.uint32))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("rows,cols,nranks,halo,pass_sweeps,iters", [(203, 150, 2, 8, 4, 70), (256, 96, 3, 16, 8, 100), (181, 130, 2, 9, 4, 33)])
def test_several_passes_per_exchange_are_bit_identical(rows, cols, nranks, halo, pass_sweeps, iters):
    """`halo` ghost rows, passes of `pass_sweeps` sweeps: one exchange per `halo` sweeps instead of one per pass."""
    bgr, scribble, edited = synth.synth_case(rows, cols, 77)
    want_state = ob.FrameState(bgr)
    want_state.solve(scribble, edited, iters)
    engines = [CpuStripEngine(bgr, scribble, edited) for _ in range(nranks)]
    results, exchanges = strips.run_local(engines, iters, halo=halo, min_strip_pixels=1, pass_sweeps=pass_sweeps)
    engines1 = [CpuStripEngine(bgr, scribble, edited) for _ in range(nranks)]
    _, exchanges1 = strips.run_local(engines1, iters, halo=pass_sweeps, min_strip_pixels=1)
    assert 0 < exchanges < exchanges1
    gotf = np.zeros_like(want_state.depth[0])
    for r, (plan, own) in enumerate(results):
        gotf[own[0]:own[1]] = engines[r].st.depth[0][own[0]:own[1]]
    assert np.array_equal(gotf.view(np.uint32), want_state.depth[0].view(np.uint32))


@pytest.mark.parametrize("rows,cols,nranks,halo,sweeps", [(203, 150, 2, 4, 19), (256, 96, 4, 8, 64), (181, 130, 3, 5, 3)])
def test_level0_only_strips_match_the_single_level_solve(rows, cols, nranks, halo, sweeps):
    """configs[4] measurement (i): a fixed number of finest-level sweeps, strip-decomposed, from the same guess."""
    bgr, scribble, edited = synth.synth_case(rows, cols, 99)
    rng = np.random.default_rng(5)
    guess = rng.uniform(0, 255, (rows, cols)).astype(np.float32)
    solo = CpuStripEngine(bgr, scribble, edited)
    solo.st.depth[0] = guess.copy()
    (res, ex0) = strips.run_local([solo], 0, halo=halo, level0_sweeps=sweeps)
    assert ex0 == 0 and res[0][1] == (0, rows)
    forced = CpuStripEngine(bgr, scribble, edited)              # one rank through the strip entry points, whole image as window
    forced.st.depth[0] = guess.copy()
    strips.run_local([forced], 0, halo=halo, level0_sweeps=sweeps, force_strip_path=True)
    assert np.array_equal(forced.st.depth[0].view(np.uint32), solo.st.depth[0].view(np.uint32))
    engines = [CpuStripEngine(bgr, scribble, edited) for _ in range(nranks)]
    for e in engines:
        e.st.depth[0] = guess.copy()
    results, exchanges = strips.run_local(engines, 0, halo=halo, level0_sweeps=sweeps)
    assert exchanges == (sweeps - 1) // halo                  # one per pass except after the last
    got = np.zeros((rows, cols), np.float32)
    for r, (plan, own) in enumerate(results):
        got[own[0]:own[1]] = engines[r].st.depth[0][own[0]:own[1]]
    assert np.array_equal(got.view(np.uint32), solo.st.depth[0].view(np.uint32))


class _Recorder:
    """Engine stand-in that only records which entry points the per-rank frame calls, in order."""

    def __init__(self, sizes, fused, staged):
        self.sizes, self.fused_halo, self.staged_halo, self.log = sizes, fused, staged, []

    def __getattr__(self, name):
        if not (name.startswith("strip_") or name in ("annotation_pyramid", "convert_rows", "solve_full", "pyrup_rows", "quantise_rows")):
            raise AttributeError(name)

        def call(*args):
            self.log.append((name,) + args)
        return call


@pytest.mark.parametrize("mode", ["fused", "staged"])
@pytest.mark.parametrize("nranks,halo,pass_sweeps,iters", [(2, 8, None, 150), (4, 16, 8, 1000), (3, 4, None, 33)])
def test_peer_memory_modes_call_sequence(mode, nranks, halo, pass_sweeps, iters):
    """The GPU-only exchange modes on the CPU: which rtdd_strip_* calls a rank makes and in which order."""
    sizes = [(1081 >> l, 700 >> l) for l in range(3)]
    plan = strips.plan_strips(sizes, nranks, halo, 1)
    for rank in range(nranks):
        eng = _Recorder(sizes, True, mode == "staged")
        co = strips.frame_coroutine(eng, rank, nranks, iters, halo, 1, pass_sweeps=pass_sweeps)
        yields = 0
        try:
            ex = next(co)
            while True:
                assert ex.send_up is None and ex.send_dn is None          # no rows travel through the host side
                eng.log.append(("yield", ex.level))
                yields += 1
                ex = co.send(None)
        except StopIteration:
            pass
        names = [c[0] for c in eng.log]
        for l in range(3):
            assert plan[l] is not None
            it = strips.level_iterations(iters, 3, l)
            sched = strips.strip_schedule(it, halo, halo if mode == "fused" else pass_sweeps, l)
            calls = [c for c in eng.log if c[0].startswith("strip_") and c[1] == l]
            passes = [c for c in calls if c[0] == "strip_pass"]
            assert [c[3] for c in passes] == [n for n, _ in sched] and sum(c[3] for c in passes) == it
            if mode == "staged":
                pushes = [i for i, c in enumerate(eng.log) if c[:2] == ("strip_push", l)]
                assert len(pushes) == sum(1 for _, e in sched if e)
                for i in pushes:                                            # push, everybody yields, then pull
                    assert eng.log[i + 1] == ("yield", l) and eng.log[i + 2] == ("strip_pull", l)
                assert not any(c[0] in ("strip_wait", "strip_push_enable") for c in calls)
            else:
                waits = [c for c in calls if c[0] == "strip_wait"]
                assert len(waits) == (1 if l > 0 else 0)                    # ghost rows final before the prolongation reads them
                assert ("strip_push_enable", 0, False) in eng.log           # nobody reads level 0's ghost rows after its last pass
            assert [c[0] for c in calls][:2] == ["strip_init", "strip_neighbours"] and calls[-1][0] == "strip_finish"
        assert names[-1] == "quantise_rows"


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from oracle import binding as ob
from realtimedepthdiffusion_b200 import strips, synth
from tests.strip_cpu_engine import CpuStripEngine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ob.set_num_threads(2)
rows, cols, iters, halo = 203, 150, 70, 4
bgr, scribble, edited = synth.synth_case(rows, cols, 321)
eng = CpuStripEngine(bgr, scribble, edited)
plan, own, exchanges = strips.run_distributed(eng, dist, iters, halo=halo, min_strip_pixels=1)
# the same frame with twice the ghost rows and two passes per exchange, and the finest level alone: same rows, fewer exchanges
eng2 = CpuStripEngine(bgr, scribble, edited)
plan2, own2, exchanges2 = strips.run_distributed(eng2, dist, iters, halo=2 * halo, min_strip_pixels=1, pass_sweeps=halo)
same2 = own2 == own and exchanges2 < exchanges and np.array_equal(eng2.st.depth[0][own[0]:own[1]].view(np.uint32), eng.st.depth[0][own[0]:own[1]].view(np.uint32))
eng3 = CpuStripEngine(bgr, scribble, edited)
guess = np.random.default_rng(3).uniform(0, 255, (rows, cols)).astype(np.float32)
eng3.st.depth[0] = guess.copy()
plan3, own3, exchanges3 = strips.run_distributed(eng3, dist, 0, halo=halo, level0_sweeps=9)
solo = CpuStripEngine(bgr, scribble, edited)
solo.st.depth[0] = guess.copy()
strips.run_local([solo], 0, halo=halo, level0_sweeps=9)
same3 = exchanges3 == 2 and np.array_equal(eng3.st.depth[0][own3[0]:own3[1]].view(np.uint32), solo.st.depth[0][own3[0]:own3[1]].view(np.uint32))
flag = torch.tensor([1 if (same2 and same3) else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
full = torch.zeros((rows, cols), dtype=torch.float32)
full[own[0]:own[1]] = torch.from_numpy(eng.st.depth[0][own[0]:own[1]])
dist.all_reduce(full)                      # disjoint rows: the sum assembles the image
if rank == 0:
    st = ob.FrameState(bgr)
    st.solve(scribble, edited, iters)
    ok = np.array_equal(full.numpy().view(np.uint32), st.depth[0].view(np.uint32))
    print("STRIPS_OK" if ok and exchanges > 0 and int(flag.item()) == 1 else "STRIPS_MISMATCH", exchanges, exchanges2, exchanges3, int(flag.item()), flush=True)
dist.destroy_process_group()
'''


def test_world_size_2_gloo_strips_match_single_solve(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=280, env=env, cwd=ROOT)
    assert "STRIPS_OK" in r.stdout, r.stdout[-3000:]


def test_batch_sharding_is_a_partition():
    """configs[3]: image i -> rank i mod N, every image exactly once."""
    for n in (1, 2, 4, 8):
        seen = []
        for rank in range(n):
            seen += list(range(rank, 256, n))
        assert sorted(seen) == list(range(256))
