"""Row-strip domain decomposition of one very large image across GPUs (BASELINE configs[4]).

The reference is single-GPU; this module adds nothing to its arithmetic.  Fine pyramid levels are cut into contiguous
row strips, one per rank; every rank keeps H ghost rows beyond each strip boundary, runs up to H sweeps between two
halo exchanges (temporally blocked, so one exchange per H sweeps, not per sweep) and exchanges the H rows of x_k and
x_{k-1} next to each boundary with its two neighbours.  Coarse levels (too small to be worth splitting) are solved
redundantly by every rank -- bit-identical everywhere, so no broadcast is needed.  Results are bit-identical to the
single-GPU solve because every owned pixel still goes through exactly the reference's per-pixel recipe
(ref: src/GPUSolver.cu:73-106,226-262) once per sweep.

The frame logic is main.cpp's (ref: src/main.cpp:232-295), written as a per-rank coroutine that yields at every
halo exchange; `run_distributed` drives it with torch.distributed P2P (NCCL over NVLink on GPUs, gloo in the CPU
tests), `run_local` drives several ranks in lockstep inside one process (single-GPU emulation, unit tests).

The engine behind it is duck-typed: `GpuStripEngine` (this file) calls librtdd.so's rtdd_strip_* entry points;
the CPU tests plug in an oracle-backed engine to check the decomposition logic itself.
"""
import ctypes as C

import torch


def plan_strips(sizes, nranks, halo, min_strip_pixels=1 << 22):
    """sizes: [(rows, cols)] per level, finest first.  Returns plan[level] = None (replicated) or a list of
    (begin, end) owned row ranges per rank.  Strip boundaries double from one level to the next finer one so that
    a rank's fine rows are the prolongation of its own coarse rows (plus ghosts).  The planning itself is the native
    library's (rtdd_plan_strips, host only)."""
    from ._native import lib
    levels = len(sizes)
    plan = [None] * levels
    if nranks <= 1:
        return plan
    IntArr = C.c_int * levels
    rows = IntArr(*[int(r) for r, _ in sizes])
    cols = IntArr(*[int(c) for _, c in sizes])
    split = IntArr()
    begin = (C.c_int * (levels * nranks))()
    end = (C.c_int * (levels * nranks))()
    rc = lib.rtdd_plan_strips(rows, cols, levels, int(nranks), int(halo), int(min_strip_pixels), split, begin, end)
    assert rc == 0, "strip shorter than the halo"
    for l in range(levels):
        if split[l]:
            plan[l] = [(begin[l * nranks + r], end[l * nranks + r]) for r in range(nranks)]
    return plan


def strip_schedule(iters, halo, pass_sweeps, level):
    """[(sweeps of the pass, exchange after it?)] for one split level (rtdd_strip_schedule, host only)."""
    from ._native import lib
    cap = max(int(iters), 1)
    sweeps = (C.c_int * cap)()
    exch = (C.c_int * cap)()
    n = lib.rtdd_strip_schedule(int(iters), int(halo), int(pass_sweeps or 0), int(level), sweeps, exch, cap)
    assert n >= 0, "rtdd_strip_schedule"
    return [(sweeps[i], bool(exch[i])) for i in range(n)]


def level_iterations(max_iterations, levels, level):
    """ref: src/main.cpp:263 (float division, truncation)"""
    import numpy as np
    return int(np.float32(max_iterations) / np.float32(2.0 ** ((levels - 1) - level)))


class Exchange:
    """What a rank hands to the driver at a yield: rows to send up/down and views to receive into."""

    def __init__(self, level, send_up, recv_up, send_dn, recv_dn):
        self.level, self.send_up, self.recv_up, self.send_dn, self.recv_dn = level, send_up, recv_up, send_dn, recv_dn


def _solve_split_level(engine, l, rank, nranks, strips_l, iters, halo, pass_sweeps=None):
    """`iters` sweeps of a level that is cut into row strips; yields an Exchange per halo exchange.
    `halo` ghost rows per open side allow `halo` sweeps between two exchanges; they run as passes of at most
    `pass_sweeps` sweeps (default: one pass of `halo` sweeps), so e.g. halo 16 with pass_sweeps 8 halves the number of
    exchanges while every pass keeps the kernel's best temporal block size."""
    rows = engine.sizes[l][0]
    a, b = strips_l[rank]
    w0, w1 = max(0, a - halo), min(rows, b + halo)
    engine.strip_init(l, w0, w1)
    staged = bool(getattr(engine, "staged_halo", False))      # peer memory through staging rows (rtdd_strip_push / _pull)
    fused = bool(getattr(engine, "fused_halo", False)) and not staged
    T = halo if (pass_sweeps is None or fused) else max(1, min(int(pass_sweeps), halo))   # the fused push is per pass
    if fused or staged:
        # the sweep passes push their boundary rows into the neighbours' ghost rows themselves (peer memory)
        up0 = max(0, strips_l[rank - 1][0] - halo) if rank > 0 else -1
        dn0 = max(0, strips_l[rank + 1][0] - halo) if rank < nranks - 1 else -1
        engine.strip_neighbours(l, a, b, halo, up0, dn0)
    k = 0
    for n, exchange in strip_schedule(iters, halo, T, l):
        if fused and l == 0 and k + n >= iters:
            engine.strip_push_enable(l, False)            # the finest level's last pass: nobody reads the ghost rows afterwards
        engine.strip_pass(l, k, n, T)
        k += n
        if fused:                                         # (T == halo: the schedule has an exchange after every pass but level 0's last)
            if k >= iters and l > 0:
                engine.strip_wait(l)                      # ghost rows must be final before the prolongation reads them
            yield Exchange(l, None, None, None, None)     # no data: only keeps single-process emulations in lockstep
        elif exchange:                                    # the last exchange of a level > 0 feeds the prolongation
            if staged:
                engine.strip_push(l)                      # boundary rows -> the neighbours' staging rows + flags
                yield Exchange(l, None, None, None, None) # (single-process emulations: everybody pushes before anybody pulls)
                engine.strip_pull(l)                      # wait for the neighbours' flags, staging rows -> my ghost rows
                continue
            xk, xkm1 = engine.strip_planes(l)
            gt, gb = a - w0, w1 - b                       # ghost rows above / below
            own0, own1 = gt, gt + (b - a)
            yield Exchange(l,
                           [xk[own0:own0 + halo], xkm1[own0:own0 + halo]] if gt else None,
                           [xk[0:gt], xkm1[0:gt]] if gt else None,
                           [xk[own1 - halo:own1], xkm1[own1 - halo:own1]] if gb else None,
                           [xk[own1:own1 + gb], xkm1[own1:own1 + gb]] if gb else None)
    if l > 0:
        engine.strip_finish(l, w0, w1)                    # owned rows + freshly exchanged ghosts
    else:
        engine.strip_finish(l, a, b)                      # finest level: ghosts are stale and not needed


def level0_coroutine(engine, rank, nranks, sweeps, halo=8, pass_sweeps=None, force_strip_path=False):
    """BASELINE configs[4], measurement (i) of SURVEY.md section 8d: only the finest level, a fixed number of sweeps from
    whatever guess engine.depth[0] holds, cut into `nranks` equal row strips (one rank: the ordinary level solve).
    Same per-pixel recipe and the same halo logic as the frame, so owned rows are bit-identical to one GPU."""
    rows = engine.sizes[0][0]
    mark = getattr(engine, "mark", lambda label: None)
    mark("begin")
    if nranks <= 1 and not force_strip_path:          # (force_strip_path: one rank through rtdd_strip_* with the whole image as its window)
        engine.solve_full(0, sweeps)
        mark("solve L0")
        return [None] * len(engine.sizes), (0, rows)
    bounds = [(r * rows) // nranks for r in range(nranks)] + [rows]
    strips0 = [(bounds[r], bounds[r + 1]) for r in range(nranks)]
    assert all(e - b >= halo for b, e in strips0), "strip shorter than the halo"
    yield from _solve_split_level(engine, 0, rank, nranks, strips0, sweeps, halo, pass_sweeps)
    mark("solve L0")
    return [strips0] + [None] * (len(engine.sizes) - 1), strips0[rank]


def frame_coroutine(engine, rank, nranks, max_iterations, halo=8, min_strip_pixels=1 << 22, gather_result=False, pass_sweeps=None):
    """One solve frame on one rank; yields an Exchange whenever halo rows must move.
    The engine's level planes (depth/scribble/edited/gray) are full-size on every rank; only this rank's window of a
    split level holds meaningful depth values."""
    sizes = engine.sizes
    L = len(sizes)
    plan = plan_strips(sizes, nranks, halo, min_strip_pixels)
    mark = getattr(engine, "mark", lambda label: None)           # optional per-phase device timestamps
    mark("begin")
    engine.annotation_pyramid()                                   # main.cpp:249 (replicated: u8 planes, cheap)
    engine.convert_rows(L - 1, 0, sizes[L - 1][0])                # main.cpp:257
    mark("annotation")
    for l in range(L - 1, -1, -1):
        iters = level_iterations(max_iterations, L, l)
        rows = sizes[l][0]
        if plan[l] is None:
            engine.solve_full(l, iters)                           # main.cpp:266
        else:
            yield from _solve_split_level(engine, l, rank, nranks, plan[l], iters, halo, pass_sweeps)
        mark("solve L%d" % l)
        if l > 0:
            nrows = sizes[l - 1][0]
            if plan[l - 1] is None:
                n0, n1 = 0, nrows
            else:
                a, b = plan[l - 1][rank]
                n0, n1 = max(0, a - halo), min(nrows, b + halo)
            engine.pyrup_rows(l, n0, n1)                          # main.cpp:272-279
            engine.convert_rows(l - 1, n0, n1)                    # main.cpp:281
            mark("prolong L%d->L%d" % (l, l - 1))
    own = (0, sizes[0][0]) if plan[0] is None else plan[0][rank]
    engine.quantise_rows(own[0], own[1])                          # main.cpp:290
    mark("quantise")
    return plan, own


def run_distributed(engine, dist, max_iterations, halo=8, min_strip_pixels=1 << 22, level0_sweeps=0, pass_sweeps=None):
    """Drive this process's rank; halo rows travel with batched isend/irecv (NCCL on GPUs, gloo on CPUs).
    level0_sweeps > 0 runs level0_coroutine (finest level only) instead of the frame."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if level0_sweeps > 0:
        co = level0_coroutine(engine, rank, world, level0_sweeps, halo, pass_sweeps)
    else:
        co = frame_coroutine(engine, rank, world, max_iterations, halo, min_strip_pixels, pass_sweeps=pass_sweeps)
    exchanges = 0
    try:
        ex = next(co)
        while True:
            ops = []
            if ex.send_up is not None:
                for t in ex.send_up:
                    ops.append(dist.P2POp(dist.isend, t, rank - 1))
                for t in ex.recv_up:
                    ops.append(dist.P2POp(dist.irecv, t, rank - 1))
            if ex.send_dn is not None:
                for t in ex.send_dn:
                    ops.append(dist.P2POp(dist.isend, t, rank + 1))
                for t in ex.recv_dn:
                    ops.append(dist.P2POp(dist.irecv, t, rank + 1))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            exchanges += 1
            ex = co.send(None)
    except StopIteration as done:
        plan, own = done.value
    return plan, own, exchanges


def enable_fused_halo_local(engines):
    """Single-process emulation: the neighbours' arenas are ordinary device pointers of the same process."""
    bases = [e.arena()[0] for e in engines]
    for r, e in enumerate(engines):
        e.set_peers(bases[r - 1] if r > 0 else None, bases[r + 1] if r + 1 < len(engines) else None)


def run_local(engines, max_iterations, halo=8, min_strip_pixels=1 << 22, level0_sweeps=0, pass_sweeps=None, force_strip_path=False):
    """All ranks in one process, in lockstep (emulation on one device / CPU unit tests)."""
    n = len(engines)
    if level0_sweeps > 0:
        cos = [level0_coroutine(e, r, n, level0_sweeps, halo, pass_sweeps, force_strip_path) for r, e in enumerate(engines)]
    else:
        cos = [frame_coroutine(e, r, n, max_iterations, halo, min_strip_pixels, pass_sweeps=pass_sweeps) for r, e in enumerate(engines)]
    results = [None] * n
    pending = [None] * n
    live = set(range(n))
    for r in range(n):
        try:
            pending[r] = next(cos[r])
        except StopIteration as done:
            results[r] = done.value
            live.discard(r)
    exchanges = 0
    while live:
        assert live == set(range(n)), "ranks must reach every exchange together"
        for r in range(n):                 # everybody's sends are snapshots of rows nobody else writes
            ex = pending[r]
            if ex.send_up is not None:
                for src, dst in zip(ex.send_up, pending[r - 1].recv_dn):
                    dst.copy_(src)
            if ex.send_dn is not None:
                for src, dst in zip(ex.send_dn, pending[r + 1].recv_up):
                    dst.copy_(src)
        exchanges += 1
        for r in range(n):
            try:
                pending[r] = cos[r].send(None)
            except StopIteration as done:
                results[r] = done.value
                live.discard(r)
    return results, exchanges


class GpuStripEngine:
    """Per-rank state for the strip solve on one GPU: full-size level planes (what main.cpp keeps in its GpuMat
    vectors, ref: src/main.cpp:116-147) plus one DepthDiffusion context; every computation goes through librtdd.so."""

    def __init__(self, ctx, bgr_dev, scribble_dev, edited_dev):
        from .api import pitched_empty
        self.ctx = ctx
        self.dev = ctx.device
        self.sizes = ctx.sizes
        self.levels = ctx.levels
        rows, cols = self.sizes[0]
        self.bgr = bgr_dev
        self.gray = [pitched_empty(rows, cols, torch.uint8, self.dev)]
        ctx.bgr2gray(self.bgr, self.gray[0])
        for l in range(1, self.levels):
            g = self.gray[l - 1]
            self.gray.append(pitched_empty((g.shape[0] + 1) // 2, (g.shape[1] + 1) // 2, torch.uint8, self.dev))
            ctx.pyrdown_gray(g, self.gray[l])
        self.scribble = [scribble_dev] + [pitched_empty(r, c, torch.uint8, self.dev, fill=0) for r, c in self.sizes[1:]]
        self.edited = [edited_dev] + [pitched_empty(r, c, torch.uint8, self.dev, channels=3, fill=0) for r, c in self.sizes[1:]]
        self.depth = [pitched_empty(r, c, torch.float32, self.dev, fill=255.0) for r, c in self.sizes]
        self.depth_u8 = pitched_empty(rows, cols, torch.uint8, self.dev, fill=0)
        self._views = {}
        self.fused_halo = False      # set by set_peers / enable_fused_halo_distributed
        self.staged_halo = False     # set by enable_staged_halo
        self.marks = None            # set to [] to collect (label, event) pairs for one frame

    # -- helpers ------------------------------------------------------------------
    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    @staticmethod
    def _pitch(t):
        return t.stride(0) * t.element_size()

    def _ck(self, rc):
        self.ctx._ck(rc)

    def mark(self, label):
        if self.marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.dev))
            self.marks.append((label, ev))

    def phase_ms(self):
        torch.cuda.synchronize(self.dev)
        return [(self.marks[i][0], self.marks[i - 1][1].elapsed_time(self.marks[i][1])) for i in range(1, len(self.marks))]

    # -- engine interface -----------------------------------------------------------
    def annotation_pyramid(self):
        for l in range(1, self.levels):
            self.ctx.pyrdown_annotation(self.scribble[l - 1], self.edited[l - 1], self.scribble[l], self.edited[l])

    def convert_rows(self, l, r0, r1):
        if r1 > r0:
            self.ctx.convert_to_float(self.edited[l][r0:r1], self.depth[l][r0:r1], self.scribble[l][r0:r1])

    def solve_full(self, l, iters):
        r, c = self.sizes[l]
        self.ctx.matrix_free_solver(self.depth[l], self.scribble[l], self.gray[l][:r, :c], iters, l)

    def pyrup_rows(self, l, r0, r1):
        from ._native import lib
        s, d = self.depth[l], self.depth[l - 1]
        self._ck(lib.rtdd_pyrup_depth_rows(self.ctx._h, self._p(s), self._pitch(s), s.shape[0], s.shape[1],
                                           self._p(d), self._pitch(d), d.shape[0], d.shape[1], int(r0), int(r1)))

    def strip_init(self, l, w0, w1):
        from ._native import lib
        r, c = self.sizes[l]
        d, s, g = self.depth[l], self.scribble[l], self.gray[l]
        self._ck(lib.rtdd_strip_init(self.ctx._h, l, self._p(d), self._pitch(d), self._p(s), self._pitch(s), self._p(g), self._pitch(g),
                                     r, c, int(w0), int(w1)))

    def strip_pass(self, l, k0, n, halo):
        from ._native import lib
        self._ck(lib.rtdd_strip_pass(self.ctx._h, l, int(k0), int(n), int(halo)))

    def strip_planes(self, l):
        """(x_k, x_{k-1}) as torch views [window rows, pitch floats] over the library's planes (no copy)."""
        from ._native import lib
        xk, xm, pitch, wb, wr = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self._ck(lib.rtdd_strip_planes(self.ctx._h, l, C.byref(xk), C.byref(xm), C.byref(pitch), C.byref(wb), C.byref(wr)))
        out = []
        for p in (xk.value, xm.value):
            key = (p, wr.value, pitch.value)
            if key not in self._views:
                class _H:
                    pass
                h = _H()
                h.__cuda_array_interface__ = {"shape": (wr.value, pitch.value // 4), "typestr": "<f4", "data": (p, False), "version": 2}
                self._views[key] = torch.as_tensor(h, device=self.dev)
            out.append(self._views[key])
        return out

    # -- fused halo push (peer memory) -----------------------------------------------------------
    def arena(self):
        from ._native import lib
        base, nbytes = C.c_void_p(), C.c_size_t()
        self._ck(lib.rtdd_arena(self.ctx._h, C.byref(base), C.byref(nbytes)))
        return base.value, nbytes.value

    def set_peers(self, above, below):
        from ._native import lib
        self._ck(lib.rtdd_strip_set_peers(self.ctx._h, C.c_void_p(above or 0), C.c_void_p(below or 0)))
        self.fused_halo = bool(above or below)

    def enable_fused_halo_distributed(self, dist):
        """Exchange CUDA IPC handles of the arenas with the two neighbouring ranks and map them."""
        from ._native import lib
        rank, world = dist.get_rank(), dist.get_world_size()
        h = C.create_string_buffer(64)
        self._ck(lib.rtdd_ipc_export(self.ctx._h, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw))
        ptrs = {}
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < world:
                p = C.c_void_p()
                self._ck(lib.rtdd_ipc_import(self.ctx._h, C.create_string_buffer(handles[nb], 64), C.byref(p)))
                ptrs[nb] = p.value
        self.set_peers(ptrs.get(rank - 1), ptrs.get(rank + 1))

    def strip_neighbours(self, l, a, b, halo, up0, dn0):
        from ._native import lib
        self._ck(lib.rtdd_strip_neighbours(self.ctx._h, l, int(a), int(b), int(halo), int(up0), int(dn0)))

    def strip_push(self, l):
        from ._native import lib
        self._ck(lib.rtdd_strip_push(self.ctx._h, l))

    def strip_pull(self, l):
        from ._native import lib
        self._ck(lib.rtdd_strip_pull(self.ctx._h, l))

    def enable_staged_halo(self):
        """After set_peers / enable_fused_halo_*: keep the plain sweep kernels and move halo rows with push / pull kernels."""
        self.ctx.set_tuning("strip_peer_staging", 1)
        self.staged_halo = True

    def strip_wait(self, l):
        from ._native import lib
        self._ck(lib.rtdd_strip_wait(self.ctx._h, l))

    def strip_push_enable(self, l, on):
        from ._native import lib
        self._ck(lib.rtdd_strip_push_enable(self.ctx._h, l, 1 if on else 0))

    def strip_finish(self, l, r0, r1):
        from ._native import lib
        d = self.depth[l]
        self._ck(lib.rtdd_strip_finish(self.ctx._h, l, self._p(d), self._pitch(d), int(r0), int(r1)))

    def quantise_rows(self, r0, r1):
        if r1 > r0:
            self.ctx.quantise_u8(self.depth[0][r0:r1], self.depth_u8[r0:r1])
