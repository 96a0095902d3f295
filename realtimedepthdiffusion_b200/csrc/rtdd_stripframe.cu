// Row-strip frames in C++: one very large image solved by N GPUs (BASELINE configs[4]) and the batch of independent images
// (configs[3]), behind the C ABI.
//
//   rtdd_strip_frame_*   one RANK's share of a frame, on that rank's context.  The frame logic is main.cpp's
//                        (ref: src/main.cpp:232-295): annotation restriction, Dirichlet injection, per level edge-weight pass +
//                        sweeps, prolongation, 8-bit map -- with the fine levels cut into row strips (rtdd_plan_strips), H ghost
//                        rows per open side, several temporally blocked passes between two halo exchanges
//                        (rtdd_strip_schedule) and the halo rows moved by the staged peer exchange (rtdd_strip_push / _pull:
//                        peer-memory stores over NVLink + system-scope flags, no NCCL, no host round trip).  The ranks may be
//                        threads of one process (rtdd_mgpu_*) or processes (peers mapped through rtdd_ipc_*).
//   rtdd_mgpu_*          one process, one host thread per GPU: creates the contexts, enables peer access, wires the
//                        neighbours and drives all ranks.  This is what a C++ host like main.cpp links against.
//
// The reference is single-GPU; nothing here changes its arithmetic: owned rows are bit-identical to one GPU (tests).

#include "rtdd_internal.h"

#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct DeviceGuard2 {
    int prev = -1;
    explicit DeviceGuard2(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard2() { if (prev >= 0) cudaSetDevice(prev); }
};

#define SF_TRY(expr)                 \
    do {                             \
        const int _rc = (expr);      \
        if (_rc) return _rc;         \
    } while (0)

int window_of(const rtdd_ctx *ctx, int level, int *a, int *b, int *w0, int *w1)
{
    const RtddStripFrame &sf = ctx->sf;
    const int rows = ctx->lv[level].rows;
    if (sf.nranks <= 1 || !sf.split[level]) { *a = 0; *b = rows; *w0 = 0; *w1 = rows; return 0; }
    *a = sf.ownBegin[level * sf.nranks + sf.rank];
    *b = sf.ownEnd[level * sf.nranks + sf.rank];
    *w0 = *a - sf.halo > 0 ? *a - sf.halo : 0;
    *w1 = *b + sf.halo < rows ? *b + sf.halo : rows;
    return 1;
}

// One split level on this rank: edge-weight pass on the window, passes + exchanges, result into the frame's depth plane.
int solve_split_level(rtdd_ctx *ctx, int l, int iters, bool lastLevelOfRun)
{
    const RtddStripFrame &sf = ctx->sf;
    RtddFrameLevel &F = ctx->fl[l];
    int a, b, w0, w1;
    window_of(ctx, l, &a, &b, &w0, &w1);
    SF_TRY(rtdd_strip_init(ctx, l, F.depth, F.depthPitch, F.scribble, F.scribblePitch, F.gray, F.grayPitch, F.rows, F.cols, w0, w1));
    const int n = sf.nranks, r = sf.rank;
    const int up0 = r > 0 ? (sf.ownBegin[l * n + r - 1] - sf.halo > 0 ? sf.ownBegin[l * n + r - 1] - sf.halo : 0) : -1;
    const int dn0 = r < n - 1 ? (sf.ownBegin[l * n + r + 1] - sf.halo > 0 ? sf.ownBegin[l * n + r + 1] - sf.halo : 0) : -1;
    SF_TRY(rtdd_strip_neighbours(ctx, l, a, b, sf.halo, up0, dn0));
    int T = sf.passSweeps;
    if (T < 1 || T > sf.halo) T = sf.halo;
    if (T > RTDD_MAX_T) T = RTDD_MAX_T;
    std::vector<int> sweeps(iters > 0 ? iters : 1), exch(iters > 0 ? iters : 1);
    // the last exchange of a level feeds the prolongation of the next finer one; the finest level of a run needs none
    const int npass = rtdd_strip_schedule(iters, sf.halo, T, lastLevelOfRun ? 0 : 1, sweeps.data(), exch.data(), (int)sweeps.size());
    if (npass < 0) return rtdd_fail(ctx, npass, "rtdd_strip_frame (schedule)");
    int k = 0;
    bool direct = false;
    for (int p = 0; p < npass; p++) {
        const bool last = (p == npass - 1);
        if (last && lastLevelOfRun && ((((uintptr_t)F.depth | F.depthPitch) & 15u) == 0)) {
            // nobody reads this level's ghost rows afterwards: the last pass writes the frame's depth plane (and, on level 0, the 8-bit map)
            SF_TRY(rtdd_strip_pass_to(ctx, l, k, sweeps[p], T, F.depth, F.depthPitch, l == 0 ? ctx->depthU8 : nullptr, l == 0 ? ctx->depthU8Pitch : 0));
            direct = true;
        } else {
            SF_TRY(rtdd_strip_pass(ctx, l, k, sweeps[p], T));
        }
        k += sweeps[p];
        if (exch[p]) {
            SF_TRY(rtdd_strip_push(ctx, l));
            SF_TRY(rtdd_strip_pull(ctx, l));
        }
    }
    if (!direct) {
        if (lastLevelOfRun) SF_TRY(rtdd_strip_finish(ctx, l, F.depth, F.depthPitch, a, b));       // ghost rows are stale and not needed
        else SF_TRY(rtdd_strip_finish(ctx, l, F.depth, F.depthPitch, w0, w1));                     // owned rows + freshly exchanged ghosts
        if (l == 0 && lastLevelOfRun) {
            SF_TRY(rtdd_quantise_u8(ctx, (const float *)((const char *)F.depth + (size_t)a * F.depthPitch), F.depthPitch,
                                    ctx->depthU8 + (size_t)a * ctx->depthU8Pitch, ctx->depthU8Pitch, b - a, F.cols));
        }
    }
    return 0;
}

int strip_frame_run(rtdd_ctx *ctx, int maxIterations, int level0Sweeps)
{
    if (!ctx) return RTDD_E_ARG;
    if (!ctx->sf.ready) return rtdd_fail(ctx, RTDD_E_STATE, "rtdd_strip_frame (rtdd_strip_frame_setup not called)");
    if (!ctx->imageSet || !ctx->lutLoaded) return rtdd_fail(ctx, RTDD_E_STATE, "rtdd_strip_frame (image / weights not set)");
    if (maxIterations < 0 || level0Sweeps < 0) return rtdd_fail(ctx, RTDD_E_ARG, "rtdd_strip_frame");
    DeviceGuard2 guard(ctx->device);
    const RtddStripFrame &sf = ctx->sf;
    const int L = ctx->levels;
    cudaStream_t s = ctx->stream;
    if (level0Sweeps > 0) {
        // only the finest level, a fixed sweep count, from whatever guess the depth plane holds
        if (sf.nranks <= 1 || !sf.split[0]) {
            RtddFrameLevel &F = ctx->fl[0];
            return rtdd_solve_level(ctx, F.depth, F.depthPitch, F.scribble, F.scribblePitch, F.gray, F.grayPitch, F.rows, F.cols, level0Sweeps, 0);
        }
        return solve_split_level(ctx, 0, level0Sweeps, true);
    }
    for (int l = 1; l < L; l++) {                                                            // main.cpp:249 (replicated: u8 planes)
        RtddFrameLevel &P = ctx->fl[l - 1], &F = ctx->fl[l];
        SF_TRY(rtdd_pyrdown_annotation(ctx, P.scribble, P.scribblePitch, P.edited, P.editedPitch, P.rows, P.cols,
                                       F.scribble, F.scribblePitch, F.edited, F.editedPitch, F.rows, F.cols));
    }
    {
        RtddFrameLevel &F = ctx->fl[L - 1];                                                  // main.cpp:257
        SF_TRY(rtdd_convert_to_float(ctx, F.edited, F.editedPitch, F.depth, F.depthPitch, F.scribble, F.scribblePitch, F.rows, F.cols));
    }
    for (int l = L - 1; l >= 0; l--) {                                                       // main.cpp:261-288
        RtddFrameLevel &F = ctx->fl[l];
        const int iters = rtdd_level_iterations(maxIterations, L, l);
        if (sf.nranks <= 1 || !sf.split[l]) {
            SF_TRY(rtdd_solve_level(ctx, F.depth, F.depthPitch, F.scribble, F.scribblePitch, F.gray, F.grayPitch, F.rows, F.cols, iters, l));
            if (l == 0)
                SF_TRY(rtdd_quantise_u8(ctx, F.depth, F.depthPitch, ctx->depthU8, ctx->depthU8Pitch, F.rows, F.cols));
        } else {
            SF_TRY(solve_split_level(ctx, l, iters, l == 0));
        }
        if (l > 0) {
            RtddFrameLevel &N = ctx->fl[l - 1];
            int a, b, n0, n1;
            window_of(ctx, l - 1, &a, &b, &n0, &n1);
            // main.cpp:272-281 on the rows this rank needs of the next finer level
            SF_TRY(rtdd_pyrup_depth_rows(ctx, F.depth, F.depthPitch, F.rows, F.cols, N.depth, N.depthPitch, N.rows, N.cols, n0, n1));
            SF_TRY(rtdd_convert_to_float(ctx, N.edited + (size_t)n0 * N.editedPitch, N.editedPitch,
                                         (float *)((char *)N.depth + (size_t)n0 * N.depthPitch), N.depthPitch,
                                         N.scribble + (size_t)n0 * N.scribblePitch, N.scribblePitch, n1 - n0, N.cols));
        }
    }
    (void)s;
    return 0;
}

}  // namespace

extern "C" {

int rtdd_strip_frame_setup(rtdd_ctx *ctx, int rank, int nranks, int halo, int passSweeps, long long minStripPixels)
{
    if (!ctx) return RTDD_E_ARG;
    if (rank < 0 || nranks < 1 || rank >= nranks || halo < 1 || halo > RTDD_MAX_HALO || passSweeps < 0 || minStripPixels < 1)
        return rtdd_fail(ctx, RTDD_E_ARG, "rtdd_strip_frame_setup");
    RtddStripFrame &sf = ctx->sf;
    const int L = ctx->levels;
    std::vector<int> rows(L), cols(L);
    for (int l = 0; l < L; l++) { rows[l] = ctx->lv[l].rows; cols[l] = ctx->lv[l].cols; }
    sf.split.assign(L, 0);
    sf.ownBegin.assign((size_t)L * nranks, 0);
    sf.ownEnd.assign((size_t)L * nranks, 0);
    const int rc = rtdd_plan_strips(rows.data(), cols.data(), L, nranks, halo, minStripPixels, sf.split.data(), sf.ownBegin.data(), sf.ownEnd.data());
    if (rc) return rtdd_fail(ctx, rc, "rtdd_strip_frame_setup (a strip would be shorter than its halo)");
    sf.rank = rank; sf.nranks = nranks; sf.halo = halo; sf.passSweeps = passSweeps;
    sf.ready = true;
    if (nranks > 1) ctx->peerStaging = true;          // halo rows travel through rtdd_strip_push / rtdd_strip_pull
    {
        DeviceGuard2 guard(ctx->device);
        const int e = rtdd_check(ctx, cudaMemsetAsync(ctx->dPeerBadWord, 0, sizeof(unsigned int), ctx->stream), "rtdd_strip_frame_setup");
        if (e) return e;
    }
    return 0;
}

int rtdd_strip_frame_solve(rtdd_ctx *ctx, int maxIterations) { return strip_frame_run(ctx, maxIterations, 0); }

int rtdd_strip_frame_level0(rtdd_ctx *ctx, int sweeps)
{
    if (sweeps < 1) return rtdd_fail(ctx, RTDD_E_ARG, "rtdd_strip_frame_level0");
    return strip_frame_run(ctx, 0, sweeps);
}

int rtdd_strip_frame_rows(rtdd_ctx *ctx, int level, int *split, int *ownBegin, int *ownEnd, int *winBegin, int *winEnd)
{
    if (!ctx) return RTDD_E_ARG;
    if (!ctx->sf.ready || level < 0 || level >= ctx->levels) return rtdd_fail(ctx, RTDD_E_ARG, "rtdd_strip_frame_rows");
    int a, b, w0, w1;
    const int sp = window_of(ctx, level, &a, &b, &w0, &w1);
    if (split) *split = sp;
    if (ownBegin) *ownBegin = a;
    if (ownEnd) *ownEnd = b;
    if (winBegin) *winBegin = w0;
    if (winEnd) *winEnd = w1;
    return 0;
}

// ref: src/main.cpp:190-230 on this rank's rows of the finest level (DepthEffect row strips); outputs are FULL
// caller planes of which only the owned rows are written
int rtdd_strip_frame_effects(rtdd_ctx *ctx, uint8_t *desat, size_t desatPitch, uint8_t *haze, size_t hazePitch, uint8_t *defocus, size_t defocusPitch)
{
    if (!ctx) return RTDD_E_ARG;
    if (!ctx->sf.ready || !ctx->imageSet) return rtdd_fail(ctx, RTDD_E_STATE, "rtdd_strip_frame_effects");
    int a, b, w0, w1;
    window_of(ctx, 0, &a, &b, &w0, &w1);
    RtddFrameLevel &F = ctx->fl[0];
    return rtdd_effects_rows(ctx, ctx->bgr, ctx->bgrPitch, F.gray, F.grayPitch, F.depth, F.depthPitch, desat, desatPitch, haze, hazePitch,
                             defocus, defocusPitch, ctx->rows, ctx->cols, a, b);
}

// ---- one process, one host thread per GPU ----------------------------------------------------------------------------

struct rtdd_mgpu {
    int n = 0, rows = 0, cols = 0, levels = 0;
    std::vector<int> devices;
    std::vector<rtdd_ctx *> ctx;
    std::vector<std::vector<rtdd_ctx *>> batchCtx;       // per GPU: contexts for independent images, several in flight (made on first use)
    float beta = 0.4f;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cvWork, cvDone;
    unsigned long long generation = 0;        // bumped per command
    int pending = 0;
    bool quit = false;
    // the current command (read by every worker)
    int cmd = 0;                               // 1 = strip frame, 2 = strip level 0, 3 = batch, 4 = upload image, 5 = upload annotation + solve + download
    int iArg = 0;
    const uint8_t *hostA = nullptr; size_t pitchA = 0;       // image / annotation
    uint8_t *hostOut = nullptr; size_t pitchOut = 0;
    int nimages = 0;
    const uint8_t *const *bgrList = nullptr; const uint8_t *const *annList = nullptr; uint8_t *const *outList = nullptr;
    size_t bgrPitch = 0, annPitch = 0, outPitch = 0;
    std::vector<int> status;
    std::vector<float> ms;
    std::string err;
};

namespace {

void mgpu_worker(rtdd_mgpu *m, int r)
{
    cudaSetDevice(m->devices[r]);
    rtdd_ctx *ctx = m->ctx[r];
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    unsigned long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cvWork.wait(lk, [&] { return m->quit || m->generation != seen; });
            if (m->quit) break;
            seen = m->generation;
        }
        int rc = 0;
        float ms = 0.0f;
        cudaEventRecord(e0, ctx->stream);
        switch (m->cmd) {
        case 1: rc = rtdd_strip_frame_solve(ctx, m->iArg); break;
        case 2: rc = rtdd_strip_frame_level0(ctx, m->iArg); break;
        case 3: {
            // BASELINE configs[3]: image i -> GPU i mod N, every image a full job from host buffers.  Up to 8 images are in flight
            // per GPU (one context and stream each): the coarse levels keep <= 16 SMs busy, other images' levels fill the rest
            // (measured at 1080p, tools/tune_batch.py: 0.89 / 0.79 / 0.70 / 0.66 / 0.67 ms per image with 3 / 4 / 6 / 8 / 12 in flight).
            const int K = 8;
            std::vector<rtdd_ctx *> &bc = m->batchCtx[r];
            while ((int)bc.size() < K && !rc) {
                rtdd_ctx *c = nullptr;
                rc = rtdd_create(m->rows, m->cols, m->levels, m->devices[r], &c);
                if (!rc) rc = rtdd_load_weights(c, m->beta);
                if (!rc) rc = rtdd_set_tuning(c, "plan_throughput", 1);      // several images in flight: SM time counts, not one image's latency
                if (c) bc.push_back(c);
            }
            int j = 0;
            for (int i = r; i < m->nimages && !rc; i += m->n, j++) {
                rtdd_ctx *c = bc[j % K];
                rc = rtdd_frame_set_image(c, m->bgrList[i], m->bgrPitch);
                if (!rc) rc = rtdd_frame_solve_host_annotation(c, m->annList[i], m->annPitch, m->iArg, nullptr, 0);
                if (!rc && m->outList && m->outList[i]) rc = rtdd_frame_read_depth_u8(c, m->outList[i], m->outPitch, 0);
            }
            for (rtdd_ctx *c : bc) { const int rs = rtdd_sync(c); if (!rc) rc = rs; }
            if (rc && !bc.empty()) ctx->err = rtdd_last_error(bc[0]);
            break;
        }
        case 4: rc = rtdd_frame_set_image(ctx, m->hostA, m->pitchA); break;
        case 5: {
            // every rank ingests the whole annotation plane (the coarse, replicated levels need all of it), solves its strips
            // and returns its own rows of the 8-bit map
            RtddFrameLevel &F = ctx->fl[0];
            cudaError_t e = cudaMemcpy2DAsync(ctx->annot, ctx->annotPitch, m->hostA, m->pitchA, (size_t)ctx->cols, ctx->rows, cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = rtdd::launch_annotation_ingest(ctx->stream, ctx->annot, ctx->annotPitch, ctx->bgr, ctx->bgrPitch, F.edited, F.editedPitch,
                                                                     F.scribble, F.scribblePitch, ctx->rows, ctx->cols);
            rc = rtdd_check(ctx, e, "rtdd_mgpu_frame_solve_host_annotation");
            if (!rc) rc = rtdd_strip_frame_solve(ctx, m->iArg);
            if (!rc && m->hostOut) {
                int a = 0, b = 0;
                rtdd_strip_frame_rows(ctx, 0, nullptr, &a, &b, nullptr, nullptr);
                e = cudaMemcpy2DAsync(m->hostOut + (size_t)a * m->pitchOut, m->pitchOut, ctx->depthU8 + (size_t)a * ctx->depthU8Pitch, ctx->depthU8Pitch,
                                      (size_t)ctx->cols, b - a, cudaMemcpyDeviceToHost, ctx->stream);
                rc = rtdd_check(ctx, e, "rtdd_mgpu_frame_solve_host_annotation (download)");
            }
            break;
        }
        default: rc = RTDD_E_ARG;
        }
        cudaEventRecord(e1, ctx->stream);
        const int rs = rtdd_sync(ctx);
        if (!rc) rc = rs;
        if (!rc) cudaEventElapsedTime(&ms, e0, e1);
        {
            std::lock_guard<std::mutex> lk(m->mu);
            m->status[r] = rc;
            m->ms[r] = ms;
            if (rc && m->err.empty()) m->err = "rank " + std::to_string(r) + ": " + rtdd_last_error(ctx);
            if (--m->pending == 0) m->cvDone.notify_all();
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int mgpu_run(rtdd_mgpu *m, int cmd, int iArg, float *msMax)
{
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->cmd = cmd; m->iArg = iArg;
        m->pending = m->n;
        m->err.clear();
        m->generation++;
    }
    m->cvWork.notify_all();
    std::unique_lock<std::mutex> lk(m->mu);
    m->cvDone.wait(lk, [&] { return m->pending == 0; });
    int rc = 0;
    float mx = 0.0f;
    for (int r = 0; r < m->n; r++) { if (m->status[r] && !rc) rc = m->status[r]; if (m->ms[r] > mx) mx = m->ms[r]; }
    if (msMax) *msMax = mx;
    return rc;
}

}  // namespace

int rtdd_mgpu_create(const int *devices, int ndevices, int rows, int cols, int levels, float beta, int halo, int passSweeps,
                     long long minStripPixels, rtdd_mgpu **out)
{
    if (!out) return RTDD_E_ARG;
    *out = nullptr;
    if (!devices || ndevices < 1 || rows < 1 || cols < 1) return RTDD_E_ARG;
    rtdd_mgpu *m = new (std::nothrow) rtdd_mgpu();
    if (!m) return RTDD_E_NOMEM;
    m->n = ndevices; m->rows = rows; m->cols = cols;
    m->levels = levels > 0 ? levels : rtdd_pyramid_levels(rows, cols);
    m->devices.assign(devices, devices + ndevices);
    m->ctx.assign(ndevices, nullptr);
    m->batchCtx.assign(ndevices, std::vector<rtdd_ctx *>());
    m->beta = beta;
    m->status.assign(ndevices, 0);
    m->ms.assign(ndevices, 0.0f);
    int rc = 0;
    for (int r = 0; r < ndevices && !rc; r++) {
        rc = ndevices > 1 ? rtdd_create_strip(rows, cols, m->levels, devices[r], ndevices, halo > 0 ? halo : 16, minStripPixels > 0 ? minStripPixels : (1LL << 22), &m->ctx[r])
                          : rtdd_create(rows, cols, m->levels, devices[r], &m->ctx[r]);
        if (!rc) rc = rtdd_load_weights(m->ctx[r], beta);
        if (!rc) rc = rtdd_strip_frame_setup(m->ctx[r], r, ndevices, halo > 0 ? halo : 16, passSweeps > 0 ? passSweeps : 8,
                                             minStripPixels > 0 ? minStripPixels : (1LL << 22));
    }
    // neighbours write each other's staging rows and flags directly: peer access both ways, arenas wired by pointer
    for (int r = 0; r + 1 < ndevices && !rc; r++) {
        const int a = devices[r], b = devices[r + 1];
        if (a == b) continue;
        int ok = 0;
        cudaDeviceCanAccessPeer(&ok, a, b);
        int ok2 = 0;
        cudaDeviceCanAccessPeer(&ok2, b, a);
        if (!ok || !ok2) { rc = RTDD_E_PEER; break; }
        cudaSetDevice(a);
        cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = RTDD_E_PEER;
        cudaGetLastError();
        cudaSetDevice(b);
        e = cudaDeviceEnablePeerAccess(a, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = RTDD_E_PEER;
        cudaGetLastError();
    }
    for (int r = 0; r < ndevices && !rc; r++)
        rc = rtdd_strip_set_peers(m->ctx[r], r > 0 ? m->ctx[r - 1]->arena : nullptr, r + 1 < ndevices ? m->ctx[r + 1]->arena : nullptr);
    if (rc) {
        for (rtdd_ctx *c : m->ctx) if (c) rtdd_destroy(c);
        delete m;
        return rc;
    }
    for (int r = 0; r < ndevices; r++) m->workers.emplace_back(mgpu_worker, m, r);
    *out = m;
    return 0;
}

int rtdd_mgpu_destroy(rtdd_mgpu *m)
{
    if (!m) return RTDD_E_ARG;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cvWork.notify_all();
    for (std::thread &t : m->workers) t.join();
    for (rtdd_ctx *c : m->ctx) if (c) rtdd_destroy(c);
    for (auto &v : m->batchCtx) for (rtdd_ctx *c : v) if (c) rtdd_destroy(c);
    delete m;
    return 0;
}

int rtdd_mgpu_devices(const rtdd_mgpu *m) { return m ? m->n : 0; }
const char *rtdd_mgpu_last_error(const rtdd_mgpu *m) { return m ? m->err.c_str() : "null handle"; }
rtdd_ctx *rtdd_mgpu_context(rtdd_mgpu *m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

int rtdd_mgpu_set_image(rtdd_mgpu *m, const uint8_t *bgrHost, size_t bgrPitch)
{
    if (!m || !bgrHost) return RTDD_E_ARG;
    m->hostA = bgrHost; m->pitchA = bgrPitch;
    return mgpu_run(m, 4, 0, nullptr);
}

int rtdd_mgpu_frame_solve_host_annotation(rtdd_mgpu *m, const uint8_t *annotationHost, size_t annotationPitch, int maxIterations,
                                          uint8_t *depthU8Host, size_t depthU8Pitch, float *msDevice)
{
    if (!m || !annotationHost) return RTDD_E_ARG;
    m->hostA = annotationHost; m->pitchA = annotationPitch;
    m->hostOut = depthU8Host; m->pitchOut = depthU8Pitch;
    return mgpu_run(m, 5, maxIterations, msDevice);
}

int rtdd_mgpu_frame_solve(rtdd_mgpu *m, int maxIterations, float *msDevice)
{
    if (!m) return RTDD_E_ARG;
    return mgpu_run(m, 1, maxIterations, msDevice);
}

int rtdd_mgpu_level0(rtdd_mgpu *m, int sweeps, float *msDevice)
{
    if (!m) return RTDD_E_ARG;
    return mgpu_run(m, 2, sweeps, msDevice);
}

int rtdd_mgpu_batch_solve(rtdd_mgpu *m, int nimages, const uint8_t *const *bgrHost, size_t bgrPitch, const uint8_t *const *annotationHost,
                          size_t annotationPitch, int maxIterations, uint8_t *const *depthU8Host, size_t depthU8Pitch, float *msDevice)
{
    if (!m || nimages < 0 || (nimages > 0 && (!bgrHost || !annotationHost))) return RTDD_E_ARG;
    m->nimages = nimages; m->bgrList = bgrHost; m->annList = annotationHost; m->outList = depthU8Host;
    m->bgrPitch = bgrPitch; m->annPitch = annotationPitch; m->outPitch = depthU8Pitch;
    return mgpu_run(m, 3, maxIterations, msDevice);
}

}  // extern "C"
