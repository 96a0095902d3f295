// A C++ host for the multi-GPU entry points of include/rtdd.h, the way a main.cpp-like application would use them
// (north_star: "host code is C++ ... through a thin C-ABI"; the loop restated inside the library is src/main.cpp:232-295).
//   mgpu_host <ngpus> <rows> <cols> <maxIterations> [minStripPixels]
// 1. configs[4]: one image cut into row strips over <ngpus> GPUs (rtdd_mgpu_frame_solve_host_annotation) -- the 8-bit depth map must
//    equal, byte for byte, the one a single context on GPU 0 produces (rtdd_frame_solve_host_annotation);
// 2. DepthEffect row strips (rtdd_strip_frame_effects on every rank's context) against the single-GPU effects;
// 3. configs[3]: a batch of 5 images, image i on GPU i mod N (rtdd_mgpu_batch_solve), against the single context.
// Prints one line per check and exits non-zero on any mismatch.
//
//   g++ -std=c++17 mgpu_host.cpp -I<repo>/include -I/usr/local/cuda/include -L<repo>/realtimedepthdiffusion_b200/lib -lrtdd
//       -L/usr/local/cuda/lib64 -lcudart -o mgpu_host
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "rtdd.h"

static void synth(int rows, int cols, unsigned seed, std::vector<uint8_t> &bgr, std::vector<uint8_t> &ann)
{
    bgr.resize((size_t)rows * cols * 3);
    ann.resize((size_t)rows * cols);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            const size_t p = (size_t)y * cols + x;
            const unsigned h = (unsigned)(x * 2654435761u) ^ (unsigned)(y * 40503u) ^ (seed * 97u);
            const int blk = ((x / 61 + seed) * 37 + (y / 47) * 91) & 255;
            bgr[3 * p] = (uint8_t)((blk + (h >> 28)) & 255);
            bgr[3 * p + 1] = (uint8_t)((blk * 3 + (h >> 27 & 7)) & 255);
            bgr[3 * p + 2] = (uint8_t)((255 - blk + (h >> 26 & 3)) & 255);
            const bool s = ((x / 13 + seed) % 9 == 3) && ((y / 7) % 8 == 2);
            const int v = ((x / 90 + y / 70 + (int)seed) % 5) * 64;
            ann[p] = s ? (uint8_t)(v > 254 ? 254 : v) : 32;                       // main.cpp:160-170: 32 = not annotated
        }
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        const int _rc = (call);                                                                      \
        if (_rc) { fprintf(stderr, "%s failed with status %d (line %d)\n", #call, _rc, __LINE__); return 2; } \
    } while (0)

int main(int argc, char **argv)
{
    const int ngpus = argc > 1 ? atoi(argv[1]) : 2, rows = argc > 2 ? atoi(argv[2]) : 1536, cols = argc > 3 ? atoi(argv[3]) : 2048;
    const int iters = argc > 4 ? atoi(argv[4]) : 1000;
    const long long minStrip = argc > 5 ? atoll(argv[5]) : 200000;
    int have = 0;
    cudaGetDeviceCount(&have);
    if (have < ngpus) { printf("skipped: %d GPUs visible, %d wanted\n", have, ngpus); return 0; }
    std::vector<int> devices(ngpus);
    for (int i = 0; i < ngpus; i++) devices[i] = i;
    std::vector<uint8_t> bgr, ann;
    synth(rows, cols, 1, bgr, ann);
    int bad = 0;

    // ---- single GPU ----
    rtdd_ctx *one = nullptr;
    const int levels = rtdd_pyramid_levels(rows, cols);
    CK(rtdd_create(rows, cols, levels, 0, &one));
    CK(rtdd_load_weights(one, 0.4f));
    CK(rtdd_frame_set_image(one, bgr.data(), (size_t)cols * 3));
    std::vector<uint8_t> want((size_t)rows * cols), got((size_t)rows * cols, 0);
    // three frames on both sides: like main.cpp, a frame starts from the previous frame's coarsest-level solution (main.cpp:257)
    for (int rep = 0; rep < 3; rep++) CK(rtdd_frame_solve_host_annotation(one, ann.data(), cols, iters, want.data(), cols));

    // ---- 1. row strips over ngpus ----
    rtdd_mgpu *m = nullptr;
    CK(rtdd_mgpu_create(devices.data(), ngpus, rows, cols, levels, 0.4f, 8, 8, minStrip, &m));
    CK(rtdd_mgpu_set_image(m, bgr.data(), (size_t)cols * 3));
    float ms = 0.0f;
    for (int rep = 0; rep < 3; rep++) {
        std::fill(got.begin(), got.end(), 0);
        const int rc = rtdd_mgpu_frame_solve_host_annotation(m, ann.data(), cols, iters, got.data(), cols, &ms);
        if (rc) { fprintf(stderr, "rtdd_mgpu_frame_solve_host_annotation: %d %s\n", rc, rtdd_mgpu_last_error(m)); return 2; }
    }
    size_t diff = 0;
    for (size_t i = 0; i < want.size(); i++) diff += (want[i] != got[i]);
    int split0 = 0;
    rtdd_strip_frame_rows(rtdd_mgpu_context(m, 0), 0, &split0, nullptr, nullptr, nullptr, nullptr);
    printf("strips: %d GPUs, %dx%d, %d levels, level 0 split %d, %.3f ms per frame (device, slowest rank), differing bytes %zu\n", ngpus, cols, rows, levels,
           split0, ms, diff);
    bad += diff != 0;

    // ---- 2. DepthEffect row strips: every rank fills its rows of its own output planes; stitch on the host ----
    {
        const size_t rowBytes = (size_t)cols * 3;
        std::vector<uint8_t> ref[3], strip[3];
        uint8_t *d[3];
        size_t pitch = 0;
        cudaSetDevice(0);
        for (int k = 0; k < 3; k++) { cudaMallocPitch((void **)&d[k], &pitch, rowBytes, rows); ref[k].resize(rowBytes * rows); strip[k].assign(rowBytes * rows, 0); }
        CK(rtdd_frame_effects(one, d[0], pitch, d[1], pitch, d[2], pitch));
        CK(rtdd_sync(one));
        for (int k = 0; k < 3; k++) { cudaMemcpy2D(ref[k].data(), rowBytes, d[k], pitch, rowBytes, rows, cudaMemcpyDeviceToHost); cudaFree(d[k]); }
        for (int r = 0; r < ngpus; r++) {
            rtdd_ctx *c = rtdd_mgpu_context(m, r);
            int a = 0, b = 0;
            rtdd_strip_frame_rows(c, 0, nullptr, &a, &b, nullptr, nullptr);
            cudaSetDevice(devices[r]);
            for (int k = 0; k < 3; k++) cudaMallocPitch((void **)&d[k], &pitch, rowBytes, rows);
            CK(rtdd_strip_frame_effects(c, d[0], pitch, d[1], pitch, d[2], pitch));
            CK(rtdd_sync(c));
            for (int k = 0; k < 3; k++) {
                cudaMemcpy2D(strip[k].data() + (size_t)a * rowBytes, rowBytes, d[k] + (size_t)a * pitch, pitch, rowBytes, b - a, cudaMemcpyDeviceToHost);
                cudaFree(d[k]);
            }
        }
        size_t de[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++)
            for (size_t i = 0; i < ref[k].size(); i++) de[k] += (ref[k][i] != strip[k][i]);
        printf("effects on row strips: differing bytes desaturation %zu haze %zu defocus %zu\n", de[0], de[1], de[2]);
        bad += (de[0] + de[1] + de[2]) != 0;
    }

    // ---- 3. batch: 5 images, image i on GPU i mod N ----
    {
        const int nimg = 5;
        std::vector<std::vector<uint8_t>> B(nimg), A(nimg), O(nimg), W(nimg);
        std::vector<const uint8_t *> pb(nimg), pa(nimg);
        std::vector<uint8_t *> po(nimg);
        for (int i = 0; i < nimg; i++) {
            synth(rows, cols, 10 + i, B[i], A[i]);
            O[i].assign((size_t)rows * cols, 0);
            W[i].assign((size_t)rows * cols, 0);
            pb[i] = B[i].data(); pa[i] = A[i].data(); po[i] = O[i].data();
            CK(rtdd_frame_set_image(one, B[i].data(), (size_t)cols * 3));
            CK(rtdd_frame_solve_host_annotation(one, A[i].data(), cols, iters, W[i].data(), cols));
        }
        const int rc = rtdd_mgpu_batch_solve(m, nimg, pb.data(), (size_t)cols * 3, pa.data(), cols, iters, po.data(), cols, &ms);
        if (rc) { fprintf(stderr, "rtdd_mgpu_batch_solve: %d %s\n", rc, rtdd_mgpu_last_error(m)); return 2; }
        size_t db = 0;
        for (int i = 0; i < nimg; i++)
            for (size_t k = 0; k < W[i].size(); k++) db += (W[i][k] != O[i][k]);
        printf("batch: %d images over %d GPUs, %.3f ms (device, slowest rank), differing bytes %zu\n", nimg, ngpus, ms, db);
        bad += db != 0;
    }
    CK(rtdd_mgpu_destroy(m));
    CK(rtdd_destroy(one));
    printf("%s\n", bad ? "MISMATCH" : "all identical");
    return bad ? 1 : 0;
}
