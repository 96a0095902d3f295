// A C++ host for the per-frame download of include/rtdd.h (ref: src/main.cpp:291): the 8-bit depth map must hold the same bytes
// whether the caller's plane is pageable (a copy after the last pass), page-locked by cudaHostRegister on an ordinary allocation,
// or allocated by cudaHostAlloc (both: stored by the last sweep pass itself, "zero_copy_out" in rtdd.h) -- over two frames, the
// second one the live loop's way (rtdd_frame_paint + rtdd_frame_solve_download).
//   pinned_map <rows> <cols> <maxIterations>
//   g++ -std=c++17 pinned_map.cpp -I<repo>/include -I/usr/local/cuda/include -L<repo>/realtimedepthdiffusion_b200/lib -lrtdd
//       -L/usr/local/cuda/lib64 -lcudart -o pinned_map
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "rtdd.h"

#define CK(call)                                                                                     \
    do {                                                                                             \
        const int _rc = (call);                                                                      \
        if (_rc) { fprintf(stderr, "%s failed with status %d (line %d)\n", #call, _rc, __LINE__); return 2; } \
    } while (0)

static int run(int rows, int cols, int iters, const std::vector<uint8_t> &bgr, const std::vector<uint8_t> &ann, uint8_t *out, size_t pitch,
               std::vector<uint8_t> &frame1, std::vector<uint8_t> &frame2)
{
    rtdd_ctx *ctx = nullptr;
    CK(rtdd_create(rows, cols, rtdd_pyramid_levels(rows, cols), 0, &ctx));
    CK(rtdd_load_weights(ctx, 0.4f));
    CK(rtdd_frame_set_image(ctx, bgr.data(), (size_t)cols * 3));
    CK(rtdd_frame_solve_host_annotation(ctx, ann.data(), (size_t)cols, iters, out, pitch));
    frame1.resize((size_t)rows * cols);
    for (int y = 0; y < rows; y++) memcpy(&frame1[(size_t)y * cols], out + (size_t)y * pitch, (size_t)cols);
    CK(rtdd_frame_paint(ctx, cols / 3, rows / 2, 200, rows / 40 + 2));
    CK(rtdd_frame_solve_download(ctx, iters, out, pitch));
    frame2.resize((size_t)rows * cols);
    for (int y = 0; y < rows; y++) memcpy(&frame2[(size_t)y * cols], out + (size_t)y * pitch, (size_t)cols);
    CK(rtdd_destroy(ctx));
    return 0;
}

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 540, cols = argc > 2 ? atoi(argv[2]) : 960, iters = argc > 3 ? atoi(argv[3]) : 1000;
    std::vector<uint8_t> bgr((size_t)rows * cols * 3), ann((size_t)rows * cols);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            const size_t p = (size_t)y * cols + x;
            const int blk = ((x / 53) * 37 + (y / 41) * 91) & 255;
            bgr[3 * p] = (uint8_t)blk; bgr[3 * p + 1] = (uint8_t)((blk * 3) & 255); bgr[3 * p + 2] = (uint8_t)(255 - blk);
            const bool s = ((x / 11) % 9 == 3) && ((y / 7) % 8 == 2);
            ann[p] = s ? (uint8_t)(((x / 90 + y / 70) % 4) * 64) : 32;                     // main.cpp:160-170: 32 = not annotated
        }
    const size_t pitch = ((size_t)cols + 3) / 4 * 4 + 64;                                   // a pitched plane, 4-byte aligned rows
    std::vector<uint8_t> a1, a2, b1, b2, c1, c2;

    std::vector<uint8_t> pageable(pitch * rows, 7);
    if (run(rows, cols, iters, bgr, ann, pageable.data(), pitch, a1, a2)) return 2;

    uint8_t *plain = (uint8_t *)aligned_alloc(4096, (pitch * rows + 4095) / 4096 * 4096);
    memset(plain, 7, pitch * rows);
    if (cudaHostRegister(plain, pitch * rows, cudaHostRegisterDefault) != cudaSuccess) { fprintf(stderr, "cudaHostRegister failed\n"); return 2; }
    if (run(rows, cols, iters, bgr, ann, plain, pitch, b1, b2)) return 2;
    int marginOk = 1;
    for (int y = 0; y < rows && marginOk; y++)
        for (size_t x = cols; x < pitch; x++) if (plain[(size_t)y * pitch + x] != 7) { marginOk = 0; break; }
    cudaHostUnregister(plain);
    free(plain);

    uint8_t *pinned = nullptr;
    if (cudaHostAlloc((void **)&pinned, pitch * rows, cudaHostAllocDefault) != cudaSuccess) { fprintf(stderr, "cudaHostAlloc failed\n"); return 2; }
    if (run(rows, cols, iters, bgr, ann, pinned, pitch, c1, c2)) return 2;
    cudaFreeHost(pinned);

    const int same = (a1 == b1) && (a1 == c1) && (a2 == b2) && (a2 == c2);
    const int moved = (a1 != a2);
    printf("pinned_map %dx%d: registered/hostalloc vs pageable %s, second frame differs from the first: %s, margin untouched: %s\n", cols, rows,
           same ? "identical" : "MISMATCH", moved ? "yes" : "NO", marginOk ? "yes" : "NO");
    return (same && moved && marginOk) ? 0 : 1;
}
