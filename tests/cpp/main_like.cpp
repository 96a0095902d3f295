// A headless stand-in for the reference's main.cpp solve loop (ref: src/main.cpp:149-155, 249-283), written only
// against the three reference-named headers and the CUDA runtime -- exactly what main.cpp itself uses.  It links
// against librtdd.so and prints a checksum of the solved depth map that tests/test_cpp_dropin.py
// compares with the same computation driven through the Python binding.
//
//   g++ -std=c++17 main_like.cpp -I<repo>/include -I/usr/local/cuda/include -L<repo>/realtimedepthdiffusion_b200/lib -lrtdd
//       -L/usr/local/cuda/lib64 -lcudart -o main_like
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "GPUDepthEffect.h"
#include "GPUImageProcessing.h"
#include "GPUSolver.h"

struct Plane {            // cv::cuda::GpuMat stand-in: pitched device memory
    unsigned char *ptr = nullptr;
    size_t step = 0;
    int rows = 0, cols = 0;
    void create(int r, int c, size_t elem) { rows = r; cols = c; cudaMallocPitch((void **)&ptr, &step, (size_t)c * elem, r); cudaMemset2D(ptr, step, 0, (size_t)c * elem, r); }
};

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 203, cols = argc > 2 ? atoi(argv[2]) : 317, maxIterations = argc > 3 ? atoi(argv[3]) : 100;
    const int levels = (int)log2((double)(std::max(std::min(cols, rows) / 45, 1))) + 1;           // main.cpp:95
    // deterministic inputs (the Python side regenerates the same bytes)
    std::vector<unsigned char> gray0((size_t)rows * cols), scribble0((size_t)rows * cols), edited0((size_t)rows * cols * 3);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            const size_t p = (size_t)y * cols + x;
            gray0[p] = (unsigned char)(((x / 23) * 37 + (y / 17) * 91 + (x * y) % 7) & 255);
            const bool s = ((x / 9) % 7 == 3) && ((y / 5) % 11 == 2);
            scribble0[p] = s ? 255 : 0;
            const unsigned char v = (unsigned char)(((x / 40 + y / 30) % 5) * 64 > 254 ? 254 : ((x / 40 + y / 30) % 5) * 64);
            edited0[3 * p] = edited0[3 * p + 1] = edited0[3 * p + 2] = s ? v : gray0[p];
        }
    std::vector<Plane> gray(levels), scribble(levels), edited(levels), depth(levels);
    std::vector<std::vector<unsigned char>> grayHost(levels);
    grayHost[0] = gray0;
    for (int l = 0; l < levels; l++) {
        const int r = (int)(rows / powf(2, l)), c = (int)(cols / powf(2, l));                       // main.cpp:103
        scribble[l].create(r, c, 1); edited[l].create(r, c, 3); depth[l].create(r, c, 4); gray[l].create(r, c, 1);
        std::vector<float> init((size_t)r * c, 255.0f);                                             // main.cpp:136
        cudaMemcpy2D(depth[l].ptr, depth[l].step, init.data(), (size_t)c * 4, (size_t)c * 4, r, cudaMemcpyHostToDevice);
        if (l > 0) {                                                                                // stand-in for cv::pyrDown: 2x2 box
            const int pr = (int)(rows / powf(2, l - 1)), pc = (int)(cols / powf(2, l - 1));
            grayHost[l].assign((size_t)r * c, 0);
            for (int y = 0; y < r; y++)
                for (int x = 0; x < c; x++) {
                    int s = 0;
                    for (int dy = 0; dy < 2; dy++) for (int dx = 0; dx < 2; dx++) s += grayHost[l - 1][(size_t)std::min(2 * y + dy, pr - 1) * pc + std::min(2 * x + dx, pc - 1)];
                    grayHost[l][(size_t)y * c + x] = (unsigned char)((s + 2) / 4);
                }
        }
        cudaMemcpy2D(gray[l].ptr, gray[l].step, grayHost[l].data(), c, c, r, cudaMemcpyHostToDevice);
    }
    cudaMemcpy2D(scribble[0].ptr, scribble[0].step, scribble0.data(), cols, cols, rows, cudaMemcpyHostToDevice);
    cudaMemcpy2D(edited[0].ptr, edited[0].step, edited0.data(), (size_t)cols * 3, (size_t)cols * 3, rows, cudaMemcpyHostToDevice);

    GPUAllocateDeviceMemory(rows, cols, levels);                                                    // main.cpp:149
    GPULoadWeights(0.4f);                                                                           // main.cpp:155
    GPUPaintImage(cols / 2, rows / 2, 192, 9, edited[0].ptr, edited[0].step, scribble[0].ptr, scribble[0].step, rows, cols);   // main.cpp:56
    for (int l = 1; l < levels; l++)                                                                // main.cpp:249
        GPUPyrDownAnnotation(scribble[l - 1].ptr, scribble[l - 1].step, edited[l - 1].ptr, edited[l - 1].step, edited[l - 1].rows, edited[l - 1].cols,
                             scribble[l].ptr, scribble[l].step, edited[l].ptr, edited[l].step, edited[l].rows, edited[l].cols);
    GPUConvertToFloat(edited[levels - 1].ptr, edited[levels - 1].step, (float *)depth[levels - 1].ptr, depth[levels - 1].step,
                      scribble[levels - 1].ptr, scribble[levels - 1].step, edited[levels - 1].rows, edited[levels - 1].cols);      // main.cpp:257
    for (int l = levels - 1; l >= 0; l--) {
        const int iters = (int)(maxIterations / powf(2.0f, (float)((levels - 1) - l)));            // main.cpp:263
        GPUMatrixFreeSolver((float *)depth[l].ptr, depth[l].step, scribble[l].ptr, scribble[l].step, gray[l].ptr, gray[l].step,
                            depth[l].rows, depth[l].cols, 0.4f, iters, 1e-5f, l);                  // main.cpp:266
        if (l > 0) {                                                                                // stand-in for cv::pyrUp: nearest neighbour
            const int r = depth[l - 1].rows, c = depth[l - 1].cols, cr = depth[l].rows, cc = depth[l].cols;
            std::vector<float> coarse((size_t)cr * cc), fine((size_t)r * c);
            cudaMemcpy2D(coarse.data(), (size_t)cc * 4, depth[l].ptr, depth[l].step, (size_t)cc * 4, cr, cudaMemcpyDeviceToHost);
            for (int y = 0; y < r; y++)
                for (int x = 0; x < c; x++) fine[(size_t)y * c + x] = coarse[(size_t)std::min(y / 2, cr - 1) * cc + std::min(x / 2, cc - 1)];
            cudaMemcpy2D(depth[l - 1].ptr, depth[l - 1].step, fine.data(), (size_t)c * 4, (size_t)c * 4, r, cudaMemcpyHostToDevice);
            GPUConvertToFloat(edited[l - 1].ptr, edited[l - 1].step, (float *)depth[l - 1].ptr, depth[l - 1].step, scribble[l - 1].ptr,
                              scribble[l - 1].step, edited[l - 1].rows, edited[l - 1].cols);       // main.cpp:281
        }
    }
    Plane art; art.create(rows, cols, 3);
    GPUSimulateHaze(edited[0].ptr, edited[0].step, (float *)depth[0].ptr, depth[0].step, art.ptr, art.step, rows, cols);           // main.cpp:220
    GPUSimulateDesaturation(edited[0].ptr, edited[0].step, gray[0].ptr, gray[0].step, (float *)depth[0].ptr, depth[0].step, art.ptr, art.step, rows, cols);
    GPUSimulateDefocus(edited[0].ptr, edited[0].step, (float *)depth[0].ptr, depth[0].step, art.ptr, art.step, rows, cols);        // main.cpp:192
    std::vector<float> out((size_t)rows * cols);
    std::vector<unsigned char> artHost((size_t)rows * cols * 3);
    cudaMemcpy2D(out.data(), (size_t)cols * 4, depth[0].ptr, depth[0].step, (size_t)cols * 4, rows, cudaMemcpyDeviceToHost);
    cudaMemcpy2D(artHost.data(), (size_t)cols * 3, art.ptr, art.step, (size_t)cols * 3, rows, cudaMemcpyDeviceToHost);
    GPUFreeDeviceMemory(levels);                                                                    // main.cpp:336
    // FNV-1a over the raw bits
    uint64_t h = 1469598103934665603ull;
    for (float f : out) { uint32_t b; memcpy(&b, &f, 4); for (int k = 0; k < 4; k++) { h ^= (b >> (8 * k)) & 255; h *= 1099511628211ull; } }
    uint64_t ha = 1469598103934665603ull;
    for (unsigned char c : artHost) { ha ^= c; ha *= 1099511628211ull; }
    const cudaError_t e = cudaDeviceSynchronize();
    printf("levels %d depth_fnv %016llx defocus_fnv %016llx cuda %d\n", levels, (unsigned long long)h, (unsigned long long)ha, (int)e);
    return e == cudaSuccess ? 0 : 1;
}
