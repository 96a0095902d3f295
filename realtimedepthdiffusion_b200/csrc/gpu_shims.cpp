// The reference's ten free functions (C++ linkage, identical signatures) as thin
// shims over the C ABI of include/rtdd.h.  They keep the reference's conventions:
// void returns, "<function>: <error>" printed and execution continues
// (ref: src/GPUSolver.cu:21-27), one process-global solver instance
// (ref: src/GPUSolver.cu:13-19), and a device-wide sync at the end of the four
// GPUSolver functions (GPUCheckError) but not after the image/effect functions.
//
// main.cpp links against these unchanged.

#include <cstdio>

#include "GPUDepthEffect.h"
#include "GPUImageProcessing.h"
#include "GPUSolver.h"
#include "rtdd.h"

namespace {

rtdd_ctx *g_ctx = nullptr;       // the reference's file-scope globals, folded into one handle
rtdd_ctx *g_aux = nullptr;       // context for image/effect calls made before GPUAllocateDeviceMemory

void report(const char *fn, rtdd_ctx *ctx, int rc)
{
    if (rc == 0) return;
    const char *msg = ctx ? rtdd_last_error(ctx) : "no context";
    std::printf("%s: %s\n", fn, msg);
}

rtdd_ctx *any_ctx(const char *fn)
{
    if (g_ctx) return g_ctx;
    if (!g_aux) {
        const int rc = rtdd_create(1, 1, 1, -1, &g_aux);
        if (rc) { std::printf("%s: cannot create a device context (status %d)\n", fn, rc); return nullptr; }
    }
    return g_aux;
}

}  // namespace

void GPUAllocateDeviceMemory(int rows, int cols, int levels)
{
    if (g_ctx) { rtdd_destroy(g_ctx); g_ctx = nullptr; }
    const int rc = rtdd_create(rows, cols, levels, -1, &g_ctx);
    if (rc) std::printf("GPUAllocateDeviceMemory: status %d\n", rc);
}

void GPUFreeDeviceMemory(int levels)
{
    (void)levels;
    if (g_ctx) {
        const int rc = rtdd_destroy(g_ctx);   // synchronises the device first, like GPUCheckError
        if (rc) std::printf("GPUFreeDeviceMemory: status %d\n", rc);
        g_ctx = nullptr;
    }
    if (g_aux) { rtdd_destroy(g_aux); g_aux = nullptr; }
}

void GPULoadWeights(float beta)
{
    if (!g_ctx) { std::printf("GPULoadWeights: GPUAllocateDeviceMemory has not been called\n"); return; }
    report("GPULoadWeights", g_ctx, rtdd_load_weights(g_ctx, beta));
}

void GPUMatrixFreeSolver(float *depthImage, size_t depthPitch, unsigned char *scribbleImage, size_t scribblePitch, unsigned char *grayImage,
	size_t grayPitch, int rows, int cols, float beta, int maxIterations, float tolerance, int level)
{
    (void)beta; (void)tolerance;   // accepted and ignored, as in the reference (src/GPUSolver.cu:274-275)
    if (!g_ctx) { std::printf("GPUMatrixFreeSolver: GPUAllocateDeviceMemory has not been called\n"); return; }
    int rc = rtdd_solve_level(g_ctx, depthImage, depthPitch, scribbleImage, scribblePitch, grayImage, grayPitch, rows, cols, maxIterations, level);
    if (rc == 0) rc = rtdd_sync(g_ctx);   // GPUCheckError's cudaThreadSynchronize (src/GPUSolver.cu:314)
    report("GPUMatrixFreeSolver", g_ctx, rc);
}

void GPUConvertToFloat(unsigned char *src, size_t srcPitch, float *dst, size_t dstPitch, unsigned char *mask, size_t maskPitch,
	int rows, int cols)
{
    rtdd_ctx *c = any_ctx("GPUConvertToFloat");
    if (c) report("GPUConvertToFloat", c, rtdd_convert_to_float(c, src, srcPitch, dst, dstPitch, mask, maskPitch, rows, cols));
}

void GPUPyrDownAnnotation(unsigned char *prevScribbleImage, size_t prevScribblePitch, unsigned char *prevEditedImage,
	size_t prevEditedPitch, int previousRows, int previousCols, unsigned char *currScribbleImage, size_t currScribblePitch,
	unsigned char *currEditedImage, size_t currEditedPitch, int currentRows, int currentCols)
{
    rtdd_ctx *c = any_ctx("GPUPyrDownAnnotation");
    if (c) report("GPUPyrDownAnnotation", c, rtdd_pyrdown_annotation(c, prevScribbleImage, prevScribblePitch, prevEditedImage, prevEditedPitch,
        previousRows, previousCols, currScribbleImage, currScribblePitch, currEditedImage, currEditedPitch, currentRows, currentCols));
}

void GPUPaintImage(int x, int y, int scribbleColor, int scribbleRadius, unsigned char *editedImage, size_t editedPitch,
	unsigned char *scribbleImage, size_t scribblePitch, int rows, int cols)
{
    rtdd_ctx *c = any_ctx("GPUPaintImage");
    if (c) report("GPUPaintImage", c, rtdd_paint(c, x, y, scribbleColor, scribbleRadius, editedImage, editedPitch, scribbleImage, scribblePitch, rows, cols));
}

void GPUSimulateDefocus(unsigned char *originalImage, size_t originalPitch, float *depthImage, size_t depthPitch,
	unsigned char *artisticImage, size_t artisticPitch, int rows, int cols)
{
    rtdd_ctx *c = any_ctx("GPUSimulateDefocus");
    if (c) report("GPUSimulateDefocus", c, rtdd_defocus(c, originalImage, originalPitch, depthImage, depthPitch, artisticImage, artisticPitch, rows, cols));
}

void GPUSimulateDesaturation(unsigned char *originalImage, size_t originalPitch, unsigned char *grayImage, size_t grayPitch,
	float *depthImage, size_t depthPitch, unsigned char *artisticImage, size_t artisticPitch, int rows, int cols)
{
    rtdd_ctx *c = any_ctx("GPUSimulateDesaturation");
    if (c) report("GPUSimulateDesaturation", c, rtdd_desaturate(c, originalImage, originalPitch, grayImage, grayPitch, depthImage, depthPitch,
        artisticImage, artisticPitch, rows, cols));
}

void GPUSimulateHaze(unsigned char *originalImage, size_t originalPitch, float *depthImage, size_t depthPitch,
	unsigned char *artisticImage, size_t artisticPitch, int rows, int cols)
{
    rtdd_ctx *c = any_ctx("GPUSimulateHaze");
    if (c) report("GPUSimulateHaze", c, rtdd_haze(c, originalImage, originalPitch, depthImage, depthPitch, artisticImage, artisticPitch, rows, cols));
}
