// GPUImageProcessing kernels (Dirichlet injection, annotation restriction, brush)
// and the pyramid ops that sit either side of the solve (BGR->gray, gray pyrDown,
// depth pyrUp, u8 quantiser).
//
// ref: src/GPUImageProcessing.cu:8-100; src/main.cpp:111-112,143-145,272-279,290.

#include "rtdd_internal.h"

namespace rtdd {

// ---------------------------------------------------------------------------
// convert: dst[y][x] = (float)src[y][3x] where mask[y][x] == 255
// ref: src/GPUImageProcessing.cu:8-21
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
convert_kernel(const uint8_t *__restrict__ src, size_t srcPitch, float *__restrict__ dst, size_t dstPitch,
               const uint8_t *__restrict__ mask, size_t maskPitch, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    if (__ldg(mask + (size_t)y * maskPitch + x) == 255) {
        float *dRow = (float *)((char *)dst + (size_t)y * dstPitch);
        dRow[x] = (float)__ldg(src + (size_t)y * srcPitch + 3 * x);
    }
}

cudaError_t launch_convert(cudaStream_t s, const uint8_t *src, size_t srcPitch, float *dst, size_t dstPitch,
                           const uint8_t *mask, size_t maskPitch, int rows, int cols)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(cols, block.x), rtdd_div_up(rows, block.y));
    convert_kernel<<<grid, block, 0, s>>>(src, srcPitch, dst, dstPitch, mask, maskPitch, rows, cols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// annotation restriction.  The reference scans rows {2y-1, 2y} x cols {2x-1, 2x}
// in row-major order and lets every scribbled hit overwrite the output, so the
// winner is the LAST hit: (2y,2x) > (2y,2x-1) > (2y-1,2x) > (2y-1,2x-1).
// Outputs are left untouched when nothing is scribbled (planes are never cleared).
// ref: src/GPUImageProcessing.cu:23-49
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pyrdown_annotation_kernel(const uint8_t *__restrict__ prevScribble, size_t prevScribblePitch,
                          const uint8_t *__restrict__ prevEdited, size_t prevEditedPitch, int previousRows, int previousCols,
                          uint8_t *__restrict__ currScribble, size_t currScribblePitch,
                          uint8_t *__restrict__ currEdited, size_t currEditedPitch, int currentRows, int currentCols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= currentCols || y >= currentRows) return;
    int hitX = -1, hitY = -1;
#pragma unroll
    for (int dy = 0; dy >= -1; dy--) {
#pragma unroll
        for (int dx = 0; dx >= -1; dx--) {
            const int px = 2 * x + dx, py = 2 * y + dy;
            if (hitX < 0 && px >= 0 && py >= 0 && px < previousCols && py < previousRows &&
                __ldg(prevScribble + (size_t)py * prevScribblePitch + px) == 255) {
                hitX = px; hitY = py;
            }
        }
    }
    if (hitX >= 0) {
        currScribble[(size_t)y * currScribblePitch + x] = 255;
        currEdited[(size_t)y * currEditedPitch + 3 * x] = __ldg(prevEdited + (size_t)hitY * prevEditedPitch + 3 * hitX);
    }
}

cudaError_t launch_pyrdown_annotation(cudaStream_t s, const uint8_t *prevScribble, size_t prevScribblePitch,
                                      const uint8_t *prevEdited, size_t prevEditedPitch, int previousRows, int previousCols,
                                      uint8_t *currScribble, size_t currScribblePitch, uint8_t *currEdited, size_t currEditedPitch,
                                      int currentRows, int currentCols)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(currentCols, block.x), rtdd_div_up(currentRows, block.y));
    pyrdown_annotation_kernel<<<grid, block, 0, s>>>(prevScribble, prevScribblePitch, prevEdited, prevEditedPitch,
                                                     previousRows, previousCols, currScribble, currScribblePitch,
                                                     currEdited, currEditedPitch, currentRows, currentCols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// brush: the reference launches a full-image grid and lets almost every thread
// return; here the grid covers only the clipped brush rectangle
// [x-r/2, x+r/2] x [y-r/2, y+r/2] (integer r/2, truncating division).
// ref: src/GPUImageProcessing.cu:51-70
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
paint_kernel(int x0, int y0, int x1, int y1, int color, uint8_t *__restrict__ edited, size_t editedPitch,
             uint8_t *__restrict__ scribble, size_t scribblePitch)
{
    const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x > x1 || y > y1) return;
    uint8_t *e = edited + (size_t)y * editedPitch + 3 * x;
    e[0] = (uint8_t)color; e[1] = (uint8_t)color; e[2] = (uint8_t)color;
    scribble[(size_t)y * scribblePitch + x] = 255;
}

cudaError_t launch_paint(cudaStream_t s, int x, int y, int color, int radius, uint8_t *edited, size_t editedPitch,
                         uint8_t *scribble, size_t scribblePitch, int rows, int cols, int *launched)
{
    const int h = radius / 2;
    *launched = 0;
    if (h < 0) return cudaSuccess;   // x - h > x + h: the reference's tests reject every thread
    int x0 = x - h, x1 = x + h, y0 = y - h, y1 = y + h;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > cols - 1) x1 = cols - 1;
    if (y1 > rows - 1) y1 = rows - 1;
    if (x1 < x0 || y1 < y0) return cudaSuccess;
    dim3 block(32, 8);
    dim3 grid(rtdd_div_up(x1 - x0 + 1, block.x), rtdd_div_up(y1 - y0 + 1, block.y));
    paint_kernel<<<grid, block, 0, s>>>(x0, y0, x1, y1, color, edited, editedPitch, scribble, scribblePitch);
    *launched = 1;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// cv::cvtColor(BGR2GRAY), 15-bit fixed point (ref: src/main.cpp:111,138)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bgr2gray_kernel(const uint8_t *__restrict__ bgr, size_t bgrPitch, uint8_t *__restrict__ gray, size_t grayPitch, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const uint8_t *p = bgr + (size_t)y * bgrPitch + 3 * x;
    const int v = (int)__ldg(p) * 3735 + (int)__ldg(p + 1) * 19235 + (int)__ldg(p + 2) * 9798 + (1 << 14);
    gray[(size_t)y * grayPitch + x] = (uint8_t)(v >> 15);
}

cudaError_t launch_bgr2gray(cudaStream_t s, const uint8_t *bgr, size_t bgrPitch, uint8_t *gray, size_t grayPitch, int rows, int cols)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(cols, block.x), rtdd_div_up(rows, block.y));
    bgr2gray_kernel<<<grid, block, 0, s>>>(bgr, bgrPitch, gray, grayPitch, rows, cols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// cv::pyrDown (u8): separable [1 4 6 4 1], BORDER_REFLECT_101, (v + 128) >> 8
// ref: src/main.cpp:112,143-145,244-246
// ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
pyrdown_gray_kernel(const uint8_t *__restrict__ src, size_t srcPitch, int srows, int scols,
                    uint8_t *__restrict__ dst, size_t dstPitch, int drows, int dcols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    const int k[5] = {1, 4, 6, 4, 1};
    int cx[5];
#pragma unroll
    for (int i = 0; i < 5; i++) cx[i] = reflect101(2 * x + i - 2, scols);
    int v = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint8_t *r = src + (size_t)reflect101(2 * y + j - 2, srows) * srcPitch;
        int h = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) h += k[i] * (int)__ldg(r + cx[i]);
        v += k[j] * h;
    }
    dst[(size_t)y * dstPitch + x] = (uint8_t)((v + 128) >> 8);
}

cudaError_t launch_pyrdown_gray(cudaStream_t s, const uint8_t *src, size_t srcPitch, int srows, int scols, uint8_t *dst, size_t dstPitch)
{
    const int drows = (srows + 1) / 2, dcols = (scols + 1) / 2;
    dim3 block(32, 8);
    dim3 grid(rtdd_div_up(dcols, block.x), rtdd_div_up(drows, block.y));
    pyrdown_gray_kernel<<<grid, block, 0, s>>>(src, srcPitch, srows, scols, dst, dstPitch, drows, dcols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// cv::pyrUp (fp32), any destination of 2n or 2n+1 per dimension.  No multiply-add
// is contracted (explicit __fmul_rn/__fadd_rn) so the result is bit-equal to
// OpenCV's unfused CPU path and to oracle_pyrup_f32.
// ref: src/main.cpp:272-279
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pyrup_h(const float *__restrict__ s, int n, int dx)
{
    // horizontal pass value at destination column dx of one source row
    if (n == 1) return __fmul_rn(__ldg(s), 8.0f);
    if (dx >= 2 * n) dx = 2 * n - 1;                       // odd destination width: repeat last column
    const int x = dx >> 1;
    if (dx & 1) {
        if (x == n - 1) return __fmul_rn(__ldg(s + x), 8.0f);
        return __fmul_rn(__fadd_rn(__ldg(s + x), __ldg(s + x + 1)), 4.0f);
    }
    if (x == 0) return __fadd_rn(__fmul_rn(__ldg(s), 6.0f), __fmul_rn(__ldg(s + 1), 2.0f));
    if (x == n - 1) return __fadd_rn(__ldg(s + x - 1), __fmul_rn(__ldg(s + x), 7.0f));
    return __fadd_rn(__fadd_rn(__ldg(s + x - 1), __fmul_rn(__ldg(s + x), 6.0f)), __ldg(s + x + 1));
}

__global__ void __launch_bounds__(256)
pyrup_depth_kernel(const float *__restrict__ src, size_t srcPitch, int srows, int scols,
                   float *__restrict__ dst, size_t dstPitch, int drows, int dcols, int dy0)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    int dy = dy0 + blockIdx.y * blockDim.y + threadIdx.y;   // rows [dy0, drows) of the destination
    if (dx >= dcols || dy >= drows) return;
    float *out = (float *)((char *)dst + (size_t)dy * dstPitch) + dx;
    if (dy >= 2 * srows) dy = 2 * srows - 2;               // odd destination height: repeat row 2n-2
    const int y = dy >> 1;
    const int ym = (y > 0) ? y - 1 : (srows > 1 ? 1 : 0);
    const int yp = (y < srows - 1) ? y + 1 : srows - 1;
    const float *s1 = (const float *)((const char *)src + (size_t)y * srcPitch);
    const float *s2 = (const float *)((const char *)src + (size_t)yp * srcPitch);
    const float r1 = pyrup_h(s1, scols, dx);
    const float r2 = pyrup_h(s2, scols, dx);
    if (dy & 1) {
        *out = __fmul_rn(__fadd_rn(r1, r2), 1.0f / 16.0f);
    } else {
        const float *s0 = (const float *)((const char *)src + (size_t)ym * srcPitch);
        const float r0 = pyrup_h(s0, scols, dx);
        *out = __fmul_rn(__fadd_rn(__fadd_rn(r0, __fmul_rn(r1, 6.0f)), r2), 1.0f / 64.0f);
    }
}

cudaError_t launch_pyrup_depth(cudaStream_t s, const float *src, size_t srcPitch, int srows, int scols,
                               float *dst, size_t dstPitch, int drows, int dcols)
{
    return launch_pyrup_depth_rows(s, src, srcPitch, srows, scols, dst, dstPitch, drows, dcols, 0, drows);
}

// rows [rowBegin, rowEnd) of the destination only (row-strip decomposition); src/dst are the full planes
cudaError_t launch_pyrup_depth_rows(cudaStream_t s, const float *src, size_t srcPitch, int srows, int scols,
                                    float *dst, size_t dstPitch, int drows, int dcols, int rowBegin, int rowEnd)
{
    if (rowEnd <= rowBegin) return cudaSuccess;
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(dcols, block.x), rtdd_div_up(rowEnd - rowBegin, block.y));
    pyrup_depth_kernel<<<grid, block, 0, s>>>(src, srcPitch, srows, scols, dst, dstPitch, rowEnd, dcols, rowBegin);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// GpuMat::convertTo(CV_8UC1): round half to even, saturate (ref: src/main.cpp:290)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quantise_kernel(const float *__restrict__ src, size_t srcPitch, uint8_t *__restrict__ dst, size_t dstPitch, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const float v = __ldg((const float *)((const char *)src + (size_t)y * srcPitch) + x);
    int q = __float2int_rn(v);          // cvt.rni.s32.f32: half to even, saturating, NaN -> 0
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    dst[(size_t)y * dstPitch + x] = (uint8_t)q;
}

cudaError_t launch_quantise(cudaStream_t s, const float *src, size_t srcPitch, uint8_t *dst, size_t dstPitch, int rows, int cols)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(cols, block.x), rtdd_div_up(rows, block.y));
    quantise_kernel<<<grid, block, 0, s>>>(src, srcPitch, dst, dstPitch, rows, cols);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
fill_f32_kernel(float *__restrict__ dst, size_t pitch, int rows, int cols, float v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    ((float *)((char *)dst + (size_t)y * pitch))[x] = v;
}

cudaError_t launch_fill_f32(cudaStream_t s, float *dst, size_t pitch, int rows, int cols, float v)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(cols, block.x), rtdd_div_up(rows, block.y));
    fill_f32_kernel<<<grid, block, 0, s>>>(dst, pitch, rows, cols, v);
    return cudaGetLastError();
}

}  // namespace rtdd
