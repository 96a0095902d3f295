// GPUImageProcessing kernels (Dirichlet injection, annotation restriction, brush)
// and the pyramid ops that sit either side of the solve (BGR->gray, gray pyrDown,
// depth pyrUp, u8 quantiser).
//
// ref: src/GPUImageProcessing.cu:8-100; src/main.cpp:111-112,143-145,272-279,290.

#include "rtdd_internal.h"
#include "pyrup_device.h"

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

// 4 pixels per thread: one 32-bit mask word decides; untouched groups (the common case) cost 1 byte per pixel
__global__ void __launch_bounds__(256)
convert4_kernel(const uint8_t *__restrict__ src, size_t srcPitch, float *__restrict__ dst, size_t dstPitch,
                const uint8_t *__restrict__ mask, size_t maskPitch, int rows, int cols)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");               // programmatic dependent launch: predecessor complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= cols || y >= rows) return;
    float *dRow = (float *)((char *)dst + (size_t)y * dstPitch) + x4;
    const uint8_t *sRow = src + (size_t)y * srcPitch + 3 * x4;
    if (x4 + 4 <= cols) {
        const unsigned int m = __ldg((const unsigned int *)(mask + (size_t)y * maskPitch + x4));
        if (!((m & 0xFFu) == 0xFFu || ((m >> 8) & 0xFFu) == 0xFFu || ((m >> 16) & 0xFFu) == 0xFFu || (m >> 24) == 0xFFu)) return;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (((m >> (8 * i)) & 0xFFu) == 0xFFu) dRow[i] = (float)__ldg(sRow + 3 * i);
    } else {
        for (int i = 0; x4 + i < cols; i++)
            if (__ldg(mask + (size_t)y * maskPitch + x4 + i) == 255) dRow[i] = (float)__ldg(sRow + 3 * i);
    }
}

cudaError_t launch_convert(cudaStream_t s, const uint8_t *src, size_t srcPitch, float *dst, size_t dstPitch,
                           const uint8_t *mask, size_t maskPitch, int rows, int cols)
{
    if ((((uintptr_t)mask | maskPitch) & 3u) == 0) {
        dim3 block(64, 4);
        dim3 grid(rtdd_div_up(rtdd_div_up(cols, 4), block.x), rtdd_div_up(rows, block.y));
        return launch_pdl(convert4_kernel, grid, block, (size_t)0, s, src, srcPitch, dst, dstPitch, mask, maskPitch, rows, cols);
    }
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(cols, block.x), rtdd_div_up(rows, block.y));
    convert_kernel<<<grid, block, 0, s>>>(src, srcPitch, dst, dstPitch, mask, maskPitch, rows, cols);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// annotation ingest: the reference's persistent annotation format is ONE gray plane, 32 = "not annotated"
// (the -a file, ref: src/main.cpp:160-170):  every pixel != 32 gets edited BGR := that value and scribble := 255;
// all other pixels keep the original image in `edited` (src/main.cpp:158) and 0 in `scribble` (:132).
// One thread = 4 pixels: 4 B of annotation + 12 B of image in, 12 B + 4 B out.  A frame that arrives from the host needs
// only this plane (1 B/px) instead of the scribble + 3-channel edited planes (4 B/px) main.cpp uploads (:236-237).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
annotation_ingest_kernel(const uint8_t *__restrict__ ann, size_t annPitch, const uint8_t *__restrict__ bgr, size_t bgrPitch,
                         uint8_t *__restrict__ edited, size_t editedPitch, uint8_t *__restrict__ scribble, size_t scribblePitch,
                         int rows, int cols, int vec)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= cols || y >= rows) return;
    const uint8_t *aRow = ann + (size_t)y * annPitch + x4;
    const uint8_t *bRow = bgr + (size_t)y * bgrPitch + 3 * x4;
    uint8_t *eRow = edited + (size_t)y * editedPitch + 3 * x4;
    uint8_t *sRow = scribble + (size_t)y * scribblePitch + x4;
    if (vec && x4 + 4 <= cols) {
        const unsigned int a = __ldg((const unsigned int *)aRow);
        unsigned int w[3] = {__ldg((const unsigned int *)bRow), __ldg((const unsigned int *)bRow + 1), __ldg((const unsigned int *)bRow + 2)};
        unsigned int m = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned int v = (a >> (8 * i)) & 0xFFu;
            if (v != 32u) {
                m |= 0xFFu << (8 * i);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int b = 3 * i + c;                       // byte index inside the 12-byte group
                    w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | (v << (8 * (b & 3)));
                }
            }
        }
        ((unsigned int *)eRow)[0] = w[0]; ((unsigned int *)eRow)[1] = w[1]; ((unsigned int *)eRow)[2] = w[2];
        *(unsigned int *)sRow = m;
    } else {
        for (int i = 0; x4 + i < cols; i++) {
            const unsigned int v = __ldg(aRow + i);
            const bool on = (v != 32u);
            for (int c = 0; c < 3; c++) eRow[3 * i + c] = on ? (uint8_t)v : __ldg(bRow + 3 * i + c);
            sRow[i] = on ? 255 : 0;
        }
    }
}

cudaError_t launch_annotation_ingest(cudaStream_t s, const uint8_t *ann, size_t annPitch, const uint8_t *bgr, size_t bgrPitch,
                                     uint8_t *edited, size_t editedPitch, uint8_t *scribble, size_t scribblePitch, int rows, int cols)
{
    const int vec = ((((uintptr_t)ann | annPitch | (uintptr_t)bgr | bgrPitch | (uintptr_t)edited | editedPitch | (uintptr_t)scribble | scribblePitch) & 3u) == 0) ? 1 : 0;
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(rtdd_div_up(cols, 4), block.x), rtdd_div_up(rows, block.y));
    return launch_pdl(annotation_ingest_kernel, grid, block, (size_t)0, s, ann, annPitch, bgr, bgrPitch, edited, editedPitch, scribble, scribblePitch,
                      rows, cols, vec);
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

// 4 output pixels per thread: the 2 x 9 mask bytes they look at are fetched as two 64-bit words (+1 byte each);
// groups without any scribble (the common case) leave after that
__global__ void __launch_bounds__(256)
pyrdown_annotation4_kernel(const uint8_t *__restrict__ prevScribble, size_t prevScribblePitch,
                           const uint8_t *__restrict__ prevEdited, size_t prevEditedPitch, int previousRows, int previousCols,
                           uint8_t *__restrict__ currScribble, size_t currScribblePitch,
                           uint8_t *__restrict__ currEdited, size_t currEditedPitch, int currentRows, int currentCols)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");               // programmatic dependent launch: predecessor complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= currentCols || y >= currentRows) return;
    const int px0 = 2 * x4;                       // window columns px0-1 .. px0+6
    const bool fast = (px0 + 8 <= previousCols) && (x4 + 4 <= currentCols);
    if (fast) {
        unsigned long long w[2] = {0ull, 0ull};
        unsigned int before[2] = {0u, 0u};
        bool any = false;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int py = 2 * y - 1 + j;
            if (py < 0 || py >= previousRows) continue;
            const uint8_t *row = prevScribble + (size_t)py * prevScribblePitch;
            w[j] = __ldg((const unsigned long long *)(row + px0));
            if (px0 > 0) before[j] = __ldg(row + px0 - 1);
            // any byte equal to 0xFF?  (classic has-zero-byte test on the complement)
            const unsigned long long v = ~w[j];
            any = any || (((v - 0x0101010101010101ull) & ~v & 0x8080808080808080ull) != 0ull) || before[j] == 255u;
        }
        if (!any) return;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int hitX = -1, hitY = -1;
#pragma unroll
            for (int dy = 0; dy >= -1; dy--) {
#pragma unroll
                for (int dx = 0; dx >= -1; dx--) {
                    const int c = 2 * i + dx;             // column relative to px0
                    const int j = dy + 1;
                    const unsigned int b = (c < 0) ? before[j] : (unsigned int)((w[j] >> (8 * c)) & 0xFFull);
                    const int py = 2 * y + dy, px = px0 + c;
                    if (hitX < 0 && px >= 0 && py >= 0 && py < previousRows && b == 255u) { hitX = px; hitY = py; }
                }
            }
            if (hitX >= 0) {
                currScribble[(size_t)y * currScribblePitch + x4 + i] = 255;
                currEdited[(size_t)y * currEditedPitch + 3 * (x4 + i)] = __ldg(prevEdited + (size_t)hitY * prevEditedPitch + 3 * hitX);
            }
        }
        return;
    }
    for (int i = 0; i < 4 && x4 + i < currentCols; i++) {
        const int x = x4 + i;
        int hitX = -1, hitY = -1;
        for (int dy = 0; dy >= -1; dy--)
            for (int dx = 0; dx >= -1; dx--) {
                const int px = 2 * x + dx, py = 2 * y + dy;
                if (hitX < 0 && px >= 0 && py >= 0 && px < previousCols && py < previousRows &&
                    __ldg(prevScribble + (size_t)py * prevScribblePitch + px) == 255) { hitX = px; hitY = py; }
            }
        if (hitX >= 0) {
            currScribble[(size_t)y * currScribblePitch + x] = 255;
            currEdited[(size_t)y * currEditedPitch + 3 * x] = __ldg(prevEdited + (size_t)hitY * prevEditedPitch + 3 * hitX);
        }
    }
}

cudaError_t launch_pyrdown_annotation(cudaStream_t s, const uint8_t *prevScribble, size_t prevScribblePitch,
                                      const uint8_t *prevEdited, size_t prevEditedPitch, int previousRows, int previousCols,
                                      uint8_t *currScribble, size_t currScribblePitch, uint8_t *currEdited, size_t currEditedPitch,
                                      int currentRows, int currentCols)
{
    if ((((uintptr_t)prevScribble | prevScribblePitch) & 7u) == 0) {
        dim3 block(64, 4);
        dim3 grid(rtdd_div_up(rtdd_div_up(currentCols, 4), block.x), rtdd_div_up(currentRows, block.y));
        return launch_pdl(pyrdown_annotation4_kernel, grid, block, (size_t)0, s, prevScribble, prevScribblePitch, prevEdited, prevEditedPitch,
                          previousRows, previousCols, currScribble, currScribblePitch, currEdited, currEditedPitch, currentRows, currentCols);
    }
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

// Interior fast path: one thread produces destination columns 4j..4j+3 of the row pair (2y, 2y+1) from source
// columns 2j-1..2j+2 of source rows y-1, y, y+1 -- the same expressions, in the same order, as pyrup_h and the scalar
// kernel (bit-identical); border columns / rows and odd-sized destinations go through the scalar expressions.
__global__ void __launch_bounds__(256)
pyrup_depth4_kernel(const float *__restrict__ src, size_t srcPitch, int srows, int scols,
                    float *__restrict__ dst, size_t dstPitch, int drows, int dcols, int rowBegin, int rowEnd)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");               // programmatic dependent launch: predecessor complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int j = blockIdx.x * blockDim.x + threadIdx.x;            // destination columns 4j..4j+3
    const int y = (rowBegin >> 1) + blockIdx.y * blockDim.y + threadIdx.y;   // source row; destination rows 2y, 2y+1
    const int dx0 = 4 * j;
    if (dx0 >= dcols || 2 * y >= rowEnd || y >= srows) return;
    const bool colFast = (2 * j - 1 >= 0) && (2 * j + 2 <= scols - 1) && (dx0 + 4 <= dcols);
    const bool rowFast = (y >= 1) && (y + 1 <= srows - 1);
    if (colFast && rowFast) {
        float r0[4], r1[4], r2[4];
        pyrup_h4((const float *)((const char *)src + (size_t)(y - 1) * srcPitch), j, r0);
        pyrup_h4((const float *)((const char *)src + (size_t)y * srcPitch), j, r1);
        pyrup_h4((const float *)((const char *)src + (size_t)(y + 1) * srcPitch), j, r2);
        float e[4], o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            e[i] = __fmul_rn(__fadd_rn(__fadd_rn(r0[i], __fmul_rn(r1[i], 6.0f)), r2[i]), 1.0f / 64.0f);
            o[i] = __fmul_rn(__fadd_rn(r1[i], r2[i]), 1.0f / 16.0f);
        }
        if (2 * y >= rowBegin) *(float4 *)((float *)((char *)dst + (size_t)(2 * y) * dstPitch) + dx0) = make_float4(e[0], e[1], e[2], e[3]);
        if (2 * y + 1 < rowEnd) *(float4 *)((float *)((char *)dst + (size_t)(2 * y + 1) * dstPitch) + dx0) = make_float4(o[0], o[1], o[2], o[3]);
        return;
    }
    // borders: the scalar expressions, pixel by pixel (also the extra row / column of an odd-sized destination)
    const int yLast = (y == srows - 1) ? drows - 1 : 2 * y + 1;    // the last source row also produces the odd extra row
    for (int dy = 2 * y; dy <= yLast; dy++) {
        if (dy < rowBegin || dy >= rowEnd) continue;
        int sy2 = dy;
        if (sy2 >= 2 * srows) sy2 = 2 * srows - 2;
        const int sy = sy2 >> 1;
        const int ym = (sy > 0) ? sy - 1 : (srows > 1 ? 1 : 0);
        const int yp = (sy < srows - 1) ? sy + 1 : srows - 1;
        const float *s0 = (const float *)((const char *)src + (size_t)ym * srcPitch);
        const float *s1 = (const float *)((const char *)src + (size_t)sy * srcPitch);
        const float *s2 = (const float *)((const char *)src + (size_t)yp * srcPitch);
        const int dxEnd = (dx0 + 4 < dcols) ? dx0 + 4 : dcols;
        for (int dx = dx0; dx < dxEnd; dx++) {
            float *out = (float *)((char *)dst + (size_t)dy * dstPitch) + dx;
            const float r1 = pyrup_h(s1, scols, dx);
            const float r2 = pyrup_h(s2, scols, dx);
            if (sy2 & 1) {
                *out = __fmul_rn(__fadd_rn(r1, r2), 1.0f / 16.0f);
            } else {
                const float r0 = pyrup_h(s0, scols, dx);
                *out = __fmul_rn(__fadd_rn(__fadd_rn(r0, __fmul_rn(r1, 6.0f)), r2), 1.0f / 64.0f);
            }
        }
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
    if ((((uintptr_t)dst | dstPitch) & 15u) == 0 && srows >= 2 && scols >= 2) {
        // source rows rowBegin/2 .. (rowEnd-1)/2 (the last source row also covers an odd destination's extra row)
        int y0 = rowBegin >> 1, y1 = (rowEnd - 1) >> 1;
        if (y1 > srows - 1) y1 = srows - 1;
        dim3 block(32, 8);
        dim3 grid(rtdd_div_up(rtdd_div_up(dcols, 4), block.x), rtdd_div_up(y1 - y0 + 1, block.y));
        return launch_pdl(pyrup_depth4_kernel, grid, block, (size_t)0, s, src, srcPitch, srows, scols, dst, dstPitch, drows, dcols, rowBegin, rowEnd);
    }
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(dcols, block.x), rtdd_div_up(rowEnd - rowBegin, block.y));
    pyrup_depth_kernel<<<grid, block, 0, s>>>(src, srcPitch, srows, scols, dst, dstPitch, rowEnd, dcols, rowBegin);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Extension, NOT on the parity path (rtdd_frame_solve_band): the guess of a level when only a band of rows is re-solved.
// Inside the band: cv::pyrUp of the coarser level's NEW solution, exactly the parity guess.  Outside: the previous solution
// plus the prolongated CHANGE of the coarser level, pyrUp(new) - pyrUp(old) -- the far-field effect of the edit without
// re-solving those rows.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
band_prolong_kernel(const float *__restrict__ newC, const float *__restrict__ oldC, size_t pitchC, int srows, int scols,
                    float *__restrict__ dst, size_t dstPitch, int drows, int dcols, int bandBegin, int bandEnd)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dcols || y >= drows) return;
    float *out = (float *)((char *)dst + (size_t)y * dstPitch) + x;
    const float up = pyrup_px(newC, pitchC, srows, scols, y, x);
    if (y >= bandBegin && y < bandEnd) *out = up;
    else *out = __fadd_rn(*out, __fsub_rn(up, pyrup_px(oldC, pitchC, srows, scols, y, x)));
}

cudaError_t launch_band_prolong(cudaStream_t s, const float *newC, const float *oldC, size_t pitchC, int srows, int scols,
                                float *dst, size_t dstPitch, int drows, int dcols, int bandBegin, int bandEnd)
{
    dim3 block(64, 4);
    dim3 grid(rtdd_div_up(dcols, block.x), rtdd_div_up(drows, block.y));
    band_prolong_kernel<<<grid, block, 0, s>>>(newC, oldC, pitchC, srows, scols, dst, dstPitch, drows, dcols, bandBegin, bandEnd);
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

__global__ void __launch_bounds__(256)
quantise8_kernel(const float *__restrict__ src, size_t srcPitch, uint8_t *__restrict__ dst, size_t dstPitch, int rows, int cols)
{
    const int x8 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x8 >= cols || y >= rows) return;
    const float *sRow = (const float *)((const char *)src + (size_t)y * srcPitch) + x8;
    uint8_t *dRow = dst + (size_t)y * dstPitch + x8;
    if (x8 + 8 <= cols) {
        const float4 a = __ldg((const float4 *)sRow), b = __ldg((const float4 *)sRow + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned int w[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int q = __float2int_rn(v[i]);
            q = q < 0 ? 0 : (q > 255 ? 255 : q);
            w[i >> 2] |= (unsigned int)q << (8 * (i & 3));
        }
        *(uint2 *)dRow = make_uint2(w[0], w[1]);
    } else {
        for (int i = 0; x8 + i < cols; i++) {
            int q = __float2int_rn(__ldg(sRow + i));
            dRow[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
    }
}

cudaError_t launch_quantise(cudaStream_t s, const float *src, size_t srcPitch, uint8_t *dst, size_t dstPitch, int rows, int cols)
{
    if ((((uintptr_t)src | srcPitch) & 15u) == 0 && (((uintptr_t)dst | dstPitch) & 7u) == 0) {
        dim3 block(64, 4);
        dim3 grid(rtdd_div_up(rtdd_div_up(cols, 8), block.x), rtdd_div_up(rows, block.y));
        quantise8_kernel<<<grid, block, 0, s>>>(src, srcPitch, dst, dstPitch, rows, cols);
        return cudaGetLastError();
    }
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
