// cv::pyrUp (fp32) as device expressions, shared by the stand-alone prolongation kernels (image_kernels.cu) and the
// fused prolongation + Dirichlet injection + edge-weight pass (solver_kernels.cu).  No multiply-add is contracted
// (explicit __fmul_rn/__fadd_rn), so every user is bit-equal to OpenCV's unfused CPU path and to oracle_pyrup_f32.
// ref: src/main.cpp:272-279
#pragma once

namespace rtdd {

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

// destination pixel (dy, dx) of a (2n | 2n+1)-sized destination from an srows x scols source: the scalar expressions
__device__ __forceinline__ float pyrup_px(const float *__restrict__ src, size_t srcPitch, int srows, int scols, int dy, int dx)
{
    if (dy >= 2 * srows) dy = 2 * srows - 2;               // odd destination height: repeat row 2n-2
    const int y = dy >> 1;
    const int ym = (y > 0) ? y - 1 : (srows > 1 ? 1 : 0);
    const int yp = (y < srows - 1) ? y + 1 : srows - 1;
    const float *s1 = (const float *)((const char *)src + (size_t)y * srcPitch);
    const float *s2 = (const float *)((const char *)src + (size_t)yp * srcPitch);
    const float r1 = pyrup_h(s1, scols, dx);
    const float r2 = pyrup_h(s2, scols, dx);
    if (dy & 1) return __fmul_rn(__fadd_rn(r1, r2), 1.0f / 16.0f);
    const float *s0 = (const float *)((const char *)src + (size_t)ym * srcPitch);
    const float r0 = pyrup_h(s0, scols, dx);
    return __fmul_rn(__fadd_rn(__fadd_rn(r0, __fmul_rn(r1, 6.0f)), r2), 1.0f / 64.0f);
}

__device__ __forceinline__ void pyrup_h4(const float *__restrict__ s, int j, float (&r)[4])
{
    const float a = __ldg(s + 2 * j - 1), b = __ldg(s + 2 * j), c = __ldg(s + 2 * j + 1), d = __ldg(s + 2 * j + 2);
    r[0] = __fadd_rn(__fadd_rn(a, __fmul_rn(b, 6.0f)), c);
    r[1] = __fmul_rn(__fadd_rn(b, c), 4.0f);
    r[2] = __fadd_rn(__fadd_rn(b, __fmul_rn(c, 6.0f)), d);
    r[3] = __fmul_rn(__fadd_rn(c, d), 4.0f);
}

}  // namespace rtdd
