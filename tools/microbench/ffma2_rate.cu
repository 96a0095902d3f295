// Microbenchmark (B200): issue rate of scalar FFMA (three distinct register operands) against packed FFMA2 (fma.rn.f32x2),
// FADD vs FADD2, at 4 warps per SM sub-partition (the occupancy of the blocked sweep kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *out, int iters, float s)
{
    float a[16], b[16], c[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { a[j] = s + j + threadIdx.x; b[j] = s * 0.5f + j; c[j] = (float)j; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; j++) c[j] = __fmaf_rn(a[j], b[j], c[j]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 r = __ffma2_rn(make_float2(a[j], a[j + 1]), make_float2(b[j], b[j + 1]), make_float2(c[j], c[j + 1]));
                c[j] = r.x; c[j + 1] = r.y;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; j++) c[j] = __fadd_rn(a[j], c[j]);
        } else if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 r = __fadd2_rn(make_float2(a[j], a[j + 1]), make_float2(c[j], c[j + 1]));
                c[j] = r.x; c[j + 1] = r.y;
            }
        } else if (MODE == 4) {          // FFMA chain mixed with min/max (alu pipe), like the sweep
#pragma unroll
            for (int j = 0; j < 16; j++) c[j] = fminf(fmaxf(__fmaf_rn(a[j], b[j], c[j]), 0.0f), 255.0f);
        } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 r = __ffma2_rn(make_float2(a[j], a[j + 1]), make_float2(b[j], b[j + 1]), make_float2(c[j], c[j + 1]));
                c[j] = fminf(fmaxf(r.x, 0.0f), 255.0f); c[j + 1] = fminf(fmaxf(r.y, 0.0f), 255.0f);
            }
        }
    }
    float t = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) t += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char *name, float *out)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(out, 100, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // 16 scalar results per thread per iteration; 4 warps per sub-partition
    const double warpOpsPerSmsp = 4.0 * 16.0 * iters;
    printf("%-28s %.3f ms  -> %.2f cycles per warp-wide scalar result per sub-partition (at 1.965 GHz)\n", name, ms, ms * 1e-3 * 1.965e9 / warpOpsPerSmsp);
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    run<0>("FFMA  (3 regs)", out);
    run<1>("FFMA2 (packed)", out);
    run<2>("FADD", out);
    run<3>("FADD2 (packed)", out);
    run<4>("FFMA + FMNMX x2", out);
    run<5>("FFMA2 + FMNMX x2", out);
    cudaFree(out);
    return 0;
}
