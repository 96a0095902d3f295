// Microbenchmark (B200): what a persistent CTA pays for writing a 2 x 128 x 64 fp32 tile (64 KB) back to global memory when all 148
// CTAs do it at the same moment -- the write-back phase of the temporally blocked sweep kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o writeback_rate writeback_rate.cu && ./writeback_rate
// Every CTA loops over `regions`: `spin` dependent FMAs per thread (the sweeps), then the write-back in one of the forms
//   0  8 STG.128 per thread straight from registers
//   1  park the rows in shared memory, lanes 0..7 of every warp hand one 512-byte row each to the bulk-copy engine
//      (cp.async.bulk.global.shared::cta), the next region waits for the engine to have read them before parking again
//   2  as 1, but one elected lane per warp issues the 8 rows
//   3  as 1, with the wait placed after half of the next region's arithmetic (the prologue that does not touch the parking area)
// Reported: time per region minus the time of the same loop without any write-back, and the cycles the issuing lanes spend in the
// issue itself (clock64 around it, warp 0).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store_row(void *dst, unsigned int src, unsigned int bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst)), "r"(src), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *x, float *p, int pitch, int regions, int spin, long long *issueCycles, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a[4][4], b[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) { a[r][i] = (float)(threadIdx.x + r + i); b[r][i] = 0.5f * i; }
    const unsigned int base = smem_u32(smem);
    long long issued = 0;
    for (int it = 0; it < regions; it++) {
        const int tile = it * gridDim.x + blockIdx.x;
        const size_t row0 = (size_t)(tile % 4096) * 64 + warp * 4;          // a 128-column strip of a tall plane
        for (int s = 0; s < spin / 2; s++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) a[r][i] = __fmaf_rn(a[r][i], 1.0001f, b[r][i]);
        if (MODE == 3 && lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (MODE == 3) __syncwarp();
        for (int s = spin / 2; s < spin; s++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) a[r][i] = __fmaf_rn(a[r][i], 1.0001f, b[r][i]);
        __syncthreads();                                                     // the sweeps end with a CTA barrier
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                *(float4 *)(x + (row0 + r) * pitch + 4 * lane) = make_float4(a[r][0], a[r][1], a[r][2], a[r][3]);
                *(float4 *)(p + (row0 + r) * pitch + 4 * lane) = make_float4(b[r][0], b[r][1], b[r][2], b[r][3]);
            }
        } else if (MODE >= 1) {
            if (MODE != 3) {
                if (lane < 8) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                *(float4 *)(smem + ((warp * 4 + r) * 2 + 0) * 512 + lane * 16) = make_float4(a[r][0], a[r][1], a[r][2], a[r][3]);
                *(float4 *)(smem + ((warp * 4 + r) * 2 + 1) * 512 + lane * 16) = make_float4(b[r][0], b[r][1], b[r][2], b[r][3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const long long t0 = clock64();
            if (MODE == 2) {
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        bulk_store_row(((j & 1) ? p : x) + (row0 + (j >> 1)) * pitch, base + (unsigned int)((warp * 8 + j) * 512), 512u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (lane < 8) {
                bulk_store_row(((lane & 1) ? p : x) + (row0 + (lane >> 1)) * pitch, base + (unsigned int)((warp * 8 + lane) * 512), 512u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (warp == 0 && lane == 0) issued += clock64() - t0;
        }
    }
    if (MODE >= 1 && lane < 8) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) issueCycles[0] = issued;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc += a[r][i];
    if (acc == 12345.f) sink[0] = acc;
}

__global__ void __launch_bounds__(512, 1) k_none(int regions, int spin, float *sink)
{
    float a[4][4], b[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) { a[r][i] = (float)(threadIdx.x + r + i); b[r][i] = 0.5f * i; }
    for (int it = 0; it < regions; it++) {
        for (int s = 0; s < spin; s++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) a[r][i] = __fmaf_rn(a[r][i], 1.0001f, b[r][i]);
        __syncthreads();
    }
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc += a[r][i];
    if (acc == 12345.f) sink[0] = acc;
}

template <class F> static float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    const int pitch = 128, regions = 64;
    const size_t rows = 4096ull * 64 + 64;
    float *x, *p, *sink; long long *cyc;
    cudaMalloc(&x, rows * pitch * 4); cudaMalloc(&p, rows * pitch * 4); cudaMalloc(&sink, 64); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int spin : {0, 100, 300}) {
        const float t0 = timeit([&] { k_none<<<148, 512>>>(regions, spin, sink); });
        printf("spin %4d: arithmetic only %.2f us per region\n", spin, t0 * 1e3f / regions);
        const char *names[4] = {"STG.128 from registers         ", "bulk rows, lanes 0..7          ", "bulk rows, one lane            ", "bulk rows, wait half a region later"};
        for (int mode = 0; mode < 4; mode++) {
            float t;
            if (mode == 0) t = timeit([&] { k<0><<<148, 512, 0>>>(x, p, pitch, regions, spin, cyc, sink); });
            else if (mode == 1) t = timeit([&] { k<1><<<148, 512, 65536>>>(x, p, pitch, regions, spin, cyc, sink); });
            else if (mode == 2) t = timeit([&] { k<2><<<148, 512, 65536>>>(x, p, pitch, regions, spin, cyc, sink); });
            else t = timeit([&] { k<3><<<148, 512, 65536>>>(x, p, pitch, regions, spin, cyc, sink); });
            long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("   %s  +%.2f us per region (%.0f GB/s aggregate while writing), issue %.0f cycles per region\n", names[mode],
                   (t - t0) * 1e3f / regions, 148.0 * 65536 / ((t - t0) * 1e-3 / regions) * 1e-9, mode ? (double)c / regions : 0.0);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
