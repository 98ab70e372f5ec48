// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Same number of floating-point FMAs in both kernels; 8 independent accumulator chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, float a, float b, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

__global__ void k_ffma2(float* out, float a, float b, int iters) {
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = pack(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
    const unsigned long long aa = pack(a, a), bb = pack(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x[i]) : "l"(x[i]), "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)x[i]) + __uint_as_float((unsigned)(x[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    float* out;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        float ms1, ms2;
        cudaEventRecord(e0); k_ffma<<<blocks, threads>>>(out, 0.999f, 0.001f, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, 0.999f, 0.001f, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1);
        const double fmas = (double)blocks * threads * iters * 16;
        printf("rep %d: FFMA %.3f ms (%.1f TFMA/s)   FFMA2 %.3f ms (%.1f TFMA/s)   ratio %.2f\n", rep, ms1,
               fmas / ms1 * 1e-9, ms2, fmas / ms2 * 1e-9, ms1 / ms2);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
