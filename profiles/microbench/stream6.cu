// Micro-benchmark: how fast can one B200 stream the record columns of the streaming unary kernel?
// Six columns of 4-byte elements (slot offsets, c0, l0, a0, W_f, gamma), N records, read once, a few
// FMAs per record -- the memory side of unary_fold_kernel without its run logic.  Variants:
//   ldg        plain 16-byte column loads, bytes in flight by occupancy (what the kernel does today)
//   ldg2       the same with the next tile's quads prefetched into a second set of registers
//   cpasync<S> per-thread cp.async ring in shared memory, S stages, no block-level synchronisation
//   bulk<S>    cp.async.bulk (TMA 1-D) + mbarrier ring, one elected producer thread per block
//   aos_*      the same bytes laid out tile by tile (24 KB contiguous per 1024 records)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream6 stream6.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int kTile = 1024, kThreads = 256, kQuad = 4;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void consume(float& s0, float& s1, float& s2, const float4 o, const float4 c, const float4 l,
                                        const float4 a, const float4 w, const float4 g) {
    s0 += g.x * (c.x + o.x * l.x) + g.y * (c.y + o.y * l.y) + g.z * (c.z + o.z * l.z) + g.w * (c.w + o.w * l.w);
    s1 += w.x * (l.x + a.x) + w.y * (l.y + a.y) + w.z * (l.z + a.z) + w.w * (l.w + a.w);
    s2 += g.x * a.x + g.y * a.y + g.z * a.z + g.w * a.w;
}

__device__ __forceinline__ void finish(float s0, float s1, float s2, float* out) {
    float s = s0 + s1 + s2;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// ---- column-major, LDG ----------------------------------------------------------------------------
template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM)
k_ldg(const float* __restrict__ col, long long n, long long tiles, float* out) {
    const unsigned lo = (unsigned)(tiles * blockIdx.x / gridDim.x) * kTile, hi = (unsigned)(tiles * (blockIdx.x + 1) / gridDim.x) * kTile;
    float s0 = 0, s1 = 0, s2 = 0;
    for (unsigned r = lo + threadIdx.x * kQuad; r < hi; r += kTile)
        consume(s0, s1, s2, ldg4(col + r), ldg4(col + n + r), ldg4(col + 2 * n + r), ldg4(col + 3 * n + r),
                ldg4(col + 4 * n + r), ldg4(col + 5 * n + r));
    finish(s0, s1, s2, out);
}

template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM)
k_ldg2(const float* __restrict__ col, long long n, long long tiles, float* out) {
    const unsigned lo = (unsigned)(tiles * blockIdx.x / gridDim.x) * kTile, hi = (unsigned)(tiles * (blockIdx.x + 1) / gridDim.x) * kTile;
    float s0 = 0, s1 = 0, s2 = 0;
    unsigned r = lo + threadIdx.x * kQuad;
    if (r >= hi) { finish(0, 0, 0, out); return; }
    float4 o = ldg4(col + r), c = ldg4(col + n + r), l = ldg4(col + 2 * n + r), a = ldg4(col + 3 * n + r),
           w = ldg4(col + 4 * n + r), g = ldg4(col + 5 * n + r);
    for (; r < hi; r += kTile) {
        const unsigned rn = r + kTile < hi ? r + kTile : r;
        const float4 o2 = ldg4(col + rn), c2 = ldg4(col + n + rn), l2 = ldg4(col + 2 * n + rn), a2 = ldg4(col + 3 * n + rn),
                     w2 = ldg4(col + 4 * n + rn), g2 = ldg4(col + 5 * n + rn);
        consume(s0, s1, s2, o, c, l, a, w, g);
        o = o2; c = c2; l = l2; a = a2; w = w2; g = g2;
    }
    finish(s0, s1, s2, out);
}

// ---- column-major, per-thread cp.async ring ---------------------------------------------------------
template <int S, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM)
k_cpasync(const float* __restrict__ col, long long n, long long tiles, float* out) {
    extern __shared__ __align__(16) float4 ring[];      // [S][6][kThreads]
    const unsigned lo = (unsigned)(tiles * blockIdx.x / gridDim.x) * kTile, hi = (unsigned)(tiles * (blockIdx.x + 1) / gridDim.x) * kTile;
    const int n_tiles = (int)((hi - lo) / kTile);
    auto issue = [&](int t) {
        if (t < n_tiles) {
            const unsigned r = lo + (unsigned)t * kTile + threadIdx.x * kQuad;
            float4* dst = ring + (size_t)(t % S) * 6 * kThreads + threadIdx.x;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const unsigned d = (unsigned)__cvta_generic_to_shared(dst + c * kThreads);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(col + (long long)c * n + r) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < S - 1; ++t) issue(t);
    float s0 = 0, s1 = 0, s2 = 0;
    for (int t = 0; t < n_tiles; ++t) {
        issue(t + S - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        const float4* src = ring + (size_t)(t % S) * 6 * kThreads + threadIdx.x;
        consume(s0, s1, s2, src[0], src[kThreads], src[2 * kThreads], src[3 * kThreads], src[4 * kThreads], src[5 * kThreads]);
    }
    finish(s0, s1, s2, out);
}

// ---- cp.async.bulk + mbarrier ring -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// AOS = true: one 24 KB copy per tile from the tile-major layout; else six 4 KB column copies.
// Consumers release a stage with an mbarrier arrive (count = warps), no __syncthreads in the loop.
template <int S, int BLOCKS_PER_SM, bool AOS>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM)
k_bulk(const float* __restrict__ col, long long n, long long tiles, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* ring = reinterpret_cast<float4*>(smem);                       // [S][6][kThreads]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)S * 6 * kThreads * 16);
    unsigned long long* empty = full + S;
    const unsigned lo = (unsigned)(tiles * blockIdx.x / gridDim.x), hi = (unsigned)(tiles * (blockIdx.x + 1) / gridDim.x);
    const int n_tiles = (int)(hi - lo);
    constexpr unsigned kColBytes = kTile * 4u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(full + s)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(empty + s)), "r"(kThreads / 32) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % S;
        const long long tile = lo + t;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(full + s)), "r"(6u * kColBytes) : "memory");
        if (AOS) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(ring + (size_t)s * 6 * kThreads)), "l"(col + tile * 6 * kTile), "r"(6u * kColBytes), "r"(smem_addr(full + s)) : "memory");
        } else {
#pragma unroll
            for (int c = 0; c < 6; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr(ring + ((size_t)s * 6 + c) * kThreads)), "l"(col + (long long)c * n + tile * kTile), "r"(kColBytes), "r"(smem_addr(full + s)) : "memory");
        }
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < S && t < n_tiles; ++t) issue(t);
    float s0 = 0, s1 = 0, s2 = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % S;
        mbar_wait(full + s, (unsigned)((t / S) & 1));
        const float4* src = ring + (size_t)s * 6 * kThreads + threadIdx.x;
        const float4 o = src[0], c = src[kThreads], l = src[2 * kThreads], a = src[3 * kThreads], w = src[4 * kThreads], g = src[5 * kThreads];
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(empty + s)) : "memory");
        if (threadIdx.x == 0 && t + S < n_tiles) {
            mbar_wait(empty + s, (unsigned)((t / S) & 1));
            issue(t + S);
        }
        consume(s0, s1, s2, o, c, l, a, w, g);
    }
    finish(s0, s1, s2, out);
}

// ---- tile-major (AoS of column tiles), LDG ----------------------------------------------------------
template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(kThreads, BLOCKS_PER_SM)
k_aos_ldg(const float* __restrict__ col, long long n, long long tiles, float* out) {
    const unsigned lo = (unsigned)(tiles * blockIdx.x / gridDim.x), hi = (unsigned)(tiles * (blockIdx.x + 1) / gridDim.x);
    float s0 = 0, s1 = 0, s2 = 0;
    for (unsigned t = lo; t < hi; ++t) {
        const float* p = col + (long long)t * 6 * kTile + threadIdx.x * kQuad;
        consume(s0, s1, s2, ldg4(p), ldg4(p + kTile), ldg4(p + 2 * kTile), ldg4(p + 3 * kTile), ldg4(p + 4 * kTile), ldg4(p + 5 * kTile));
    }
    finish(s0, s1, s2, out);
}

// grid-stride over tiles with many blocks (hardware block scheduler balances the SMs)
__global__ void __launch_bounds__(kThreads, 8)
k_ldg_waves(const float* __restrict__ col, long long n, long long tiles, float* out) {
    float s0 = 0, s1 = 0, s2 = 0;
    const unsigned r = blockIdx.x * kTile + threadIdx.x * kQuad;
    consume(s0, s1, s2, ldg4(col + r), ldg4(col + n + r), ldg4(col + 2 * n + r), ldg4(col + 3 * n + r),
            ldg4(col + 4 * n + r), ldg4(col + 5 * n + r));
    finish(s0, s1, s2, out);
}

template <typename F>
static void run(const char* name, long long n, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch(i & 1);
    cudaDeviceSynchronize();
    float best = 1e9f, sum = 0.f;
    const int reps = 20;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(e0);
        launch(i & 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
        sum += ms;
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = 24.0 * n;
    printf("%-28s mean %7.2f us  best %7.2f us  %7.1f GB/s (mean)  %7.1f GB/s (best)  %s\n", name, sum / reps * 1e3, best * 1e3,
           bytes / (sum / reps * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
    const long long n = (argc > 1 ? atoll(argv[1]) : 7000000ll) / kTile * kTile;
    const long long tiles = n / kTile;
    float* buf[2];
    float* out;
    for (int i = 0; i < 2; ++i) {                 // two copies, alternated: a repeat never finds its bytes in L2
        cudaMalloc(&buf[i], 24 * n);
        std::vector<float> h(6 * n);
        for (long long j = 0; j < 6 * n; ++j) h[j] = (float)((j * 2654435761u) & 1023) * 1e-3f;
        cudaMemcpy(buf[i], h.data(), 24 * n, cudaMemcpyHostToDevice);
    }
    cudaMalloc(&out, 4);
    cudaMemset(out, 0, 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("records %lld (%.1f MB), %d SMs\n", n, 24.0 * n / 1e6, sms);

    run("ldg 4 blocks/SM", n, [&](int b) { k_ldg<4><<<sms * 4, kThreads>>>(buf[b], n, tiles, out); });
    run("ldg 8 blocks/SM", n, [&](int b) { k_ldg<8><<<sms * 8, kThreads>>>(buf[b], n, tiles, out); });
    run("ldg 4/SM, 4 waves", n, [&](int b) { k_ldg<4><<<sms * 16, kThreads>>>(buf[b], n, tiles, out); });
    run("ldg one tile per block", n, [&](int b) { k_ldg_waves<<<(unsigned)tiles, kThreads>>>(buf[b], n, tiles, out); });
    run("ldg2 (reg prefetch) 3/SM", n, [&](int b) { k_ldg2<3><<<sms * 3, kThreads>>>(buf[b], n, tiles, out); });
    run("ldg2 (reg prefetch) 4/SM", n, [&](int b) { k_ldg2<4><<<sms * 4, kThreads>>>(buf[b], n, tiles, out); });
    run("aos ldg 4/SM", n, [&](int b) { k_aos_ldg<4><<<sms * 4, kThreads>>>(buf[b], n, tiles, out); });
    run("aos ldg 8/SM", n, [&](int b) { k_aos_ldg<8><<<sms * 8, kThreads>>>(buf[b], n, tiles, out); });
#define CPA(S, B)                                                                                              \
    {                                                                                                          \
        const size_t sm = (size_t)S * 6 * kThreads * 16;                                                       \
        cudaFuncSetAttribute(k_cpasync<S, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);           \
        run("cpasync S=" #S " " #B "/SM", n, [&](int b) { k_cpasync<S, B><<<sms * B, kThreads, sm>>>(buf[b], n, tiles, out); }); \
    }
    CPA(2, 4) CPA(3, 3) CPA(4, 2) CPA(2, 3) CPA(3, 2)
#define BULK(S, B, A)                                                                                          \
    {                                                                                                          \
        const size_t sm = (size_t)S * 6 * kThreads * 16 + 2 * S * 8;                                           \
        cudaFuncSetAttribute(k_bulk<S, B, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);           \
        run(A ? "bulk aos S=" #S " " #B "/SM" : "bulk cols S=" #S " " #B "/SM", n,                               \
            [&](int b) { k_bulk<S, B, A><<<sms * B, kThreads, sm>>>(buf[b], n, tiles, out); });                  \
    }
    BULK(2, 4, false) BULK(4, 2, false) BULK(3, 3, false) BULK(2, 4, true) BULK(4, 2, true) BULK(3, 3, true) BULK(8, 1, true)
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
