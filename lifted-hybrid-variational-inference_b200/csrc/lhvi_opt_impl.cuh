// Device code of the elementwise kernels around the factor pass -- ELBO / G_w reduction, cross-GPU
// exchange, optimiser step, batched queries -- shared by lhvi_opt.cu (one launch each) and the
// persistent iteration kernel (lhvi_iter_impl.cuh), which runs reduction + exchange + step between
// two grid barriers of the same launch.
#pragma once
#include "lhvi_common.cuh"

namespace lhvi {

// ---- G_w / energy: deterministic two-stage sum of the per-block partial rows ---------------

// partials is a sequence of regions of LHVI_PARTIAL_ROWS rows; row 0 of a region holds the number
// of valid data rows that follow (written by the factor kernel that owns the region).
template <typename real>
__global__ void __launch_bounds__(1024)
elbo_reduce_kernel(const double* __restrict__ partials, long long regions, int K, real* __restrict__ out) {
    __shared__ double s[32 * (LHVI_MAX_K + 1)];
    __shared__ double res[LHVI_MAX_K + 1];
    double acc[LHVI_MAX_K + 1];
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;
    const int W = K + 1;
    for (long long reg = 0; reg < regions; ++reg) {
        const double* base = partials + reg * LHVI_PARTIAL_ROWS * W;
        const int valid = (int)base[0];
        for (int r = threadIdx.x; r < valid; r += blockDim.x)
            for (int i = 0; i < W; ++i) acc[i] += base[(long long)(1 + r) * W + i];
    }
    block_sum_to(acc, W, s, res);
    if (threadIdx.x <= K) out[threadIdx.x] = (real)res[threadIdx.x];
}

// ---- step counter and Adam bias corrections, kept on the device so an iteration is
//      replayable from a CUDA graph (VarInference.py:253,272-273) ---------------------------

static __global__ void step_tick_kernel(double* step, double b1, double b2) {
    const double t = step[0] + 1.0;
    step[0] = t;
    step[1] = 1.0 - pow(b1, t);
    step[2] = 1.0 - pow(b2, t);
}

// ---- finish: partial-row reduction + step tick + cross-GPU exchange in one launch -----------
//
// Block 0 does what elbo_reduce_kernel and step_tick_kernel do.  With world > 1 every block then
// takes a slice of the exchanged vector x = [G_w | energy | grad[idx[*]]]: it stores the slice into
// slot `rank` of every peer's receive buffer over NVLink, publishes its sequence number with a
// system-scope release, waits until all `world` flags of its own buffer carry the number, and
// adds the slots in rank order (identical rounding on every rank -> replicated parameters stay
// bit-identical).  Receive buffers are double-buffered by sequence parity: a rank can be at most
// one exchange ahead of the slowest one, because completing exchange s needs every rank's flag s.

constexpr int kFinishThreads = 512;
constexpr int kFinishRegions = 64;
constexpr long long kSpinLimit = 4000000000ll;     // ~2 s of SM clocks

template <typename real>
struct FinishArgs {
    const double* partials;
    long long regions;
    int K;
    real* grad;
    long long n_param;
    double* step;
    double b1, b2;
    int world, rank;
    long long n_idx;
    const int* idx;
    real* recv[LHVI_MAX_PEERS];
    unsigned long long* flags[LHVI_MAX_PEERS];
    unsigned long long* seq;
    int* status;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Block-wide: sum the valid partial rows into grad[n_param ..] (G_w, energy) and, if a.step is
// given, advance the step counter.  Call with every thread of ONE block.
template <typename real>
__device__ void finish_reduce(const FinishArgs<real>& a, double* s, double* res) {
    const int W = a.K + 1;
    // all region headers first (one round trip), then one flat pass over the valid rows
    __shared__ int s_start[kFinishRegions + 1];
    const int regions = (int)(a.regions < kFinishRegions ? a.regions : kFinishRegions);
    if (threadIdx.x < regions)
        s_start[threadIdx.x + 1] = (int)a.partials[(long long)threadIdx.x * LHVI_PARTIAL_ROWS * W];
    __syncthreads();
    if (threadIdx.x == 0) {
        s_start[0] = 0;
        for (int i = 0; i < regions; ++i) s_start[i + 1] += s_start[i];
    }
    __syncthreads();
    double acc[LHVI_MAX_K + 1];
    for (int i = 0; i < W; ++i) acc[i] = 0.0;
    const int total = s_start[regions];
    for (int f = threadIdx.x; f < total; f += blockDim.x) {
        int reg = 0;
        while (f >= s_start[reg + 1]) ++reg;
        const double* row = a.partials + ((long long)reg * LHVI_PARTIAL_ROWS + 1 + (f - s_start[reg])) * W;
        for (int i = 0; i < W; ++i) acc[i] += row[i];
    }
    // more regions than the header table holds: the rest one by one
    for (long long reg = regions; reg < a.regions; ++reg) {
        const double* base = a.partials + reg * LHVI_PARTIAL_ROWS * W;
        const int valid = (int)base[0];
        for (int r = threadIdx.x; r < valid; r += blockDim.x)
            for (int i = 0; i < W; ++i) acc[i] += base[(long long)(1 + r) * W + i];
    }
    block_sum_to(acc, W, s, res);
    if (threadIdx.x < W) a.grad[a.n_param + threadIdx.x] = (real)res[threadIdx.x];
    if (threadIdx.x == 0 && a.step != nullptr) {
        // b^t as a running product (b^(t-1) = 1 - step[.]): no pow() on the critical path
        a.step[0] = a.step[0] + 1.0;
        a.step[1] = 1.0 - (1.0 - a.step[1]) * a.b1;
        a.step[2] = 1.0 - (1.0 - a.step[2]) * a.b2;
    }
    __syncthreads();
}

// The cross-GPU sum of x = [G_w | energy | grad[idx[*]]] over peer memory; `nblocks` blocks
// (blockIdx.x < nblocks) take part, each with its slice.
template <typename real>
__device__ void finish_exchange(const FinishArgs<real>& a, int nblocks) {
    const int W = a.K + 1;
    const long long n_x = a.n_idx + W;
    const long long chunk = (n_x + nblocks - 1) / nblocks;
    const long long lo = blockIdx.x * chunk;
    const long long hi = lo + chunk < n_x ? lo + chunk : n_x;
    const unsigned long long seq = a.seq[blockIdx.x] + 1ull;
    const size_t parity_base = (size_t)(seq & 1ull) * a.world * n_x;

    for (long long j = lo + threadIdx.x; j < hi; j += blockDim.x) {
        const long long src = j < W ? a.n_param + j : (long long)a.idx[j - W];
        const real v = a.grad[src];
        const size_t dst = parity_base + (size_t)a.rank * n_x + j;
        for (int p = 0; p < a.world; ++p) a.recv[p][dst] = v;
    }
    __syncthreads();
    // one thread per peer publishes the sequence number: each fences (cumulative over the block's
    // stores, which the barrier ordered before it) and then stores its flag, so the world - 1 NVLink
    // round trips overlap instead of queueing behind one thread's release stores
    if (threadIdx.x < a.world) {
        __threadfence_system();
        st_relaxed_sys(a.flags[threadIdx.x] + (size_t)a.rank * nblocks + blockIdx.x, seq);
    }
    if (threadIdx.x < a.world) {
        const unsigned long long* f = a.flags[a.rank] + (size_t)threadIdx.x * nblocks + blockIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > kSpinLimit) { *a.status = 1; break; }
        }
    }
    __syncthreads();
    const real* mine = a.recv[a.rank] + parity_base;
    for (long long j = lo + threadIdx.x; j < hi; j += blockDim.x) {
        real sum = real(0);
        for (int q = 0; q < a.world; ++q) sum += __ldcg(mine + (size_t)q * n_x + j);
        const long long src = j < W ? a.n_param + j : (long long)a.idx[j - W];
        a.grad[src] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) a.seq[blockIdx.x] = seq;
}

template <typename real>
__global__ void __launch_bounds__(kFinishThreads)
finish_kernel(const FinishArgs<real> a) {
    __shared__ double s[(kFinishThreads / 32) * (LHVI_MAX_K + 1)];
    __shared__ double res[LHVI_MAX_K + 1];
    if (blockIdx.x == 0) finish_reduce<real>(a, s, res);
    if (a.world <= 1) return;
    finish_exchange<real>(a, (int)gridDim.x);
}

// ---- optimiser step ------------------------------------------------------------------------

template <typename real>
struct StepArgs {
    int K;
    long long n_vars, n_param;
    const uint8_t* kind;
    const int* dim;
    const int* off;
    real* eta;
    real* tau;
    real* grad;
    real* m1;
    real* m2;
    real* wstate;
    const double* step;
    real lr, b1, b2, eps, var_floor;
    int sgd, zero_grad;
};

template <typename real> struct PairVec;
template <> struct PairVec<float> { using type = float4; static constexpr int pairs = 2; };
template <> struct PairVec<double> { using type = double2; static constexpr int pairs = 1; };

template <typename real>
__device__ __forceinline__ real moved(real theta, real g, real& m1, real& m2, const StepArgs<real>& a,
                                      real c1, real c2) {
    if (a.sgd) return theta - a.lr * g;
    m1 = m1 * a.b1 + (real(1) - a.b1) * g;
    m2 = m2 * a.b2 + (real(1) - a.b2) * g * g;
    // eps sits outside the square root (VarInference.py:272-273)
    return theta - (a.lr * (m1 / c1)) / (Math<real>::sqrt(m2 / c2) + a.eps);
}

// mixture weights: one thread (K <= 8)
template <typename real>
__device__ void step_mixture_weights(const StepArgs<real>& a, real c1, real c2) {
    using M = Math<real>;
    const int K = a.K;
    real* w_tau = a.wstate;
    real* w = a.wstate + K;
    real* m1 = a.wstate + 2 * K;
    real* m2 = a.wstate + 3 * K;
    const real* G = a.grad + a.n_param;
    real dot = real(0);
    for (int k = 0; k < K; ++k) dot += G[k] * w[k];
    real mx = real(-1e30);
    for (int k = 0; k < K; ++k) {
        const real gk = w[k] * (G[k] - dot);                     // VarInference.py:90
        w_tau[k] = moved<real>(w_tau[k], gk, m1[k], m2[k], a, c1, c2);
        mx = w_tau[k] > mx ? w_tau[k] : mx;
    }
    real z = real(0);
    for (int k = 0; k < K; ++k) z += M::exp(w_tau[k] - mx);
    for (int k = 0; k < K; ++k) w[k] = M::exp(w_tau[k] - mx) / z;
}

// the same with one lane per component (lanes 0..K-1 of ONE full warp; every load of a lane is
// independent, so the whole step costs one round trip to memory instead of a chain of them)
template <typename real>
__device__ void step_mixture_weights_warp(const StepArgs<real>& a, real c1, real c2) {
    using M = Math<real>;
    const int K = a.K, k = threadIdx.x & 31;
    const bool on = k < K;
    real* ws = a.wstate;
    const real G = on ? a.grad[a.n_param + k] : real(0);
    const real w = on ? ws[K + k] : real(0);
    real wt = on ? ws[k] : real(0);
    real m1 = on ? ws[2 * K + k] : real(0), m2 = on ? ws[3 * K + k] : real(0);
    real dot = G * w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const real gk = w * (G - dot);                                   // VarInference.py:90
    if (on) wt = moved<real>(wt, gk, m1, m2, a, c1, c2);
    real mx = on ? wt : real(-1e30);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const real v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
    const real e = on ? M::exp(wt - mx) : real(0);
    real z = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (on) {
        ws[k] = wt;
        ws[K + k] = e / z;
        if (!a.sgd) { ws[2 * K + k] = m1; ws[3 * K + k] = m2; }
    }
}

// one variable: softmax Jacobian (discrete), Adam / SGD, variance clip, re-normalisation, and the
// reset of the gradient slots it consumed
template <typename real>
__device__ __forceinline__ void step_variable(const StepArgs<real>& a, long long v, real c1, real c2) {
    using M = Math<real>;
    const int K = a.K;
    const int off = a.off[v];
    if (a.kind[v] == 0) {
        // slots are 16-byte aligned (8 for K == 1): move them with vector loads / stores
        using V = typename PairVec<real>::type;
        constexpr int PER = PairVec<real>::pairs;                 // (mu, var) pairs per vector
        const int nvec = (K + PER - 1) / PER;
        for (int c = 0; c < nvec; ++c) {
            const int i = off + 2 * PER * c;
            if (PER == 2 && K == 1) {                              // lone pair: scalar fallback
                a.eta[i] = moved<real>(a.eta[i], a.grad[i], a.m1[i], a.m2[i], a, c1, c2);
                real var = moved<real>(a.eta[i + 1], a.grad[i + 1], a.m1[i + 1], a.m2[i + 1], a, c1, c2);
                a.eta[i + 1] = var < a.var_floor ? a.var_floor : var;
                if (a.zero_grad) { a.grad[i] = real(0); a.grad[i + 1] = real(0); }
                continue;
            }
            V ve = *reinterpret_cast<V*>(a.eta + i);
            const V vg = *reinterpret_cast<const V*>(a.grad + i);
            V vm = *reinterpret_cast<V*>(a.m1 + i);
            V vu = *reinterpret_cast<V*>(a.m2 + i);
            real* e = reinterpret_cast<real*>(&ve);
            const real* gq = reinterpret_cast<const real*>(&vg);
            real* m = reinterpret_cast<real*>(&vm);
            real* u = reinterpret_cast<real*>(&vu);
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                if (PER * c + p < K) {
                    e[2 * p] = moved<real>(e[2 * p], gq[2 * p], m[2 * p], u[2 * p], a, c1, c2);
                    const real var = moved<real>(e[2 * p + 1], gq[2 * p + 1], m[2 * p + 1], u[2 * p + 1], a, c1, c2);
                    e[2 * p + 1] = var < a.var_floor ? a.var_floor : var;   // VarInference.py:281
                }
            }
            *reinterpret_cast<V*>(a.eta + i) = ve;
            *reinterpret_cast<V*>(a.m1 + i) = vm;
            *reinterpret_cast<V*>(a.m2 + i) = vu;
            if (a.zero_grad) {
                V zero;
                real* zq = reinterpret_cast<real*>(&zero);
#pragma unroll
                for (int p = 0; p < 2 * PER; ++p) zq[p] = real(0);
                *reinterpret_cast<V*>(a.grad + i) = zero;
            }
        }
    } else {
        const int D = a.dim[v];
        for (int k = 0; k < K; ++k) {
            const int row = off + k * D;
            real dot = real(0);
            for (int d = 0; d < D; ++d) dot += a.grad[row + d] * a.eta[row + d];
            real mx = real(-1e30);
            for (int d = 0; d < D; ++d) {
                const real p = a.eta[row + d];
                const real gt = p * (a.grad[row + d] - dot);          // VarInference.py:160
                const real t = moved<real>(a.tau[row + d], gt, a.m1[row + d], a.m2[row + d], a, c1, c2);
                a.tau[row + d] = t;
                mx = t > mx ? t : mx;
            }
            real z = real(0);
            for (int d = 0; d < D; ++d) z += M::exp(a.tau[row + d] - mx);
            const real iz = real(1) / z;
            for (int d = 0; d < D; ++d) a.eta[row + d] = M::exp(a.tau[row + d] - mx) * iz;
            if (a.zero_grad)
                for (int d = 0; d < D; ++d) a.grad[row + d] = real(0);
        }
    }
}

template <typename real>
__global__ void __launch_bounds__(256)
param_step_kernel(const StepArgs<real> a) {
    const real c1 = (real)a.step[1], c2 = (real)a.step[2];
    if (blockIdx.x == 0 && threadIdx.x == 0) step_mixture_weights<real>(a, c1, c2);
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < a.n_vars;
         v += (long long)gridDim.x * blockDim.x)
        step_variable<real>(a, v, c1, c2);
}

// ---- finish + step in one launch ---------------------------------------------------------------
//
// Block 0 does lhvi_finish's work (partial rows -> G_w, energy; with several GPUs the exchange)
// and then steps the mixture weights and the *shared* variables (a.n_vars - n_owned of them, listed
// last), whose gradients only exist after the exchange.  Every other block steps owned variables
// right away: they depend on nothing block 0 does, so the reduction and the NVLink round trip are
// off the critical path and one launch level of the iteration is gone.  The step counter must
// already be advanced for this iteration (lhvi_step_tick, launched beside the factor kernels).
template <typename real>
__global__ void __launch_bounds__(256)
finish_step_kernel(const FinishArgs<real> f, const StepArgs<real> a, long long n_owned) {
    __shared__ double s[(256 / 32) * (LHVI_MAX_K + 1)];
    __shared__ double res[LHVI_MAX_K + 1];
    const real c1 = (real)a.step[1], c2 = (real)a.step[2];
    if (blockIdx.x == 0) {
        finish_reduce<real>(f, s, res);
        if (f.world > 1) finish_exchange<real>(f, 1);
        __syncthreads();
        if (threadIdx.x == 0) step_mixture_weights<real>(a, c1, c2);
        for (long long v = n_owned + threadIdx.x; v < a.n_vars; v += blockDim.x) step_variable<real>(a, v, c1, c2);
        return;
    }
    for (long long v = (blockIdx.x - 1) * (long long)blockDim.x + threadIdx.x; v < n_owned;
         v += (long long)(gridDim.x - 1) * blockDim.x)
        step_variable<real>(a, v, c1, c2);
}

// The same launch for a model whose variables are all continuous and whose slots are contiguous (variable v
// at element v * slot): one thread per 16-byte VECTOR of the parameter array instead of one per variable.
// Adjacent threads then touch adjacent vectors -- the per-variable loop reads every second 16 bytes per
// request (ncu on the bench model: lg_throttle 7.5 and long_scoreboard 19 stall cycles per issue, 24 us) --
// and the kernel carries neither the variable table nor the discrete branch.  `slot_vecs` vectors per slot;
// pairs beyond K in a slot's last vector are padding and pass through unchanged.
template <typename real>
__global__ void __launch_bounds__(256)
finish_step_flat_kernel(const FinishArgs<real> f, const StepArgs<real> a, long long n_vec, int slot_vecs) {
    __shared__ double s[(256 / 32) * (LHVI_MAX_K + 1)];
    __shared__ double res[LHVI_MAX_K + 1];
    const real c1 = (real)a.step[1], c2 = (real)a.step[2];
    if (blockIdx.x == 0) {
        finish_reduce<real>(f, s, res);
        __syncthreads();
        if (threadIdx.x == 0) step_mixture_weights<real>(a, c1, c2);
        return;
    }
    using V = typename PairVec<real>::type;
    constexpr int PER = PairVec<real>::pairs;
    const int K = a.K;
    for (long long i = (blockIdx.x - 1) * (long long)blockDim.x + threadIdx.x; i < n_vec;
         i += (long long)(gridDim.x - 1) * blockDim.x) {
        const int c = (int)(i % slot_vecs);
        const long long e0 = i * (2 * PER);
        V ve = *reinterpret_cast<const V*>(a.eta + e0);
        const V vg = *reinterpret_cast<const V*>(a.grad + e0);
        V vm = *reinterpret_cast<const V*>(a.m1 + e0);
        V vu = *reinterpret_cast<const V*>(a.m2 + e0);
        real* e = reinterpret_cast<real*>(&ve);
        const real* gq = reinterpret_cast<const real*>(&vg);
        real* m = reinterpret_cast<real*>(&vm);
        real* u = reinterpret_cast<real*>(&vu);
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            if (PER * c + p < K) {
                e[2 * p] = moved<real>(e[2 * p], gq[2 * p], m[2 * p], u[2 * p], a, c1, c2);
                const real var = moved<real>(e[2 * p + 1], gq[2 * p + 1], m[2 * p + 1], u[2 * p + 1], a, c1, c2);
                e[2 * p + 1] = var < a.var_floor ? a.var_floor : var;   // VarInference.py:281
            }
        }
        *reinterpret_cast<V*>(a.eta + e0) = ve;
        *reinterpret_cast<V*>(a.m1 + e0) = vm;
        *reinterpret_cast<V*>(a.m2 + e0) = vu;
        V zero;
        real* zq = reinterpret_cast<real*>(&zero);
#pragma unroll
        for (int p = 0; p < 2 * PER; ++p) zq[p] = real(0);
        *reinterpret_cast<V*>(a.grad + e0) = zero;
    }
}

// ---- batched belief queries ------------------------------------------------------------------

template <typename real>
__global__ void __launch_bounds__(256)
mixture_belief_kernel(int K, long long n, const int* __restrict__ q_off, const int* __restrict__ q_dim,
                      const uint8_t* __restrict__ q_kind, const real* __restrict__ x,
                      const real* __restrict__ eta, const real* __restrict__ w, real* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const real* p = eta + q_off[i];
        real b = real(0);
        if (q_kind[i] == 0) {
            for (int k = 0; k < K; ++k) b += w[k] * norm_pdf<real>(x[i], p[2 * k], p[2 * k + 1]);
        } else {
            const int D = q_dim[i];
            const int d = (int)x[i];
            if (d >= 0 && d < D)
                for (int k = 0; k < K; ++k) b += w[k] * p[k * D + d];
        }
        out[i] = b;
    }
}

// ---- batched MAP queries ----------------------------------------------------------------------
// Continuous variable: start at the component mean with the largest belief, then safeguarded
// Newton ascent on the one-dimensional mixture (the reference starts at the same point and calls
// scipy's BFGS, VarInference.py:355-376).  Discrete variable: arg-max state of the mixture marginal.
// Always evaluated in double: the cost is negligible next to one iteration.

template <typename real>
__device__ double mixture_density(const real* __restrict__ p, const real* __restrict__ w, int K, double x) {
    double b = 0.0;
    for (int k = 0; k < K; ++k) b += (double)w[k] * norm_pdf_d(x, (double)p[2 * k], (double)p[2 * k + 1]);
    return b;
}

template <typename real>
__global__ void __launch_bounds__(128)
mixture_map_kernel(int K, long long n, const int* __restrict__ q_off, const int* __restrict__ q_dim,
                   const uint8_t* __restrict__ q_kind, const real* __restrict__ eta, const real* __restrict__ w,
                   real* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const real* p = eta + q_off[i];
        if (q_kind[i] != 0) {
            const int D = q_dim[i];
            int best = 0;
            double best_v = -1.0;
            for (int d = 0; d < D; ++d) {
                double v = 0.0;
                for (int k = 0; k < K; ++k) v += (double)w[k] * (double)p[k * D + d];
                if (v > best_v) { best_v = v; best = d; }
            }
            out[i] = (real)best;
            continue;
        }
        double x = (double)p[0], fx = mixture_density<real>(p, w, K, x);
        for (int k = 1; k < K; ++k) {
            const double v = mixture_density<real>(p, w, K, (double)p[2 * k]);
            if (v > fx) { fx = v; x = (double)p[2 * k]; }
        }
        for (int it = 0; it < 60; ++it) {
            double g1 = 0.0, g2 = 0.0;
            for (int k = 0; k < K; ++k) {
                const double mu = (double)p[2 * k], var = (double)p[2 * k + 1];
                const double d = x - mu, pk = (double)w[k] * norm_pdf_d(x, mu, var);
                g1 -= pk * d / var;
                g2 += pk * (d * d / (var * var) - 1.0 / var);
            }
            double step = g2 < 0.0 ? -g1 / g2 : (g1 > 0.0 ? 0.1 : (g1 < 0.0 ? -0.1 : 0.0));
            double fc = fx;
            for (int bt = 0; bt < 30; ++bt) {                 // never accept a lower density
                fc = mixture_density<real>(p, w, K, x + step);
                if (!(fc < fx)) break;
                step *= 0.5;
            }
            if (fc < fx) break;
            x += step;
            fx = fc;
            if (fabs(step) < 1e-13) break;
        }
        out[i] = (real)x;
    }
}

static inline unsigned grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)b;
}

}  // namespace lhvi
