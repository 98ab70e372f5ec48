// Elementwise kernels around the factor pass: ELBO / G_w reduction, the optimiser step
// (softmax Jacobians + Adam or SGD + variance clip + re-normalisation) and batched belief
// queries.  All are HBM-streaming kernels over the flat parameter vector.
#include <cstring>

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

__global__ void step_tick_kernel(double* step, double b1, double b2) {
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

static unsigned grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)b;
}

}  // namespace lhvi

using namespace lhvi;

extern "C" int lhvi_elbo_reduce(const lhvi_model* m, int64_t rows, void* stream) {
    if (!m || !m->partials || !m->grad || rows < 0) { set_error("lhvi_elbo_reduce: null buffer or negative rows"); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_elbo_reduce: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    if (m->dtype == LHVI_F64)
        elbo_reduce_kernel<double><<<1, 1024, 0, s>>>(m->partials, regions, m->K, (double*)m->grad + m->n_param);
    else
        elbo_reduce_kernel<float><<<1, 1024, 0, s>>>(m->partials, regions, m->K, (float*)m->grad + m->n_param);
    return check_launch("elbo_reduce_kernel");
}

template <typename real>
static int finish_t(const lhvi_model* m, long long regions, double* step, double b1, double b2,
                    const lhvi_exchange* x, cudaStream_t s) {
    FinishArgs<real> a;
    a.partials = m->partials; a.regions = regions; a.K = m->K;
    a.grad = (real*)m->grad; a.n_param = m->n_param;
    a.step = step; a.b1 = b1; a.b2 = b2;
    a.world = 1; a.rank = 0; a.n_idx = 0; a.idx = nullptr; a.seq = nullptr; a.status = nullptr;
    for (int p = 0; p < LHVI_MAX_PEERS; ++p) { a.recv[p] = nullptr; a.flags[p] = nullptr; }
    unsigned blocks = 1;
    if (x != nullptr && x->world > 1) {
        a.world = x->world; a.rank = x->rank; a.n_idx = x->n_idx; a.idx = x->idx;
        a.seq = (unsigned long long*)x->seq; a.status = x->status;
        for (int p = 0; p < x->world; ++p) {
            a.recv[p] = (real*)x->recv[p];
            a.flags[p] = (unsigned long long*)x->flags[p];
        }
        blocks = (unsigned)x->blocks;
    }
    finish_kernel<real><<<blocks, kFinishThreads, 0, s>>>(a);
    return check_launch("finish_kernel");
}

extern "C" int lhvi_finish(const lhvi_model* m, int64_t rows, double* step, double b1, double b2,
                           const lhvi_exchange* x, void* stream) {
    if (!m || !m->partials || !m->grad || rows < 0) { set_error("lhvi_finish: null buffer or negative rows"); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_finish: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    if (x != nullptr && x->world > 1) {
        if (x->world > LHVI_MAX_PEERS) { set_error("lhvi_finish: world=%d exceeds LHVI_MAX_PEERS=%d", x->world, LHVI_MAX_PEERS); return LHVI_ELIMIT; }
        if (x->rank < 0 || x->rank >= x->world) { set_error("lhvi_finish: rank %d outside world %d", x->rank, x->world); return LHVI_EINVAL; }
        if (x->blocks < 1 || x->blocks > 32) { set_error("lhvi_finish: blocks=%d out of range 1..32", x->blocks); return LHVI_EINVAL; }
        if (x->n_idx < 0 || (x->n_idx > 0 && !x->idx) || !x->seq || !x->status) { set_error("lhvi_finish: incomplete exchange descriptor"); return LHVI_EINVAL; }
        for (int p = 0; p < x->world; ++p)
            if (!x->recv[p] || !x->flags[p]) { set_error("lhvi_finish: peer %d has no mapped buffer", p); return LHVI_EINVAL; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    return m->dtype == LHVI_F64 ? finish_t<double>(m, regions, step, b1, b2, x, s)
                                : finish_t<float>(m, regions, step, b1, b2, x, s);
}

// ---- peer-visible buffers (CUDA IPC) ---------------------------------------------------------

static int cuda_fail(const char* what, cudaError_t e) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return LHVI_ECUDA;
}

extern "C" int lhvi_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
    if (bytes <= 0 || !ptr || !handle) { set_error("lhvi_peer_alloc: bad argument"); return LHVI_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == LHVI_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail("lhvi_peer_alloc: cudaMalloc", e);
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail("lhvi_peer_alloc", e); }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return LHVI_OK;
}

extern "C" int lhvi_peer_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) { set_error("lhvi_peer_open: bad argument"); return LHVI_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail("lhvi_peer_open: cudaIpcOpenMemHandle", e);
    *ptr = p;
    return LHVI_OK;
}

extern "C" int lhvi_peer_close(void* ptr) {
    if (!ptr) return LHVI_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? LHVI_OK : cuda_fail("lhvi_peer_close", e);
}

extern "C" int lhvi_peer_free(void* ptr) {
    if (!ptr) return LHVI_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? LHVI_OK : cuda_fail("lhvi_peer_free", e);
}

extern "C" int lhvi_step_tick(double* step, double b1, double b2, void* stream) {
    if (!step) { set_error("lhvi_step_tick: null step buffer"); return LHVI_EINVAL; }
    step_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, b1, b2);
    return check_launch("step_tick_kernel");
}

template <typename real>
static int param_step_t(int K, int64_t n_vars, const uint8_t* kind, const int32_t* dim, const int32_t* off,
                        void* eta, void* tau, void* grad, int64_t n_param, void* m1, void* m2,
                        void* wstate, const double* step, double lr, double b1, double b2, double eps,
                        double var_floor, int sgd, int zero_grad, cudaStream_t s) {
    StepArgs<real> a;
    a.K = K; a.n_vars = n_vars; a.n_param = n_param;
    a.kind = kind; a.dim = dim; a.off = off;
    a.eta = (real*)eta; a.tau = (real*)tau; a.grad = (real*)grad;
    a.m1 = (real*)m1; a.m2 = (real*)m2; a.wstate = (real*)wstate; a.step = step;
    a.lr = (real)lr; a.b1 = (real)b1; a.b2 = (real)b2; a.eps = (real)eps; a.var_floor = (real)var_floor;
    a.sgd = sgd; a.zero_grad = zero_grad;
    param_step_kernel<real><<<grid_for(n_vars, 256), 256, 0, s>>>(a);
    return check_launch("param_step_kernel");
}

extern "C" int lhvi_param_step(int dtype, int K, int64_t n_vars, const uint8_t* var_kind,
                               const int32_t* var_dim, const int32_t* var_off, void* eta, void* tau,
                               void* grad, int64_t n_param, void* mom1, void* mom2, void* wstate,
                               const double* step, double lr, double b1, double b2, double eps,
                               double var_threshold, int sgd, int zero_grad, void* stream) {
    if (!eta || !tau || !grad || !mom1 || !mom2 || !wstate || !step) { set_error("lhvi_param_step: null buffer"); return LHVI_EINVAL; }
    if (n_vars > 0 && (!var_kind || !var_dim || !var_off)) { set_error("lhvi_param_step: null variable table"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == LHVI_F64
        ? param_step_t<double>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, zero_grad, s)
        : param_step_t<float>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, zero_grad, s);
}

template <typename real>
static int finish_step_t(const lhvi_model* m, long long regions, const lhvi_exchange* x, int K, int64_t n_vars,
                         int64_t n_owned, const uint8_t* kind, const int32_t* dim, const int32_t* off, void* eta,
                         void* tau, void* m1, void* m2, void* wstate, const double* step, double lr, double b1,
                         double b2, double eps, double var_floor, int sgd, cudaStream_t s) {
    FinishArgs<real> f;
    f.partials = m->partials; f.regions = regions; f.K = m->K;
    f.grad = (real*)m->grad; f.n_param = m->n_param;
    f.step = nullptr; f.b1 = b1; f.b2 = b2;
    f.world = 1; f.rank = 0; f.n_idx = 0; f.idx = nullptr; f.seq = nullptr; f.status = nullptr;
    for (int p = 0; p < LHVI_MAX_PEERS; ++p) { f.recv[p] = nullptr; f.flags[p] = nullptr; }
    if (x != nullptr && x->world > 1) {
        f.world = x->world; f.rank = x->rank; f.n_idx = x->n_idx; f.idx = x->idx;
        f.seq = (unsigned long long*)x->seq; f.status = x->status;
        for (int p = 0; p < x->world; ++p) {
            f.recv[p] = (real*)x->recv[p];
            f.flags[p] = (unsigned long long*)x->flags[p];
        }
    }
    StepArgs<real> a;
    a.K = K; a.n_vars = n_vars; a.n_param = m->n_param;
    a.kind = kind; a.dim = dim; a.off = off;
    a.eta = (real*)eta; a.tau = (real*)tau; a.grad = (real*)m->grad;
    a.m1 = (real*)m1; a.m2 = (real*)m2; a.wstate = (real*)wstate; a.step = step;
    a.lr = (real)lr; a.b1 = (real)b1; a.b2 = (real)b2; a.eps = (real)eps; a.var_floor = (real)var_floor;
    a.sgd = sgd; a.zero_grad = 1;
    finish_step_kernel<real><<<1 + grid_for(n_owned > 0 ? n_owned : 1, 256), 256, 0, s>>>(f, a, (long long)n_owned);
    return check_launch("finish_step_kernel");
}

extern "C" int lhvi_finish_step(const lhvi_model* m, int64_t rows, const lhvi_exchange* x, int64_t n_vars,
                                int64_t n_owned, const uint8_t* var_kind, const int32_t* var_dim,
                                const int32_t* var_off, void* tau, void* mom1, void* mom2, void* wstate,
                                const double* step, double lr, double b1, double b2, double eps,
                                double var_threshold, int sgd, void* stream) {
    if (!m || !m->partials || !m->grad || !m->eta || rows < 0) { set_error("lhvi_finish_step: null buffer or negative rows"); return LHVI_EINVAL; }
    if (!tau || !mom1 || !mom2 || !wstate || !step) { set_error("lhvi_finish_step: null buffer"); return LHVI_EINVAL; }
    if (n_vars > 0 && (!var_kind || !var_dim || !var_off)) { set_error("lhvi_finish_step: null variable table"); return LHVI_EINVAL; }
    if (n_owned < 0 || n_owned > n_vars) { set_error("lhvi_finish_step: n_owned=%lld outside 0..n_vars=%lld", (long long)n_owned, (long long)n_vars); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_finish_step: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    if (x != nullptr && x->world > 1) {
        if (x->world > LHVI_MAX_PEERS) { set_error("lhvi_finish_step: world=%d exceeds LHVI_MAX_PEERS=%d", x->world, LHVI_MAX_PEERS); return LHVI_ELIMIT; }
        if (x->rank < 0 || x->rank >= x->world) { set_error("lhvi_finish_step: rank %d outside world %d", x->rank, x->world); return LHVI_EINVAL; }
        if (x->blocks != 1) { set_error("lhvi_finish_step: the fused launch exchanges with one block (blocks=%d): use lhvi_finish + lhvi_param_step", x->blocks); return LHVI_EINVAL; }
        if (x->n_idx < 0 || (x->n_idx > 0 && !x->idx) || !x->seq || !x->status) { set_error("lhvi_finish_step: incomplete exchange descriptor"); return LHVI_EINVAL; }
        for (int p = 0; p < x->world; ++p)
            if (!x->recv[p] || !x->flags[p]) { set_error("lhvi_finish_step: peer %d has no mapped buffer", p); return LHVI_EINVAL; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    void* eta = const_cast<void*>(m->eta);
    return m->dtype == LHVI_F64
        ? finish_step_t<double>(m, regions, x, m->K, n_vars, n_owned, var_kind, var_dim, var_off, eta, tau, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s)
        : finish_step_t<float>(m, regions, x, m->K, n_vars, n_owned, var_kind, var_dim, var_off, eta, tau, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s);
}

extern "C" int lhvi_mixture_belief(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                                   const uint8_t* q_kind, const void* x, const void* eta, const void* w,
                                   void* out, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!q_off || !q_dim || !q_kind || !x || !eta || !w || !out) { set_error("lhvi_mixture_belief: null buffer"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LHVI_F64)
        mixture_belief_kernel<double><<<grid_for(n, 256), 256, 0, s>>>(K, n, q_off, q_dim, q_kind, (const double*)x, (const double*)eta, (const double*)w, (double*)out);
    else
        mixture_belief_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(K, n, q_off, q_dim, q_kind, (const float*)x, (const float*)eta, (const float*)w, (float*)out);
    return check_launch("mixture_belief_kernel");
}

extern "C" int lhvi_mixture_map(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                                const uint8_t* q_kind, const void* eta, const void* w, void* out, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!q_off || !q_dim || !q_kind || !eta || !w || !out) { set_error("lhvi_mixture_map: null buffer"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LHVI_F64)
        mixture_map_kernel<double><<<grid_for(n, 128), 128, 0, s>>>(K, n, q_off, q_dim, q_kind, (const double*)eta, (const double*)w, (double*)out);
    else
        mixture_map_kernel<float><<<grid_for(n, 128), 128, 0, s>>>(K, n, q_off, q_dim, q_kind, (const float*)eta, (const float*)w, (float*)out);
    return check_launch("mixture_map_kernel");
}

// ---- compact host layout <-> padded device slots ------------------------------------------------

namespace lhvi {
template <typename real, bool PACK>
__global__ void __launch_bounds__(256)
state_move_kernel(long long n, const int* __restrict__ map, const real* __restrict__ src, real* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (PACK) dst[i] = src[map[i]];
        else dst[map[i]] = src[i];
    }
}
}  // namespace lhvi

template <bool PACK>
static int state_move(int dtype, int64_t n, const int32_t* map, const void* src, void* dst, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!map || !src || !dst) { lhvi::set_error("lhvi_state_pack/unpack: null buffer"); return LHVI_EINVAL; }
    if (dtype != LHVI_F32 && dtype != LHVI_F64) { lhvi::set_error("dtype %d is neither LHVI_F32 nor LHVI_F64", dtype); return LHVI_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == LHVI_F64)
        lhvi::state_move_kernel<double, PACK><<<(unsigned)blocks, 256, 0, s>>>(n, map, (const double*)src, (double*)dst);
    else
        lhvi::state_move_kernel<float, PACK><<<(unsigned)blocks, 256, 0, s>>>(n, map, (const float*)src, (float*)dst);
    return lhvi::check_launch("state_move_kernel");
}

extern "C" int lhvi_state_pack(int dtype, int64_t n, const int32_t* map, const void* state, void* packed, void* stream) {
    return state_move<true>(dtype, n, map, state, packed, stream);
}

extern "C" int lhvi_state_unpack(int dtype, int64_t n, const int32_t* map, const void* packed, void* state, void* stream) {
    return state_move<false>(dtype, n, map, packed, state, stream);
}
