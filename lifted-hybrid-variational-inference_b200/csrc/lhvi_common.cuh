// Shared device helpers for the LHVI kernels (sm_100a).
//
// Numerical contract: the reference computes everything in float64 with
//   q_k(x)  = exp(-(x-mu)^2 / (2 var)) / (2.506628274631 * var)      (VarInference.py:26-30)
//   F       = log(psi + 1e-100) - log(b + 1e-100)                    (VarInference.py:76)
// The double instantiation evaluates those expressions literally.  The float instantiation
// cannot represent 1e-100, so it uses the mathematically equivalent branches below and
// falls back to double arithmetic on the (rare) points where float would underflow.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lhvi.h"

namespace lhvi {

constexpr double kEps = 1e-100;          // the reference's floor inside both logs
constexpr double kSqrt2Pi = 2.506628274631;   // the reference's literal, not sqrt(2*pi)
constexpr double kLogEps = -230.25850929940458;  // log(1e-100)

template <typename real> struct Math;

template <> struct Math<double> {
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    // log(psi + 1e-100) with psi = exp(q)
    static __device__ __forceinline__ double log_psi(double q) {
        // for q > -180 the floor changes the result by < 1e-21 absolute: below double rounding
        return q > -180.0 ? q : ::log(::exp(q) + kEps);
    }
    // log(b + 1e-100)
    static __device__ __forceinline__ double log_belief(double b) { return ::log(b + kEps); }
    static __device__ __forceinline__ bool belief_underflow(double) { return false; }
};

// float cannot hold 1e-100: log(exp(q) + 1e-100) - q = log1p(exp(-230.26 - q)) is below half a
// float ulp of q for q > -218, so only deep-tail points take this out-of-line double evaluation
static __device__ __noinline__ float log_psi_deep_tail(float q) {
    return (float)::log(::exp((double)q) + kEps);
}

template <> struct Math<float> {
#ifdef LHVI_FAST_MATH
    static __device__ __forceinline__ float exp(float x) { return __expf(x); }
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
#else
    static __device__ __forceinline__ float exp(float x) { return expf(x); }
    static __device__ __forceinline__ float log(float x) { return logf(x); }
#endif
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float log_psi(float q) {
        if (q > -210.0f) return q;       // the 1e-100 floor is below float resolution up here
        return log_psi_deep_tail(q);
    }
    // valid only when !belief_underflow(b); callers recompute in double otherwise
    static __device__ __forceinline__ float log_belief(float b) { return log(b); }
    static __device__ __forceinline__ bool belief_underflow(float b) { return !(b > 1e-30f); }
};

// Reference component density (note the 1/var normaliser).
template <typename real>
__device__ __forceinline__ real norm_pdf(real x, real mu, real var) {
    real u = x - mu;
    real inv = Math<real>::rcp(var);
    return Math<real>::exp(real(-0.5) * u * u * inv) * (inv * real(1.0 / kSqrt2Pi));
}

__device__ __forceinline__ double norm_pdf_d(double x, double mu, double var) {
    double u = x - mu;
    return ::exp(-0.5 * u * u / var) / (kSqrt2Pi * var);
}

// ---- reductions -------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sums of `nvals` doubles per thread; thread 0 of the block writes them to out[].
// `scratch` must hold (blockDim.x / 32) * nvals doubles of shared memory.
// Every factor kernel ends with this: block b writes its partial sums to data row b of the
// launch's region and block 0 records how many rows are valid in the header row, so the
// reduction needs no zero-filled buffer.

__device__ __forceinline__ void block_sum_to(const double* vals, int nvals, double* scratch,
                                             double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = 0; i < nvals; ++i) {
        double s = warp_sum(vals[i]);
        if (lane == 0) scratch[warp * nvals + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < nvals) {
        double s = 0.0;
        for (int wp = 0; wp < nwarps; ++wp) s += scratch[wp * nvals + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// Which slice of a record group a thread block works on: block `bid` of the `nblocks` blocks working
// on the group (the launch's blockIdx / gridDim for a per-group kernel; a share of the persistent
// grid inside the iteration kernel).  accum != nullptr: the block adds its (G_w, energy) sums to
// these K + 1 doubles with atomics instead of writing a row of `partials` (the iteration kernel:
// nothing is left to reduce after the grid barrier).
struct BlockSlice {
    int bid, nblocks;
    double* accum;
};

__device__ __forceinline__ void publish_partials(const double* vals, int nvals, double* scratch, double* region,
                                                 const BlockSlice bs) {
    if (bs.accum != nullptr) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int i = 0; i < nvals; ++i) {
            double s = warp_sum(vals[i]);
            if (lane == 0) scratch[warp * nvals + i] = s;
        }
        __syncthreads();
        if (threadIdx.x < nvals) {
            double s = 0.0;
            for (int wp = 0; wp < nwarps; ++wp) s += scratch[wp * nvals + threadIdx.x];
            atomicAdd(bs.accum + threadIdx.x, s);
        }
        __syncthreads();
        return;
    }
    if (bs.bid == 0 && threadIdx.x == 0) region[-nvals] = (double)bs.nblocks;
    block_sum_to(vals, nvals, scratch, region + (long long)bs.bid * nvals);
}

__device__ __forceinline__ void publish_partials(const double* vals, int nvals, double* scratch, double* region) {
    publish_partials(vals, nvals, scratch, region, BlockSlice{(int)blockIdx.x, (int)gridDim.x, nullptr});
}

// ---- typed view of a record group ---------------------------------------------------------

template <typename real>
struct GroupView {
    int nd, nc, ng, ne, node, weighted, pure;
    int no_cat;           // lhvi_group::no_category_grad
    int pot_kind;         // LHVI_POT_*
    int dims[LHVI_MAX_AXES];
    long long n;
    const int* pot;
    const int* poff;
    const real* egval;
    const real* egvar;
    const real* ecval;
    const real* wf;
    const real* gam;
    const real* nscale;
    const real* fold;
    long long n_pad;
    const int* run_start;
    const int* run_key;
    const int* run_hid;
    const int* hub_keys;
    long long n_runs;
    const real* run_node;       // lhvi_group::run_node / run_una_pot / run_una_w (fused records of the run variable)
    const int* run_una_pot;
    const real* run_una_w;
    const real* cst_q;          // lhvi_group::cst_* (constant records evaluated by the streaming kernel)
    const real* cst_wf;
    long long cst_n;
    // model
    int K, T;
    const real* quad;
    const real* ptab;
    const real* eta;
    const real* w;
    real* grad;
    double* partials;   // this launch's first data row (the region's header row sits just before)
};

template <typename real>
inline GroupView<real> make_view(const lhvi_model* m, const lhvi_group* g, int64_t row0) {
    GroupView<real> v;
    v.nd = g->nd; v.nc = g->nc; v.ng = g->ng; v.ne = g->ne;
    v.node = g->node; v.weighted = g->weighted; v.pure = g->pure;
    v.no_cat = g->no_category_grad;
    v.pot_kind = g->pot_kind;
    for (int i = 0; i < LHVI_MAX_AXES; ++i) v.dims[i] = g->dims[i];
    v.n = g->n;
    v.pot = g->pot; v.poff = g->poff;
    v.egval = (const real*)g->egval; v.egvar = (const real*)g->egvar;
    v.ecval = (const real*)g->ecval; v.wf = (const real*)g->wf;
    v.gam = (const real*)g->gam; v.nscale = (const real*)g->nscale;
    v.fold = (const real*)g->fold; v.n_pad = g->n_pad;
    v.run_start = g->run_start; v.run_key = g->run_key; v.run_hid = g->run_hid;
    v.hub_keys = g->hub_keys; v.n_runs = g->n_runs;
    v.run_node = (const real*)g->run_node; v.run_una_pot = g->run_una_pot; v.run_una_w = (const real*)g->run_una_w;
    v.cst_q = (const real*)g->cst_q; v.cst_wf = (const real*)g->cst_wf; v.cst_n = g->cst_n;
    v.K = m->K; v.T = m->T;
    v.quad = (const real*)m->quad; v.ptab = (const real*)m->ptab;
    v.eta = (const real*)m->eta; v.w = (const real*)m->w;
    v.grad = (real*)m->grad;
    // row `row0` of the launch's region is a header (valid row count), data rows follow
    v.partials = m->partials + (row0 + 1) * (m->K + 1);
    return v;
}

// error reporting shared by the translation units
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// launchers implemented per translation unit
int launch_generic(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s);
int launch_spec(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s);   // 1 = no specialisation
bool spec_available(const lhvi_model* m, const lhvi_group* g);

}  // namespace lhvi
