// Elementwise kernels around the factor pass: ELBO / G_w reduction, the optimiser step
// (softmax Jacobians + Adam or SGD + variance clip + re-normalisation) and batched belief
// queries.  All are HBM-streaming kernels over the flat parameter vector.
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
    const real* grad;
    real* m1;
    real* m2;
    real* wstate;
    const double* step;
    real lr, b1, b2, eps, var_floor;
    int sgd;
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

template <typename real>
__global__ void __launch_bounds__(256)
param_step_kernel(const StepArgs<real> a) {
    using M = Math<real>;
    const int K = a.K;
    const real c1 = (real)a.step[1], c2 = (real)a.step[2];

    // mixture weights: one thread (K <= 8)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
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

    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < a.n_vars;
         v += (long long)gridDim.x * blockDim.x) {
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
            }
        }
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

extern "C" int lhvi_step_tick(double* step, double b1, double b2, void* stream) {
    if (!step) { set_error("lhvi_step_tick: null step buffer"); return LHVI_EINVAL; }
    step_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, b1, b2);
    return check_launch("step_tick_kernel");
}

template <typename real>
static int param_step_t(int K, int64_t n_vars, const uint8_t* kind, const int32_t* dim, const int32_t* off,
                        void* eta, void* tau, const void* grad, int64_t n_param, void* m1, void* m2,
                        void* wstate, const double* step, double lr, double b1, double b2, double eps,
                        double var_floor, int sgd, cudaStream_t s) {
    StepArgs<real> a;
    a.K = K; a.n_vars = n_vars; a.n_param = n_param;
    a.kind = kind; a.dim = dim; a.off = off;
    a.eta = (real*)eta; a.tau = (real*)tau; a.grad = (const real*)grad;
    a.m1 = (real*)m1; a.m2 = (real*)m2; a.wstate = (real*)wstate; a.step = step;
    a.lr = (real)lr; a.b1 = (real)b1; a.b2 = (real)b2; a.eps = (real)eps; a.var_floor = (real)var_floor;
    a.sgd = sgd;
    param_step_kernel<real><<<grid_for(n_vars, 256), 256, 0, s>>>(a);
    return check_launch("param_step_kernel");
}

extern "C" int lhvi_param_step(int dtype, int K, int64_t n_vars, const uint8_t* var_kind,
                               const int32_t* var_dim, const int32_t* var_off, void* eta, void* tau,
                               const void* grad, int64_t n_param, void* mom1, void* mom2, void* wstate,
                               const double* step, double lr, double b1, double b2, double eps,
                               double var_threshold, int sgd, void* stream) {
    if (!eta || !tau || !grad || !mom1 || !mom2 || !wstate || !step) { set_error("lhvi_param_step: null buffer"); return LHVI_EINVAL; }
    if (n_vars > 0 && (!var_kind || !var_dim || !var_off)) { set_error("lhvi_param_step: null variable table"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == LHVI_F64
        ? param_step_t<double>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s)
        : param_step_t<float>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s);
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
