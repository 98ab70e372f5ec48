// Template-specialised factor kernels for records with hidden DISCRETE arguments (and for quadrature
// degree 10): compile-time K / T / argument counts / state count, every table in registers, vector
// REDs for the gradient scatter.
//
// Reference: gradient_category_tau (VarInference.py:133-160) next to gradient_w_tau / gradient_mu_var /
// free_energy for factors that mix hidden discrete and hidden continuous arguments (the hybrid MLN
// formulas of Demo/HMLN, MLNPotential.py:36-37), with expectation() enumerating the discrete states
// and walking the quadrature grid of the continuous ones (VarInference.py:40-55).
//
// One thread per record.  Per mixture component k the thread tabulates the continuous axes (nodes,
// K cross densities per node) in registers, then enumerates the D^ND discrete configurations with a
// fully unrolled loop -- the configuration selects the coefficient block of the quadratic
// log-potential -- and walks the T^NC grid inside, evaluating
//     F = log(psi + 1e-100) - log(b + 1e-100)
// ONCE per point and feeding the energy, G_w, the (mu, var) gradients and the categorical gradients
// G_c[v,k,d] -= gamma * (W / eta_v[k,d]) * F from it (fused single pass).  The
// arithmetic is the generic kernel's (lhvi_generic.cu), so fp64 results agree to rounding; what
// changes is that nothing lives in local memory and that shapes the generic kernel loops over at run
// time are unrolled.
//
// Three flavours as in lhvi_spec_impl.cuh: full (F = log psi - log b), pure (unary split: F = log psi),
// node (F = log b: the variables' entropy terms).  All hidden discrete arguments of a group must have
// the same number of states D (2 or 3); other groups stay with the generic kernel.
#pragma once
#include "lhvi_common.cuh"
#include "lhvi_spec_impl.cuh"
#include "lhvi_hyb_sigs.h"

namespace lhvi {

constexpr int kHybThreads = 128;

template <int B, int E> struct IPow { static constexpr int value = B * IPow<B, E - 1>::value; };
template <int B> struct IPow<B, 0> { static constexpr int value = 1; };

// gradient of one hidden argument -> global memory
template <int N, int K, typename real>
__device__ __forceinline__ void hyb_scatter(real* p, const real (&v)[N]) {
    if constexpr (K >= 2) {
        red_vec<N>(p, v);               // every slot is a multiple of 16 bytes when K >= 2
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) atomicAdd(p + i, v[i]);
    }
}

// log(b + 1e-100) of one grid point with the belief evaluated in double from the parameters (float
// only: taken when the float products may have flushed to zero); out of line so that the unrolled
// grid loops do not carry its registers
template <typename real, int K, int ND, int NC>
struct HybPoint {
    real w[K];
    real pd[ND > 0 ? ND : 1][K];        // eta_a[k2][d_a] at this configuration
    real mu[NC > 0 ? NC : 1][K], var[NC > 0 ? NC : 1][K];
    real x[NC > 0 ? NC : 1];
};

template <typename real, int K, int ND, int NC>
__device__ __noinline__ real hyb_log_belief_double(const HybPoint<real, K, ND, NC> p) {
    double bd = 0.0;
    for (int k2 = 0; k2 < K; ++k2) {
        double v = (double)p.w[k2];
        for (int a = 0; a < ND; ++a) v *= (double)p.pd[a][k2];
        for (int c = 0; c < NC; ++c) v *= norm_pdf_d((double)p.x[c], (double)p.mu[c][k2], (double)p.var[c][k2]);
        bd += v;
    }
    return (real)::log(bd + kEps);
}

template <typename real, int K, int T, int ND, int D, int NC, int NE, int FL>
__global__ void __launch_bounds__(kHybThreads)
factor_hyb_kernel(const GroupView<real> g) {
    using M = Math<real>;
    constexpr int NH = ND + NC, NCT = NC + NE;
    constexpr int NCOEF = NCT == 0 ? 1 : (NCT + 1) * (NCT + 2) / 2;
    constexpr int NCFG = IPow<D, ND>::value;
    constexpr int NDS = ND > 0 ? ND : 1, NCS = NC > 0 ? NC : 1, NES = NE > 0 ? NE : 1;
    constexpr int TS = NC > 0 ? T : 1;
    static_assert(NC <= 2, "at most two hidden continuous arguments");
    static_assert(FL != kNode || NH == 1, "node records have one hidden argument");

    __shared__ real s_quad[2 * (T > 0 ? T : 1)];
    __shared__ real s_w[K];
    __shared__ double s_scratch[(kHybThreads / 32) * (K + 1)];
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    __syncthreads();
    real xi[TS], om[TS], wk[K];
#pragma unroll
    for (int t = 0; t < TS; ++t) { xi[t] = NC > 0 ? s_quad[t] : real(0); om[t] = NC > 0 ? s_quad[T + t] : real(1); }
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = s_w[k];

    double acc[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;

    const bool weighted = g.weighted != 0;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < g.n;
         r += (long long)gridDim.x * blockDim.x) {
        const real wf = (weighted || FL == kNode) ? g.wf[r] : real(1);
        int off[NH > 0 ? NH : 1];
#pragma unroll
        for (int a = 0; a < NH; ++a) off[a] = g.poff[a * g.n + r];
        real ev[NES];
#pragma unroll
        for (int e = 0; e < NE; ++e) ev[e] = g.ecval[e * g.n + r];
        const real* coef0 = FL == kNode ? nullptr : g.ptab + g.pot[r];

        // ---- parameters of the hidden arguments
        real pd[NDS][K][D];                 // categorical probabilities eta_a[k][d]
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            real slot[K * D];
            if constexpr (K >= 2) {
                load_vec<K * D>(g.eta + off[a], slot);
            } else {
#pragma unroll
                for (int i = 0; i < K * D; ++i) slot[i] = g.eta[off[a] + i];
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int d = 0; d < D; ++d) pd[a][k][d] = slot[k * D + d];
        }
        real mu[NCS][K], var[NCS][K];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            real slot[2 * K];
            load_vec<2 * K>(g.eta + off[ND + c], slot);
#pragma unroll
            for (int k = 0; k < K; ++k) { mu[c][k] = slot[2 * k]; var[c][k] = slot[2 * k + 1]; }
        }

        real gd[NDS][K][D], gc[NCS][2 * K];
#pragma unroll
        for (int a = 0; a < NDS; ++a)
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int d = 0; d < D; ++d) gd[a][k][d] = real(0);

#pragma unroll
        for (int k = 0; k < K; ++k) {
            // ---- the LAST continuous axis under component k is tabulated (nodes and the K cross
            // densities per node, in registers, its loop unrolled); with two continuous arguments the
            // first one is the outer loop and is evaluated on the fly (rolled when T is large: a
            // degree-10 rule has 100 grid points per configuration)
            constexpr int CL = NC > 0 ? NC - 1 : 0;
            real x[TS], q[K][TS];
            if constexpr (NC > 0) {
                const real sd = M::sqrt(real(2) * var[CL][k]);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    x[t] = sd * xi[t] + mu[CL][k];
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) q[k2][t] = norm_pdf<real>(x[t], mu[CL][k2], var[CL][k2]);
                }
            }
            const real sd0 = NC >= 2 ? M::sqrt(real(2) * var[0][k]) : real(0);
            real Ek = real(0), am[NCS], av[NCS];
#pragma unroll
            for (int c = 0; c < NCS; ++c) { am[c] = real(0); av[c] = real(0); }

            // ---- discrete configurations (fully unrolled), the grid of the continuous axes inside
#pragma unroll
            for (int cfg = 0; cfg < NCFG; ++cfg) {
                int dg[NDS];
                {
                    int rest = cfg;
#pragma unroll
                    for (int a = ND - 1; a >= 0; --a) { dg[a] = rest % D; rest /= D; }
                }
                real wd = real(1);                          // prod_a eta_a[k][d_a]
                real pk[K];                                 // w_k2 prod_a eta_a[k2][d_a]
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2) pk[k2] = wk[k2];
#pragma unroll
                for (int a = 0; a < ND; ++a) {
                    wd *= pd[a][k][dg[a]];
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk[k2] *= pd[a][k2][dg[a]];
                }
                // quadratic log-potential of this configuration, reduced by the point evidence
                real cst = real(0), lin[NCS], A[NCS][NCS];
#pragma unroll
                for (int i = 0; i < NCS; ++i) {
                    lin[i] = real(0);
#pragma unroll
                    for (int j = 0; j < NCS; ++j) A[i][j] = real(0);
                }
                if constexpr (FL != kNode) {
                    const real* cf = coef0 + cfg * NCOEF;
                    real cq[NCOEF];
#pragma unroll
                    for (int i = 0; i < NCOEF; ++i) cq[i] = __ldg(cf + i);
                    cst = cq[0];
                    if constexpr (NCT > 0) {
                        real full[NCT];                     // values of the point-evidence arguments (entries NC ..)
                        int p = 1 + NCT;
#pragma unroll
                        for (int i = 0; i < NC; ++i) lin[i] = cq[1 + i];
#pragma unroll
                        for (int e = 0; e < NE; ++e) { full[NC + e] = ev[e]; cst += cq[1 + NC + e] * ev[e]; }
#pragma unroll
                        for (int i = 0; i < NCT; ++i) {
#pragma unroll
                            for (int j = i; j < NCT; ++j) {
                                const real a_ij = cq[p++];
                                if (i < NC && j < NC) A[i < NCS ? i : 0][j < NCS ? j : 0] = a_ij;
                                else if (i < NC) lin[i < NCS ? i : 0] += a_ij * full[j];
                                else cst += a_ij * full[i] * full[j];
                            }
                        }
                    }
                }

                constexpr int T0 = NC >= 2 ? T : 1, T1 = NC >= 1 ? T : 1;
                constexpr int kOuterUnroll = T <= 4 ? T0 : 1;
#pragma unroll kOuterUnroll
                for (int t0 = 0; t0 < T0; ++t0) {
                    // outer axis (argument 0 of two): node, weight, cross densities, partial quadratic
                    real xv0 = real(0), W0 = wd, qv0 = cst, lin1 = NC >= 2 ? lin[NCS - 1] : real(0);
                    real pk0[K];
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk0[k2] = pk[k2];
                    if constexpr (NC >= 2) {
                        xv0 = sd0 * s_quad[t0] + mu[0][k];
                        W0 *= s_quad[T + t0];
                        if constexpr (FL != kNode) {
                            qv0 += xv0 * (lin[0] + A[0][0] * xv0);
                            lin1 += A[0][NCS - 1] * xv0;
                        }
                        if constexpr (FL != kPure) {
#pragma unroll
                            for (int k2 = 0; k2 < K; ++k2) pk0[k2] *= norm_pdf<real>(xv0, mu[0][k2], var[0][k2]);
                        }
                    }
                    real S0 = real(0);                       // sum over the inner axis of omega F
#pragma unroll
                    for (int t1 = 0; t1 < T1; ++t1) {
                        real W = W0, qv = qv0, xv1 = real(0);
                        if constexpr (NC >= 1) {
                            W *= om[t1];
                            xv1 = x[t1];
                            if constexpr (FL != kNode) qv += xv1 * ((NC >= 2 ? lin1 : lin[0]) + A[CL][CL] * xv1);
                        }
                        real lb = real(0);
                        if constexpr (FL != kPure) {
                            real b = real(0);
#pragma unroll
                            for (int k2 = 0; k2 < K; ++k2) b += NC >= 1 ? pk0[k2] * q[k2][t1] : pk0[k2];
                            if (M::belief_underflow(b)) {
                                HybPoint<real, K, ND, NC> pt;
#pragma unroll
                                for (int k2 = 0; k2 < K; ++k2) {
                                    pt.w[k2] = wk[k2];
#pragma unroll
                                    for (int a = 0; a < ND; ++a) pt.pd[a][k2] = pd[a][k2][dg[a]];
#pragma unroll
                                    for (int c = 0; c < NC; ++c) { pt.mu[c][k2] = mu[c][k2]; pt.var[c][k2] = var[c][k2]; }
                                }
                                if constexpr (NC >= 2) pt.x[0] = xv0;
                                if constexpr (NC >= 1) pt.x[CL] = xv1;
                                lb = hyb_log_belief_double<real, K, ND, NC>(pt);
                            } else {
                                lb = M::log_belief(b);
                            }
                        }
                        real F;
                        if constexpr (FL == kNode) F = lb;
                        else if constexpr (FL == kPure) F = (NCT == 0 ? cst : M::log_psi(qv));
                        else F = (NCT == 0 ? cst : M::log_psi(qv)) - lb;
                        const real S = W * F;
                        Ek += S;
                        if constexpr (NC >= 1) { const real dx = xv1 - mu[CL][k]; am[CL] += S * dx; av[CL] += S * (dx * dx - var[CL][k]); }
                        if constexpr (NC >= 2) { const real dx = xv0 - mu[0][k]; am[0] += S * dx; av[0] += S * (dx * dx - var[0][k]); }
                        if constexpr (ND > 0) S0 += (NC >= 1 ? om[t1] : real(1)) * F;
                    }
                    // categorical gradients: the weight of every OTHER axis times F
                    if constexpr (ND > 0) {
                        const real Wc = NC >= 2 ? S0 * s_quad[T + t0] : S0;
#pragma unroll
                        for (int a = 0; a < ND; ++a) {
                            real Wo = Wc;
#pragma unroll
                            for (int a2 = 0; a2 < ND; ++a2)
                                if (a2 != a) Wo *= pd[a2][k][dg[a2]];
                            gd[a][k][dg[a]] -= Wo;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const real inv = M::rcp(var[c][k]);
                gc[c][2 * k] = -am[c] * inv;
                gc[c][2 * k + 1] = -av[c] * (real(0.5) * inv * inv);
            }
            acc[k] -= (double)(wf * Ek);
            acc[K] -= (double)(wf * wk[k] * Ek);
        }

        // ---- scatter the parameter gradients
        const real nscale = FL == kNode ? g.nscale[r] : real(1);
#pragma unroll
        for (int a = 0; a < ND; ++a) {
            if (g.no_cat) continue;                      // compat="reference": lhvi_category_grad_reference
            const real gam = ((weighted && FL != kNode) ? g.gam[a * g.n + r] : real(1)) * nscale;
            if (gam == real(0)) continue;
            real v[K * D];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int d = 0; d < D; ++d) v[k * D + d] = gam * gd[a][k][d];
            hyb_scatter<K * D, K, real>(g.grad + off[a], v);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const real gam = ((weighted && FL != kNode) ? g.gam[(ND + c) * g.n + r] : real(1)) * nscale;
            if (gam == real(0)) continue;
            real v[2 * K];
#pragma unroll
            for (int i = 0; i < 2 * K; ++i) v[i] = gam * gc[c][i];
            hyb_scatter<2 * K, K, real>(g.grad + off[ND + c], v);
        }
    }

    publish_partials(acc, K + 1, s_scratch, g.partials);
}

template <typename real, int K, int T, int ND, int D, int NC, int NE, int FL>
static int launch_hyb_one(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const GroupView<real> v = make_view<real>(m, g, row0);
    long long blocks = (g->n + kHybThreads - 1) / kHybThreads;
    if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
    if (blocks < 1) blocks = 1;
    factor_hyb_kernel<real, K, T, ND, D, NC, NE, FL><<<(unsigned)blocks, kHybThreads, 0, s>>>(v);
    return check_launch("factor_hyb_kernel");
}

// returns 1 when there is no kernel for the group
template <typename real, int K>
int launch_hyb_k(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    if (!hyb_available(m, g)) return 1;
    const int code = hyb_code(g->nd, g->nc, g->ne, hyb_flavour(g));
    const int D = hyb_states(g);
    if (g->nd > 0 && g->nc == 0) {
        switch (code) {
#define X(ND_, NC_, NE_, FL_)                                                                     \
            case hyb_code(ND_, NC_, NE_, FL_):                                                    \
                if (D == 2) return launch_hyb_one<real, K, 1, ND_, 2, NC_, NE_, FL_>(m, g, row0, s); \
                if constexpr (ND_ <= 3) return launch_hyb_one<real, K, 1, ND_, 3, NC_, NE_, FL_>(m, g, row0, s); \
                return 1;
            LHVI_HYB_DISCRETE(X)
#undef X
            default: return 1;
        }
    }
    if (g->nd > 0) {
        switch (code) {
#define X(ND_, NC_, NE_, FL_)                                                                     \
            case hyb_code(ND_, NC_, NE_, FL_):                                                    \
                if (m->T == 3) {                                                                  \
                    if (D == 2) return launch_hyb_one<real, K, 3, ND_, 2, NC_, NE_, FL_>(m, g, row0, s); \
                    if constexpr (ND_ + NC_ <= 3) return launch_hyb_one<real, K, 3, ND_, 3, NC_, NE_, FL_>(m, g, row0, s); \
                    return 1;                                                                     \
                }                                                                                 \
                if constexpr (K <= 2) {                                                           \
                    if (D == 2) return launch_hyb_one<real, K, 10, ND_, 2, NC_, NE_, FL_>(m, g, row0, s); \
                    if constexpr (ND_ + NC_ <= 2) return launch_hyb_one<real, K, 10, ND_, 3, NC_, NE_, FL_>(m, g, row0, s); \
                }                                                                                 \
                return 1;
            LHVI_HYB_MIXED(X)
#undef X
            default: return 1;
        }
    }
    if (g->nc == 0) {
        switch (code) {
#define X(ND_, NC_, NE_, FL_) \
            case hyb_code(ND_, NC_, NE_, FL_): return launch_hyb_one<real, K, 1, 0, 2, NC_, NE_, FL_>(m, g, row0, s);
            LHVI_HYB_CONST(X)
#undef X
            default: return 1;
        }
    }
    if constexpr (K <= 2) {
        switch (code) {
#define X(ND_, NC_, NE_, FL_) \
            case hyb_code(ND_, NC_, NE_, FL_): return launch_hyb_one<real, K, 10, 0, 2, NC_, NE_, FL_>(m, g, row0, s);
            LHVI_HYB_CONT(X)
#undef X
            default: return 1;
        }
    }
    return 1;
}

}  // namespace lhvi
