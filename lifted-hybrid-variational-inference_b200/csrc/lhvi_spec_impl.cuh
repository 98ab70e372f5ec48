// Template-specialised factor kernels: compile-time K / T / argument counts, every table in
// registers, hierarchical gradient reduction.
//
// Scope: record groups without hidden discrete arguments (after evidence folding this covers
// the Gaussian, linear-Gaussian and hybrid-MLN-with-observed-relations potentials, i.e. all the
// large groups of BASELINE configs 2-5) with K <= 3 and T = 3.  Anything else runs in
// lhvi_generic.cu.
//
// Per record (one thread): the quadratic log-potential is reduced by the point evidence once,
// then for each mixture component k the thread tabulates nodes and cross-densities of every
// axis in registers and walks the T^NA grid with a compile-time recursion; F is evaluated once
// per grid point.  exp(-xi_t^2) replaces the own-component density (x - mu_k = sqrt(2 var) xi_t),
// which removes 1/K of the exponentials.
//
// Gradient scatter: warp-level segmented reduction over runs of equal parameter offsets, then
// per argument either vector REDs straight to global memory (distinct variables per lane) or a
// per-block shared-memory accumulator cache flushed once at the end (hub variables shared by
// long runs of records; selected by lhvi_group::hub_mask).
#pragma once
#include "lhvi_common.cuh"

namespace lhvi {

constexpr int kSpecThreads = 256;
constexpr int kCacheSlots = 64;

// ---- vector helpers ------------------------------------------------------------------------

template <int N>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[N]) {
    if constexpr (N == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < (N + 3) / 4; ++c) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c);
            if (4 * c + 0 < N) v[4 * c + 0] = t.x;
            if (4 * c + 1 < N) v[4 * c + 1] = t.y;
            if (4 * c + 2 < N) v[4 * c + 2] = t.z;
            if (4 * c + 3 < N) v[4 * c + 3] = t.w;
        }
    }
}

template <int N>
__device__ __forceinline__ void load_vec(const double* __restrict__ p, double (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N / 2; ++c) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p) + c);
        v[2 * c] = t.x; v[2 * c + 1] = t.y;
    }
}

template <int N>
__device__ __forceinline__ void red_vec(float* p, const float (&v)[N]) {
    if constexpr (N == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (int c = 0; c < (N + 3) / 4; ++c) {
            float4 t;
            t.x = 4 * c + 0 < N ? v[4 * c + 0] : 0.f;
            t.y = 4 * c + 1 < N ? v[4 * c + 1] : 0.f;
            t.z = 4 * c + 2 < N ? v[4 * c + 2] : 0.f;
            t.w = 4 * c + 3 < N ? v[4 * c + 3] : 0.f;
            atomicAdd(reinterpret_cast<float4*>(p) + c, t);     // REDG.E.ADD.F32x4
        }
    }
}

template <int N>
__device__ __forceinline__ void red_vec(double* p, const double (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) atomicAdd(p + i, v[i]);
}

// ---- rare float path: belief underflow, redo one grid point in double ------------------------

template <typename real, int NC, int NG>
struct PointCtx {
    int poff[NC > 0 ? NC : 1];
    real x[NC + NG > 0 ? NC + NG : 1];
    real egval[NG > 0 ? NG : 1], egvar[NG > 0 ? NG : 1];
};

template <typename real, int K, int NC, int NG>
__device__ __noinline__ real slow_log_belief(const real* __restrict__ eta, const real* s_w,
                                              const PointCtx<real, NC, NG> c) {
    double b = 0.0;
    for (int k2 = 0; k2 < K; ++k2) {
        double p = (double)s_w[k2];
        for (int a = 0; a < NC; ++a)
            p *= norm_pdf_d((double)c.x[a], (double)eta[c.poff[a] + 2 * k2], (double)eta[c.poff[a] + 2 * k2 + 1]);
        for (int j = 0; j < NG; ++j) p *= norm_pdf_d((double)c.x[NC + j], (double)c.egval[j], (double)c.egvar[j]);
        b += p;
    }
    return (real)::log(b + kEps);
}

// ---- compile-time grid walk ------------------------------------------------------------------

template <typename real, int K, int T, int NC, int NG, int NE, bool NODE>
struct Ctx {
    static constexpr int NA = NC + NG;
    static constexpr int NCT = NC + NG + NE;
    static constexpr int NQ = NCT > 0 ? NCT : 1;
    real x[NA > 0 ? NA : 1][T];            // node positions under the current component
    real q[NA > 0 ? NA : 1][K][T];         // cross densities q_{k'}(x_t)
    real qw[T], xi[T];
    real A[NQ][NQ];                        // upper-triangular quadratic coefficients
    real nscale;
    real m1[NC > 0 ? NC : 1], m2[NC > 0 ? NC : 1];   // sum W F xi_t, sum W F xi_t^2 per hidden axis
    const real* eta;
    const real* s_w;
    PointCtx<real, NC, NG> pt;
};

template <typename real, int K, int T, int NC, int NG, int NE, bool NODE, int AX>
struct Walk {
    using C = Ctx<real, K, T, NC, NG, NE, NODE>;
    static __device__ __forceinline__ real run(C& c, const real (&pk)[K], real Wout, real cst,
                                               const real (&lin)[C::NQ]) {
        if constexpr (AX == C::NA) {
            real b = pk[0];
#pragma unroll
            for (int k2 = 1; k2 < K; ++k2) b += pk[k2];
            real lb;
            if (Math<real>::belief_underflow(b)) lb = slow_log_belief<real, K, NC, NG>(c.eta, c.s_w, c.pt);
            else lb = Math<real>::log_belief(b);
            if constexpr (NODE) return c.nscale * lb;
            else return Math<real>::log_psi(cst) - lb;
        } else {
            real ret = real(0);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                real pk2[K];
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2) pk2[k2] = pk[k2] * c.q[AX][k2][t];
                const real xv = c.x[AX][t];
                real cst2 = cst;
                real lin2[C::NQ];
#pragma unroll
                for (int j = 0; j < C::NQ; ++j) lin2[j] = lin[j];
                if constexpr (!NODE) {
                    cst2 = cst + xv * (lin[AX] + c.A[AX][AX] * xv);
#pragma unroll
                    for (int j = AX + 1; j < C::NA; ++j) lin2[j] = lin[j] + c.A[AX][j] * xv;
                }
                c.pt.x[AX] = xv;
                const real R = Walk<real, K, T, NC, NG, NE, NODE, AX + 1>::run(c, pk2, Wout * c.qw[t], cst2, lin2);
                const real wr = c.qw[t] * R;
                ret += wr;
                if constexpr (AX < NC) {
                    const real m = Wout * wr * c.xi[t];
                    c.m1[AX] += m;
                    c.m2[AX] += m * c.xi[t];
                }
            }
            return ret;
        }
    }
};

struct SpecLaunch {
    long long chunk;      // records per block (multiple of the block size)
    int hub_mask;         // bit a: hidden argument a accumulates through the shared-memory cache
};

template <typename real, int K, int T, int NC, int NG, int NE, bool NODE, bool WEIGHTED>
__global__ void __launch_bounds__(kSpecThreads, 2)
factor_spec_kernel(const GroupView<real> g, const SpecLaunch L) {
    using M = Math<real>;
    using C = Ctx<real, K, T, NC, NG, NE, NODE>;
    constexpr int NA = C::NA, NCT = C::NCT, NQ = C::NQ, NV = 2 * K;
    constexpr int NCS = NC > 0 ? NC : 1;

    __shared__ real s_quad[2 * T];
    __shared__ real s_eq[T];
    __shared__ real s_w[K];
    __shared__ int s_tag[NCS][kCacheSlots];
    __shared__ real s_val[NCS][kCacheSlots][NV];
    __shared__ double s_scratch[(kSpecThreads / 32) * (K + 1)];

    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_eq[i] = M::exp(-g.quad[i] * g.quad[i]);
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    for (int i = threadIdx.x; i < NCS * kCacheSlots; i += blockDim.x) (&s_tag[0][0])[i] = -1;
    for (int i = threadIdx.x; i < NCS * kCacheSlots * NV; i += blockDim.x) (&s_val[0][0][0])[i] = real(0);
    __syncthreads();

    C c;
    real eq[T], wk[K];
#pragma unroll
    for (int t = 0; t < T; ++t) { c.xi[t] = s_quad[t]; c.qw[t] = s_quad[T + t]; eq[t] = s_eq[t]; }
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = s_w[k];
    c.eta = g.eta;
    c.s_w = s_w;

    double acc[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;

    const int lane = threadIdx.x & 31;
    const long long lo = (long long)blockIdx.x * L.chunk;
    const long long hi = lo + L.chunk < g.n ? lo + L.chunk : g.n;

    for (long long base = lo; base < hi; base += blockDim.x) {
        const long long r = base + threadIdx.x;
        const bool active = r < hi;
        int key[NCS];
        real gv[NCS][NV];
#pragma unroll
        for (int a = 0; a < NCS; ++a) {
            key[a] = -1;
#pragma unroll
            for (int i = 0; i < NV; ++i) gv[a][i] = real(0);
        }

        if (active) {
            const real wf = WEIGHTED ? __ldcs(g.wf + r) : real(1);
            real mu[NCS][K], var[NCS][K], hvar[NCS][K], nrm[NCS][K];
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                key[a] = __ldcs(g.poff + a * g.n + r);
                c.pt.poff[a] = key[a];
                real slot[NV];
                load_vec<NV>(g.eta + key[a], slot);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    mu[a][k] = slot[2 * k];
                    var[a][k] = slot[2 * k + 1];
                    const real inv = M::rcp(var[a][k]);
                    hvar[a][k] = real(-0.5) * inv;
                    nrm[a][k] = inv * real(1.0 / kSqrt2Pi);
                }
            }
            real egval[NG > 0 ? NG : 1], egs[NG > 0 ? NG : 1], egh[NG > 0 ? NG : 1], egn[NG > 0 ? NG : 1];
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                egval[j] = __ldcs(g.egval + j * g.n + r);
                const real v = __ldcs(g.egvar + j * g.n + r);
                c.pt.egval[j] = egval[j];
                c.pt.egvar[j] = v;
                egs[j] = M::sqrt(real(2) * v);
                const real inv = M::rcp(v);
                egh[j] = real(-0.5) * inv;
                egn[j] = inv * real(1.0 / kSqrt2Pi);
            }

            // quadratic log-potential, reduced by the point evidence
            real cst0 = real(0), lin0[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) lin0[j] = real(0);
            if constexpr (NODE) {
                c.nscale = __ldcs(g.nscale + r);
            } else {
                const real* cf = g.ptab + __ldcs(g.pot + r);
                cst0 = __ldg(cf);
                if constexpr (NCT > 0) {
#pragma unroll
                    for (int i = 0; i < NCT; ++i) lin0[i] = __ldg(cf + 1 + i);
                    int p = 1 + NCT;
#pragma unroll
                    for (int i = 0; i < NCT; ++i)
#pragma unroll
                        for (int j = i; j < NCT; ++j) c.A[i][j] = __ldg(cf + p++);
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const int j = NA + e;
                        const real xv = __ldcs(g.ecval + e * g.n + r);
                        cst0 += xv * (lin0[j] + c.A[j][j] * xv);
#pragma unroll
                        for (int i = 0; i < j; ++i) lin0[i] += c.A[i][j] * xv;
#pragma unroll
                        for (int i = j + 1; i < NCT; ++i) lin0[i] += c.A[j][i] * xv;
                    }
                }
            }

#pragma unroll
            for (int k = 0; k < K; ++k) {
                // tabulate the axes under component k
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    const real s = M::sqrt(real(2) * var[a][k]);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real dx = s * c.xi[t];
                        c.x[a][t] = dx + mu[a][k];
#pragma unroll
                        for (int k2 = 0; k2 < K; ++k2) {
                            if (k2 == k) {
                                c.q[a][k2][t] = eq[t] * nrm[a][k2];      // exp(-xi^2) / (sqrt(2pi) var)
                            } else {
                                const real u = dx + (mu[a][k] - mu[a][k2]);
                                c.q[a][k2][t] = M::exp(hvar[a][k2] * u * u) * nrm[a][k2];
                            }
                        }
                    }
                    c.m1[a] = real(0);
                    c.m2[a] = real(0);
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) {
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        c.x[NC + j][t] = egs[j] * c.xi[t] + egval[j];
                        const real qe = eq[t] * egn[j];
#pragma unroll
                        for (int k2 = 0; k2 < K; ++k2) c.q[NC + j][k2][t] = qe;
                    }
                }
                real pk[K];
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2) pk[k2] = wk[k2];
                const real Ek = Walk<real, K, T, NC, NG, NE, NODE, 0>::run(c, pk, real(1), cst0, lin0);

#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    // g_mu = -sum W F (x-mu) / var ; g_var = -sum W F ((x-mu)^2 - var) / (2 var^2)
                    const real inv = real(-2) * hvar[a][k];                 // 1 / var
                    const real s = M::sqrt(real(2) * var[a][k]);
                    gv[a][2 * k] = -(s * c.m1[a]) * inv;
                    gv[a][2 * k + 1] = -(c.m2[a] - real(0.5) * Ek) * inv;
                }
                acc[k] -= (double)(wf * Ek);
                acc[K] -= (double)(wf * wk[k] * Ek);
            }
            if constexpr (WEIGHTED) {
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    const real gam = __ldcs(g.gam + a * g.n + r);
#pragma unroll
                    for (int i = 0; i < NV; ++i) gv[a][i] *= gam;
                }
            }
        }

        // ---- scatter: warp segmented reduction over runs of equal offsets, then RED / cache
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            const int prev = __shfl_up_sync(0xffffffffu, key[a], 1);
            const bool head = lane == 0 || prev != key[a];
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            bool writer;
            if (heads == 1u) {                       // one run: butterfly
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    real v = gv[a][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    gv[a][i] = v;
                }
                writer = lane == 0;
            } else if (heads == 0xffffffffu) {       // all distinct
                writer = true;
            } else {                                 // segmented inclusive scan, tails write
                const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int seg_up = __shfl_up_sync(0xffffffffu, seg, o);
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const real up = __shfl_up_sync(0xffffffffu, gv[a][i], o);
                        if (lane >= o && seg_up == seg) gv[a][i] += up;
                    }
                }
                writer = lane == 31 || ((heads >> (lane + 1)) & 1u);
            }
            if (writer && key[a] >= 0) {
                bool done = false;
                if ((L.hub_mask >> a) & 1) {
                    const int slot = (key[a] / (NV <= 2 ? 2 : ((NV + 3) / 4) * 4)) & (kCacheSlots - 1);
                    const int old = atomicCAS(&s_tag[a][slot], -1, key[a]);
                    if (old == -1 || old == key[a]) {
#pragma unroll
                        for (int i = 0; i < NV; ++i) atomicAdd(&s_val[a][slot][i], gv[a][i]);
                        done = true;
                    }
                }
                if (!done) red_vec<NV>(g.grad + key[a], gv[a]);
            }
        }
    }

    block_sum_to(acc, K + 1, s_scratch, g.partials + (long long)blockIdx.x * (K + 1));

    // flush the hub cache (block_sum_to ends with a barrier, so every shared atomic has landed)
    if constexpr (NC > 0) {
        for (int i = threadIdx.x; i < NC * kCacheSlots; i += blockDim.x) {
            const int a = i / kCacheSlots, slot = i % kCacheSlots;
            const int tag = s_tag[a][slot];
            if (tag >= 0) {
                real v[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) v[j] = s_val[a][slot][j];
                red_vec<NV>(g.grad + tag, v);
            }
        }
    }
}

// ---- dispatch ----------------------------------------------------------------------------------

template <typename real, int K, int T, int NC, int NG, int NE, bool NODE>
static int launch_one(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    GroupView<real> v = make_view<real>(m, g, row0);
    long long blocks = (g->n + kSpecThreads - 1) / kSpecThreads;
    const long long persistent = 148 * 2;
    if (blocks > persistent) blocks = persistent;
    if (blocks > LHVI_PARTIAL_ROWS) blocks = LHVI_PARTIAL_ROWS;
    SpecLaunch L;
    L.chunk = ((g->n + blocks - 1) / blocks + kSpecThreads - 1) / kSpecThreads * kSpecThreads;
    blocks = (g->n + L.chunk - 1) / L.chunk;
    L.hub_mask = g->hub_mask >> g->nd;
    if (g->weighted)
        factor_spec_kernel<real, K, T, NC, NG, NE, NODE, true><<<(unsigned)blocks, kSpecThreads, 0, s>>>(v, L);
    else
        factor_spec_kernel<real, K, T, NC, NG, NE, NODE, false><<<(unsigned)blocks, kSpecThreads, 0, s>>>(v, L);
    return check_launch("factor_spec_kernel");
}

// signature code = NC*100 + NG*10 + NE (ND == 0 only)
#define LHVI_SPEC_SIGS(X) X(1, 0, 0) X(2, 0, 0) X(1, 0, 1) X(1, 1, 0) X(2, 0, 1) X(2, 1, 0) X(3, 0, 0) X(1, 0, 2)

template <typename real, int K, int T>
int launch_kt(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    if (g->node) {
        if (g->nc == 1 && g->ng == 0) return launch_one<real, K, T, 1, 0, 0, true>(m, g, row0, s);
        if (g->nc == 0 && g->ng == 1) return launch_one<real, K, T, 0, 1, 0, true>(m, g, row0, s);
        return 1;
    }
    const int code = g->nc * 100 + g->ng * 10 + g->ne;
    switch (code) {
#define X(NC_, NG_, NE_) case NC_ * 100 + NG_ * 10 + NE_: return launch_one<real, K, T, NC_, NG_, NE_, false>(m, g, row0, s);
        LHVI_SPEC_SIGS(X)
#undef X
        default: return 1;
    }
}

}  // namespace lhvi
