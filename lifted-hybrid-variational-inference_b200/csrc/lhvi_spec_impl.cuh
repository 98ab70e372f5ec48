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
// which removes 1/K of the exponentials.  Three record flavours share the code:
//   full   F = log psi - log b      (two or more integrated arguments)
//   pure   F = log psi              (unary split, see lowering.py: no density, exp or log at all)
//   node   F = log b                (the variables' entropy terms)
//
// The hot walk is branch-free: it assumes no floor is active (log(psi+1e-100) = log psi,
// log(b+1e-100) = log b, no float underflow) while tracking min(q) and min(b); if a bound
// trips, that (record, component) is redone by a checked walk that evaluates the reference's
// expressions literally (in double where float cannot).
//
// Gradient scatter: at most one argument per group is declared a hub (lhvi_group::hub_mask);
// each thread keeps a running accumulator for it across its records and flushes through a
// per-block shared-memory cache, so a hub costs no shuffles and O(blocks) global atomics.  Other
// arguments use a warp-level segmented reduction over runs of equal offsets followed by vector
// REDs (REDG.E.ADD.F32x4) to global memory.
#pragma once
#include "lhvi_common.cuh"
#include "lhvi_spec_sigs.h"

namespace lhvi {

constexpr int kSpecThreads = 256;
constexpr int kCacheSlots = 32;

// ---- scalar math in the form the hot loop wants ------------------------------------------------

template <typename real> struct Fast;

template <> struct Fast<float> {
    static constexpr float kExpScale = 1.4426950408889634f;     // exponents are kept in log2 units
    static __device__ __forceinline__ float exp_scaled(float x) {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float log_belief(float b) {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b));
        return y * 0.6931471805599453f;
    }
    static constexpr float kQFloor = -80.0f;      // below: exp(q) is within 1e-35 of the 1e-100 floor scale
    static constexpr float kBFloor = 1e-30f;      // below: float products may have flushed to zero
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float sqrt(float x) { return __fsqrt_rn(x); }
};

template <> struct Fast<double> {
    static constexpr double kExpScale = 1.0;
    static __device__ __forceinline__ double exp_scaled(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log_belief(double b) { return ::log(b + kEps); }
    static constexpr double kQFloor = -180.0;     // below: log(exp(q)+1e-100) differs from q in double
    static constexpr double kBFloor = -1.0;       // never trips: log(b + 1e-100) is always literal
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
};

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

// ---- checked leaf: the reference's expressions, literally ---------------------------------------

template <typename real, int NC, int NG>
struct PointCtx {
    int poff[NC > 0 ? NC : 1];
    real x[NC + NG > 0 ? NC + NG : 1];
    real egval[NG > 0 ? NG : 1], egvar[NG > 0 ? NG : 1];
};

// log(b + 1e-100) with the belief recomputed in double from the parameters
template <typename real, int K, int NC, int NG>
__device__ __noinline__ real checked_log_belief(const real* __restrict__ eta, const real* s_w,
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

// log(exp(q) + 1e-100)
template <typename real>
__device__ __noinline__ real checked_log_psi(real q) {
    return (real)::log(::exp((double)q) + kEps);
}

// ---- compile-time grid walk ------------------------------------------------------------------

enum Flavour { kFull = 0, kPure = 1, kNode = 2 };

template <typename real, int K, int T, int NC, int NG, int NE>
struct Ctx {
    static constexpr int NA = NC + NG;
    static constexpr int NCT = NC + NG + NE;
    static constexpr int NQ = NCT > 0 ? NCT : 1;
    real x[NA > 0 ? NA : 1][T];            // node positions under the current component
    real q[NA > 0 ? NA : 1][K][T];         // cross densities q_{k'}(x_t)
    real qw[T], xi[T];
    real A[NQ][NQ];                        // upper-triangular quadratic coefficients
    real m1[NC > 0 ? NC : 1], m2[NC > 0 ? NC : 1];   // sum W F xi_t, sum W F xi_t^2 per hidden axis
    real qmin, bmin;                       // bounds tracked by the unchecked walk
    const real* eta;
    const real* s_w;
    PointCtx<real, NC, NG> pt;
};

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool CHECKED, int AX>
struct Walk {
    using C = Ctx<real, K, T, NC, NG, NE>;
    using F = Fast<real>;
    static __device__ __forceinline__ real run(C& c, const real (&pk)[K], real Wout, real cst,
                                               const real (&lin)[C::NQ]) {
        if constexpr (AX == C::NA) {
            real lpsi = cst, lb = real(0);
            if constexpr (FL != kNode) {
                if constexpr (CHECKED) lpsi = checked_log_psi<real>(cst);
                else c.qmin = cst < c.qmin ? cst : c.qmin;
            }
            if constexpr (FL != kPure) {
                real b = pk[0];
#pragma unroll
                for (int k2 = 1; k2 < K; ++k2) b += pk[k2];
                if constexpr (CHECKED) {
                    lb = checked_log_belief<real, K, NC, NG>(c.eta, c.s_w, c.pt);
                } else {
                    c.bmin = b < c.bmin ? b : c.bmin;
                    lb = F::log_belief(b);
                }
            }
            if constexpr (FL == kNode) return lb;
            else if constexpr (FL == kPure) return lpsi;
            else return lpsi - lb;
        } else {
            real ret = real(0);
#pragma unroll
            for (int t = 0; t < T; ++t) {
                real pk2[K];
                if constexpr (FL != kPure) {
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk2[k2] = pk[k2] * c.q[AX][k2][t];
                } else {
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk2[k2] = pk[k2];
                }
                const real xv = c.x[AX][t];
                real cst2 = cst;
                real lin2[C::NQ];
#pragma unroll
                for (int j = 0; j < C::NQ; ++j) lin2[j] = lin[j];
                if constexpr (FL != kNode) {
                    cst2 = cst + xv * (lin[AX] + c.A[AX][AX] * xv);
#pragma unroll
                    for (int j = AX + 1; j < C::NA; ++j) lin2[j] = lin[j] + c.A[AX][j] * xv;
                }
                if constexpr (CHECKED) c.pt.x[AX] = xv;
                const real R = Walk<real, K, T, NC, NG, NE, FL, CHECKED, AX + 1>::run(c, pk2, Wout * c.qw[t], cst2, lin2);
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
};

// HUB: index of the hidden argument accumulated per thread across records (-1: none)
template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool WEIGHTED, int HUB>
__global__ void __launch_bounds__(kSpecThreads, (FL == kFull && NC + NG >= 2) ? 2 : 3)
factor_spec_kernel(const GroupView<real> g, const SpecLaunch L) {
    using F = Fast<real>;
    using C = Ctx<real, K, T, NC, NG, NE>;
    constexpr int NA = C::NA, NCT = C::NCT, NQ = C::NQ, NV = 2 * K;
    constexpr int NCS = NC > 0 ? NC : 1;
    constexpr int kSlotElems = NV <= 2 ? 2 : ((NV + 3) / 4) * 4;

    __shared__ real s_quad[2 * T];
    __shared__ real s_eq[T];
    __shared__ real s_w[K];
    __shared__ int s_tag[kCacheSlots];
    __shared__ real s_val[kCacheSlots][NV];
    __shared__ double s_scratch[(kSpecThreads / 32) * (K + 1)];

    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_eq[i] = (real)::exp(-(double)g.quad[i] * (double)g.quad[i]);
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    for (int i = threadIdx.x; i < kCacheSlots; i += blockDim.x) s_tag[i] = -1;
    for (int i = threadIdx.x; i < kCacheSlots * NV; i += blockDim.x) (&s_val[0][0])[i] = real(0);
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

    // running accumulator of the hub argument
    int hub_key = -1;
    real hub_acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hub_acc[i] = real(0);

    // one thread's partial sums for a hub variable -> shared cache (or global on a slot clash)
    auto flush_hub = [&](int key, const real (&v)[NV]) {
        const int slot = (key / kSlotElems) & (kCacheSlots - 1);
        const int old = atomicCAS(&s_tag[slot], -1, key);
        if (old == -1 || old == key) {
#pragma unroll
            for (int i = 0; i < NV; ++i) atomicAdd(&s_val[slot][i], v[i]);
        } else {
            red_vec<NV>(g.grad + key, v);
        }
    };

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
            const real wf = (WEIGHTED || FL == kNode) ? __ldcs(g.wf + r) : real(1);
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
                    if constexpr (FL != kPure) {
                        const real inv = F::rcp(var[a][k]);
                        hvar[a][k] = real(-0.5) * F::kExpScale * inv;
                        nrm[a][k] = inv * real(1.0 / kSqrt2Pi);
                    }
                }
            }
            real egval[NG > 0 ? NG : 1], egs[NG > 0 ? NG : 1], egn[NG > 0 ? NG : 1];
#pragma unroll
            for (int j = 0; j < NG; ++j) {
                egval[j] = __ldcs(g.egval + j * g.n + r);
                const real v = __ldcs(g.egvar + j * g.n + r);
                c.pt.egval[j] = egval[j];
                c.pt.egvar[j] = v;
                egs[j] = F::sqrt(real(2) * v);
                egn[j] = F::rcp(v) * real(1.0 / kSqrt2Pi);
            }

            // quadratic log-potential, reduced by the point evidence
            real cst0 = real(0), lin0[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) lin0[j] = real(0);
            if constexpr (FL != kNode) {
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

            real Eks[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // tabulate the axes under component k
                real sdev[NCS];
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    sdev[a] = F::sqrt(real(2) * var[a][k]);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real dx = sdev[a] * c.xi[t];
                        c.x[a][t] = dx + mu[a][k];
                        if constexpr (FL != kPure) {
#pragma unroll
                            for (int k2 = 0; k2 < K; ++k2) {
                                if (k2 == k) {
                                    c.q[a][k2][t] = eq[t] * nrm[a][k2];      // exp(-xi^2) / (sqrt(2pi) var)
                                } else {
                                    const real u = dx + (mu[a][k] - mu[a][k2]);
                                    c.q[a][k2][t] = F::exp_scaled(hvar[a][k2] * (u * u)) * nrm[a][k2];
                                }
                            }
                        }
                    }
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

#pragma unroll
                for (int a = 0; a < NC; ++a) { c.m1[a] = real(0); c.m2[a] = real(0); }
                c.qmin = real(0);
                c.bmin = real(1);
                real Ek = Walk<real, K, T, NC, NG, NE, FL, false, 0>::run(c, pk, real(1), cst0, lin0);
                if (c.qmin < F::kQFloor || c.bmin < F::kBFloor) {
                    // a floor is active somewhere on this grid: redo it with the literal formulas
#pragma unroll
                    for (int a = 0; a < NC; ++a) { c.m1[a] = real(0); c.m2[a] = real(0); }
                    Ek = Walk<real, K, T, NC, NG, NE, FL, true, 0>::run(c, pk, real(1), cst0, lin0);
                }
                Eks[k] = Ek;

#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    // g_mu = -sum W F (x-mu) / var ; g_var = -sum W F ((x-mu)^2 - var) / (2 var^2)
                    const real inv = F::rcp(var[a][k]);
                    gv[a][2 * k] = -(sdev[a] * c.m1[a]) * inv;
                    gv[a][2 * k + 1] = -(c.m2[a] - real(0.5) * Ek) * inv;
                }
            }
            real e_sum = real(0);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                acc[k] -= (double)(wf * Eks[k]);
                e_sum += wk[k] * Eks[k];
            }
            acc[K] -= (double)(wf * e_sum);

            if constexpr (FL == kNode) {
                const real gs = __ldcs(g.nscale + r);
#pragma unroll
                for (int a = 0; a < NC; ++a)
#pragma unroll
                    for (int i = 0; i < NV; ++i) gv[a][i] *= gs;
            } else if constexpr (WEIGHTED) {
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    const real gam = __ldcs(g.gam + a * g.n + r);
#pragma unroll
                    for (int i = 0; i < NV; ++i) gv[a][i] *= gam;
                }
            }
        }

        // ---- scatter
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            if (a == HUB) {
                // per-thread running sum while the hub variable stays the same
                if (active) {
                    if (key[a] != hub_key) {
                        if (hub_key >= 0) flush_hub(hub_key, hub_acc);
                        hub_key = key[a];
#pragma unroll
                        for (int i = 0; i < NV; ++i) hub_acc[i] = real(0);
                    }
#pragma unroll
                    for (int i = 0; i < NV; ++i) hub_acc[i] += gv[a][i];
                }
                continue;
            }
            // warp segmented reduction over runs of equal offsets, then vector REDs
            const int prev = __shfl_up_sync(0xffffffffu, key[a], 1);
            const bool head = lane == 0 || prev != key[a];
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            bool writer;
            if (heads == 0xffffffffu) {              // all distinct (the common case)
                writer = true;
            } else if (heads == 1u) {                // one run: butterfly
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    real v = gv[a][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    gv[a][i] = v;
                }
                writer = lane == 0;
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
            if (writer && key[a] >= 0) red_vec<NV>(g.grad + key[a], gv[a]);
        }
    }

    // ---- end of the block's chunk: flush running sums (warp-combined when the warp agrees)
    if constexpr (HUB >= 0) {
        const int k0 = __shfl_sync(0xffffffffu, hub_key, 0);
        if (__all_sync(0xffffffffu, hub_key == k0)) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                real v = hub_acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                hub_acc[i] = v;
            }
            if (lane == 0 && hub_key >= 0) flush_hub(hub_key, hub_acc);
        } else if (hub_key >= 0) {
            flush_hub(hub_key, hub_acc);
        }
    }

    block_sum_to(acc, K + 1, s_scratch, g.partials + (long long)blockIdx.x * (K + 1));

    // flush the hub cache (block_sum_to ends with a barrier, so every shared atomic has landed)
    if constexpr (HUB >= 0) {
        for (int slot = threadIdx.x; slot < kCacheSlots; slot += blockDim.x) {
            const int tag = s_tag[slot];
            if (tag >= 0) {
                real v[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) v[j] = s_val[slot][j];
                red_vec<NV>(g.grad + tag, v);
            }
        }
    }
}

// ---- dispatch ----------------------------------------------------------------------------------

template <typename KernelT>
static int resident_blocks(KernelT kernel) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSpecThreads, 0);
    return (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
}

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool WEIGHTED, int HUB>
static int launch_final(const GroupView<real>& v, cudaStream_t s) {
    auto kernel = factor_spec_kernel<real, K, T, NC, NG, NE, FL, WEIGHTED, HUB>;
    static int resident = 0;                 // one persistent wave; each block owns a contiguous chunk
    if (resident == 0) resident = resident_blocks(kernel);
    long long blocks = (v.n + kSpecThreads - 1) / kSpecThreads;
    if (blocks > resident) blocks = resident;
    if (blocks > LHVI_PARTIAL_ROWS) blocks = LHVI_PARTIAL_ROWS;
    SpecLaunch L;
    L.chunk = ((v.n + blocks - 1) / blocks + kSpecThreads - 1) / kSpecThreads * kSpecThreads;
    blocks = (v.n + L.chunk - 1) / L.chunk;
    kernel<<<(unsigned)blocks, kSpecThreads, 0, s>>>(v, L);
    return check_launch("factor_spec_kernel");
}

template <typename real, int K, int T, int NC, int NG, int NE, int FL>
static int launch_one(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const GroupView<real> v = make_view<real>(m, g, row0);
    const bool weighted = g->weighted != 0 || FL == kNode;
    // the hub argument: lowest set bit of hub_mask among the hidden continuous arguments
    int hub = -1;
    for (int a = 0; a < NC; ++a)
        if ((g->hub_mask >> (g->nd + a)) & 1) { hub = a; break; }
    if constexpr (NC == 0) {
        return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, -1>(v, s)
                        : launch_final<real, K, T, NC, NG, NE, FL, false, -1>(v, s);
    } else if constexpr (NC == 1) {
        if (hub == 0)
            return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, 0>(v, s)
                            : launch_final<real, K, T, NC, NG, NE, FL, false, 0>(v, s);
        return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, -1>(v, s)
                        : launch_final<real, K, T, NC, NG, NE, FL, false, -1>(v, s);
    } else if constexpr (NC == 2) {
        if (hub == 0)
            return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, 0>(v, s)
                            : launch_final<real, K, T, NC, NG, NE, FL, false, 0>(v, s);
        if (hub == 1)
            return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, 1>(v, s)
                            : launch_final<real, K, T, NC, NG, NE, FL, false, 1>(v, s);
        return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, -1>(v, s)
                        : launch_final<real, K, T, NC, NG, NE, FL, false, -1>(v, s);
    } else {
        return weighted ? launch_final<real, K, T, NC, NG, NE, FL, true, -1>(v, s)
                        : launch_final<real, K, T, NC, NG, NE, FL, false, -1>(v, s);
    }
}


template <typename real, int K, int T>
int launch_kt(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    if (g->node) {
        if (g->nc == 1 && g->ng == 0) return launch_one<real, K, T, 1, 0, 0, kNode>(m, g, row0, s);
        if (g->nc == 0 && g->ng == 1) return launch_one<real, K, T, 0, 1, 0, kNode>(m, g, row0, s);
        return 1;
    }
    if (g->pure) {
        const int code = g->nc * 10 + g->ne;
        if (g->ng != 0) return 1;
        switch (code) {
#define X(NC_, NE_) case NC_ * 10 + NE_: return launch_one<real, K, T, NC_, 0, NE_, kPure>(m, g, row0, s);
            LHVI_SPEC_PURE(X)
#undef X
            default: return 1;
        }
    }
    const int code = g->nc * 100 + g->ng * 10 + g->ne;
    switch (code) {
#define X(NC_, NG_, NE_) case NC_ * 100 + NG_ * 10 + NE_: return launch_one<real, K, T, NC_, NG_, NE_, kFull>(m, g, row0, s);
        LHVI_SPEC_FULL(X)
#undef X
        default: return 1;
    }
}

}  // namespace lhvi
