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
#include <type_traits>

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
    // the walk works in log2 units: F2 = log2(psi) - log2(b); kUnit converts back at the end
    static constexpr float kUnit = 0.6931471805599453f;
    static __device__ __forceinline__ float log_belief(float b) {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(b));
        return y;
    }
    // log(exp(q) + 1e-100) - q = log1p(exp(-230.26 - q)) exceeds half a float ulp of q only for
    // q < -218; above the threshold (in log2 units here) the floor is invisible in float
    static constexpr float kQFloor = -210.0f * 1.4426950408889634f;
    static constexpr float kBFloor = 1e-30f;      // below: float products may have flushed to zero
    // MUFU approximations (about 1 ulp); the IEEE versions cost ~8 instructions and a branch each
    static __device__ __forceinline__ float rcp(float x) {
        float y;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float sqrt(float x) {
        float y;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
};

template <> struct Fast<double> {
    static constexpr double kExpScale = 1.0;
    static constexpr double kUnit = 1.0;
    static __device__ __forceinline__ double exp_scaled(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log_belief(double b) { return ::log(b + kEps); }
    static constexpr double kQFloor = -180.0;     // below: log(exp(q)+1e-100) differs from q in double
    static constexpr double kBFloor = -1.0;       // never trips: log(b + 1e-100) is always literal
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
};

// ---- vector helpers ------------------------------------------------------------------------

// (plain loads, not ld.global.nc: these read parameter slots, which the persistent iteration kernel
// rewrites between two passes of the same launch)
template <int N>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[N]) {
    if constexpr (N == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < (N + 3) / 4; ++c) {
            const float4 t = *(reinterpret_cast<const float4*>(p) + c);
            if (4 * c + 0 < N) v[4 * c + 0] = t.x;
            if (4 * c + 1 < N) v[4 * c + 1] = t.y;
            if (4 * c + 2 < N) v[4 * c + 2] = t.z;
            if (4 * c + 3 < N) v[4 * c + 3] = t.w;
        }
    }
}

template <int N>
__device__ __forceinline__ void load_vec(const double* p, double (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N / 2; ++c) {
        const double2 t = *(reinterpret_cast<const double2*>(p) + c);
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

// log(b + 1e-100) (in the walk's units) with the belief recomputed in double from the parameters
template <typename real, int K, int NC, int NG>
__device__ __noinline__ real checked_log_belief(const real* eta, const real* s_w,
                                                 const PointCtx<real, NC, NG> c) {
    double b = 0.0;
    for (int k2 = 0; k2 < K; ++k2) {
        double p = (double)s_w[k2];
        for (int a = 0; a < NC; ++a)
            p *= norm_pdf_d((double)c.x[a], (double)eta[c.poff[a] + 2 * k2], (double)eta[c.poff[a] + 2 * k2 + 1]);
        for (int j = 0; j < NG; ++j) p *= norm_pdf_d((double)c.x[NC + j], (double)c.egval[j], (double)c.egvar[j]);
        b += p;
    }
    return (real)(::log(b + kEps) / (double)Fast<real>::kUnit);
}

// log(exp(q) + 1e-100), argument and result in the walk's units
template <typename real>
__device__ __noinline__ real checked_log_psi(real q) {
    const double unit = (double)Fast<real>::kUnit;
    return (real)(::log(::exp((double)q * unit) + kEps) / unit);
}

// ---- compile-time grid walk ------------------------------------------------------------------

#ifndef LHVI_FLAVOURS_DEFINED
#define LHVI_FLAVOURS_DEFINED
enum Flavour { kFull = 0, kPure = 1, kNode = 2 };
#endif

template <typename real, int K, int T, int NC, int NG, int NE>
struct Ctx {
    static constexpr int NA = NC + NG;
    static constexpr int NCT = NC + NG + NE;
    static constexpr int NQ = NCT > 0 ? NCT : 1;
    real x[NA > 0 ? NA : 1][T];            // node positions under the current component
    real q[NA > 0 ? NA : 1][K][T];         // cross densities q_{k'}(x_t)
    real w0[T], w1[T], w2[T];              // omega_t, omega_t xi_t, omega_t xi_t^2
    real A[NQ][NQ];                        // upper-triangular quadratic coefficients
    real m1[NC > 0 ? NC : 1], m2[NC > 0 ? NC : 1];   // sum W F xi_t, sum W F xi_t^2 per hidden axis
    real qmin;                             // smallest log psi seen by the unchecked walk
    const real* eta;
    const real* s_w;
    PointCtx<real, NC, NG> pt;
};

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool CHECKED, int AX>
struct Walk {
    using C = Ctx<real, K, T, NC, NG, NE>;
    using F = Fast<real>;
    // the unchecked full walk integrates log b only (log psi has closed-form moments, see the kernel)
    static constexpr bool WITH_PSI = FL == kPure || (FL == kFull && CHECKED);
    // returns sum over the inner axes of (prod inner omega) * F ; Wout = prod of outer omegas
    static __device__ __forceinline__ real run(C& c, const real (&pk)[K], real Wout, real cst,
                                               const real (&lin)[C::NQ]) {
        if constexpr (AX == C::NA) {
            real lpsi = cst, lb = real(0);
            if constexpr (WITH_PSI) {
                if constexpr (CHECKED) lpsi = checked_log_psi<real>(cst);
                else c.qmin = F::min(c.qmin, cst);
            }
            if constexpr (FL != kPure) {
                if constexpr (CHECKED) {
                    lb = checked_log_belief<real, K, NC, NG>(c.eta, c.s_w, c.pt);
                } else {
                    real b = pk[0];
#pragma unroll
                    for (int k2 = 1; k2 < K; ++k2) b += pk[k2];
                    lb = F::log_belief(b);
                }
            }
            if constexpr (FL == kNode) return lb;
            else if constexpr (FL == kPure) return lpsi;
            else if constexpr (CHECKED) return lpsi - lb;
            else return -lb;
        } else {
            real R[T];
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
                real cst2 = cst;
                real lin2[C::NQ];
#pragma unroll
                for (int j = 0; j < C::NQ; ++j) lin2[j] = lin[j];
                if constexpr (WITH_PSI) {
                    const real xv = c.x[AX][t];
                    cst2 = cst + xv * (lin[AX] + c.A[AX][AX] * xv);
#pragma unroll
                    for (int j = AX + 1; j < C::NA; ++j) lin2[j] = lin[j] + c.A[AX][j] * xv;
                }
                if constexpr (CHECKED) c.pt.x[AX] = c.x[AX][t];
                R[t] = Walk<real, K, T, NC, NG, NE, FL, CHECKED, AX + 1>::run(c, pk2, Wout * c.w0[t], cst2, lin2);
            }
            real s0 = real(0), s1 = real(0), s2 = real(0);
            if constexpr (!CHECKED) {
                // the rule is mirror-symmetric (xi_t = -xi_{T-1-t}, equal weights, xi = 0 in the middle
                // of an odd rule; lhvi_model::rule_symmetric): pair the mirror nodes
#pragma unroll
                for (int t = 0; t < T / 2; ++t) {
                    const real sm = R[t] + R[T - 1 - t];
                    s0 += c.w0[t] * sm;
                    if constexpr (AX < NC) {
                        s1 += c.w1[T - 1 - t] * (R[T - 1 - t] - R[t]);
                        s2 += c.w2[t] * sm;
                    }
                }
                if constexpr ((T & 1) != 0) s0 += c.w0[T / 2] * R[T / 2];
            } else {
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    s0 += c.w0[t] * R[t];
                    if constexpr (AX < NC) {
                        s1 += c.w1[t] * R[t];
                        s2 += c.w2[t] * R[t];
                    }
                }
            }
            if constexpr (AX < NC) {
                c.m1[AX] += Wout * s1;
                c.m2[AX] += Wout * s2;
            }
            return s0;
        }
    }
};

struct SpecLaunch {
    long long chunk;      // records per block (multiple of the block size)
};

// Shared memory of one block of the record-major walk (a static allocation in the per-group kernel,
// a view of the dynamic allocation inside the persistent iteration kernel).
template <typename real, int K, int T, int NC, int FL, int HUB>
struct SpecShared {
    static constexpr int NV = 2 * K;
    static constexpr bool kHubTab = FL == kFull && HUB >= 0;
    static constexpr int TP = (T + 3) / 4 * 4;
    static constexpr int kSlotElems = NV <= 2 ? 2 : ((NV + 3) / 4) * 4;
    static constexpr int kStageArgs = FL == kFull ? NC - (HUB >= 0 ? 1 : 0) : 0;
    static constexpr int kSlotBytes = kSlotElems * (int)sizeof(real);
    static constexpr bool kStage = kStageArgs > 0 && 2 * kStageArgs * kSlotBytes * kSpecThreads <= 32768;
    alignas(16) unsigned char stage[kStage ? 2 * kStageArgs * kSlotBytes * kSpecThreads : 16];
    // axis table of the hub variable the block is working on (full records with a hub argument):
    // cross densities q_{k2}(x_{k,t}) and the per-component scalars, built once per hub and block
    alignas(16) real hq[kHubTab ? K : 1][kHubTab ? K : 1][TP];
    double scratch[(kSpecThreads / 32) * (K + 1)];
    real quad[2 * T];
    real eq[T];
    real w[K];
    real val[kCacheSlots][NV];
    real hpar[5][K];                   // mu, var, hvar, nrm, sdev
    real mom[5];                       // cm0, cm2, cm4, cm22, max |xi| (computed once per block)
    int tag[kCacheSlots];
};

// HUB: index of the hidden argument accumulated per thread across records (-1: none)
template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool WEIGHTED, int HUB>
__device__ __forceinline__ void
factor_spec_body(const GroupView<real>& g, const SpecLaunch& L, const BlockSlice bs,
                 SpecShared<real, K, T, NC, FL, HUB>& sh) {
    using F = Fast<real>;
    using C = Ctx<real, K, T, NC, NG, NE>;
    constexpr int NA = C::NA, NCT = C::NCT, NQ = C::NQ, NV = 2 * K;
    constexpr int NCS = NC > 0 ? NC : 1;
    constexpr int kSlotElems = NV <= 2 ? 2 : ((NV + 3) / 4) * 4;

    auto& s_quad = sh.quad;
    auto& s_eq = sh.eq;
    auto& s_w = sh.w;
    auto& s_tag = sh.tag;
    auto& s_val = sh.val;
    auto& s_scratch = sh.scratch;
    constexpr bool kHubTab = FL == kFull && HUB >= 0;
    auto& s_hq = sh.hq;
    auto& s_hpar = sh.hpar;
    auto& s_mom = sh.mom;
    auto& s_stage = sh.stage;

    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_eq[i] = (real)::exp(-(double)g.quad[i] * (double)g.quad[i]);
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    for (int i = threadIdx.x; i < kCacheSlots; i += blockDim.x) s_tag[i] = -1;
    for (int i = threadIdx.x; i < kCacheSlots * NV; i += blockDim.x) (&s_val[0][0])[i] = real(0);
    __syncthreads();

    C c;
    real eq[T], xi[T], wk[K];
    real eq_min = real(1);
#pragma unroll
    for (int t = 0; t < T; ++t) {
        xi[t] = s_quad[t];
        c.w0[t] = s_quad[T + t];
        c.w1[t] = c.w0[t] * xi[t];
        c.w2[t] = c.w1[t] * xi[t];
        eq[t] = s_eq[t];
        eq_min = eq[t] < eq_min ? eq[t] : eq_min;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = s_w[k];
    c.eta = g.eta;
    c.s_w = s_w;

    // Full records: log psi is a quadratic, so its quadrature sums are closed forms in the rule's
    // even moments (the rule is mirror-symmetric, odd moments vanish) -- only -log b is walked.
    //   sum W = M0^NA, sum W xi_a^2 = M0^(NA-1) M2, sum W xi_a^4 = M0^(NA-1) M4,
    //   sum W xi_a^2 xi_b^2 = M0^(NA-2) M2^2
    if constexpr (FL == kFull) {
        if (threadIdx.x == 0) {
            real M0 = real(0), M2 = real(0), M4 = real(0), xm = real(0);
            for (int t = 0; t < T; ++t) {
                const real x = s_quad[t], om = s_quad[T + t];
                M0 += om;
                M2 += om * x * x;
                M4 += om * x * x * x * x;
                xm = fabs(x) > xm ? fabs(x) : xm;
            }
            real p2 = real(1);                 // M0^(NA-2)
            for (int a = 0; a + 2 < NA; ++a) p2 *= M0;
            const real p1 = NA >= 2 ? p2 * M0 : real(1);
            s_mom[0] = p1 * M0;
            s_mom[1] = p1 * M2;
            s_mom[2] = p1 * M4;
            s_mom[3] = NA >= 2 ? p2 * M2 * M2 : real(0);
            s_mom[4] = xm;
        }
        __syncthreads();
    }
    const volatile real* v_mom = s_mom;     // re-read where used rather than recomputed per record
    int tab_key = -1;                      // hub variable whose table sits in s_hq / s_hpar

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

    // pure unary records: the quadrature nodes depend on the variable only, so they are kept
    // across records and rebuilt when the variable changes (never, inside a hub's run)
    constexpr bool kUnaryPure = FL == kPure && NC == 1 && NG == 0;
    int node_key = -1;
    real node_x[K][T], node_s[K], node_inv[K];

    const int lane = threadIdx.x & 31;
    const long long lo = (long long)bs.bid * L.chunk;
    const long long hi = lo + L.chunk < g.n ? lo + L.chunk : g.n;

    // software pipeline: record columns are fetched one tile ahead (plain register arrays --
    // nothing here may have its address taken, or the prefetch lands in local memory and the
    // store stalls on the load).  The general path prefetches what a record needs first
    // (offsets, potential id, evidence); the light unary-pure path also prefetches the weights.
    constexpr bool kPrefetchWeights = kUnaryPure && WEIGHTED;
    int c_pot = 0, n_pot = 0;
    int c_poff[NCS], n_poff[NCS], f_poff[NCS];      // current, next, and far (two tiles ahead)
    real c_egv[NG > 0 ? NG : 1], c_egs[NG > 0 ? NG : 1], n_egv[NG > 0 ? NG : 1], n_egs[NG > 0 ? NG : 1];
    real c_ec[NE > 0 ? NE : 1], n_ec[NE > 0 ? NE : 1];
    real c_wf = real(1), n_wf = real(1), c_gam = real(1), n_gam = real(1);
#define LHVI_FETCH(RR, POT, POFF, EGV, EGS, EC, WF, GAM)                                  \
    do {                                                                                   \
        if constexpr (FL != kNode) POT = __ldcs(g.pot + (RR));                             \
        _Pragma("unroll") for (int a = 0; a < NC; ++a) POFF[a] = __ldcs(g.poff + a * g.n + (RR)); \
        _Pragma("unroll") for (int j = 0; j < NG; ++j) {                                   \
            EGV[j] = __ldcs(g.egval + j * g.n + (RR));                                     \
            EGS[j] = __ldcs(g.egvar + j * g.n + (RR));                                     \
        }                                                                                  \
        _Pragma("unroll") for (int e = 0; e < NE; ++e) EC[e] = __ldcs(g.ecval + e * g.n + (RR)); \
        if constexpr (kPrefetchWeights) { WF = __ldcs(g.wf + (RR)); GAM = __ldcs(g.gam + (RR)); } \
    } while (0)
    // non-hub parameter slots are prefetched into L1 one tile ahead, which needs their offsets
    // two tiles ahead (f_poff)
    constexpr bool kSlotPrefetch = FL == kFull || (FL == kNode && NC > 0);
    // ... and from there into shared memory with cp.async (one tile ahead, double-buffered, laid
    // out [buffer][argument][16-byte chunk][thread] so that the reads are conflict-free): the gather
    // latency of the non-hub slots is then off the record's critical path without costing registers
    constexpr int kStageArgs = FL == kFull ? NC - (HUB >= 0 ? 1 : 0) : 0;
    constexpr int kSlotBytes = kSlotElems * (int)sizeof(real);
    constexpr int kChunkBytes = kSlotBytes < 16 ? kSlotBytes : 16;
    constexpr int kChunks = kSlotBytes / kChunkBytes;
    constexpr bool kStage = kStageArgs > 0 && 2 * kStageArgs * kSlotBytes * kSpecThreads <= 32768;
    static_assert(kStage == SpecShared<real, K, T, NC, FL, HUB>::kStage, "stage buffer size");
    auto stage_slots = [&](int buf, const int (&off)[NCS]) {
        if constexpr (kStage) {
            int sa = 0;
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                if (a == HUB) continue;
                if (off[a] >= 0) {
#pragma unroll
                    for (int ch = 0; ch < kChunks; ++ch) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(
                            s_stage + ((((buf * kStageArgs + sa) * kChunks + ch) * kSpecThreads + threadIdx.x) * kChunkBytes));
                        const char* src = reinterpret_cast<const char*>(g.eta + off[a]) + ch * kChunkBytes;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(kChunkBytes) : "memory");
                    }
                }
                ++sa;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
#pragma unroll
    for (int a = 0; a < NCS; ++a) { c_poff[a] = -1; n_poff[a] = -1; f_poff[a] = -1; }
    if (lo + threadIdx.x < hi) {
        const long long r0 = lo + threadIdx.x;
        LHVI_FETCH(r0, c_pot, c_poff, c_egv, c_egs, c_ec, c_wf, c_gam);
    }
    if constexpr (kSlotPrefetch) {
        if (lo + blockDim.x + threadIdx.x < hi) {
#pragma unroll
            for (int a = 0; a < NC; ++a) f_poff[a] = __ldcs(g.poff + a * g.n + lo + blockDim.x + threadIdx.x);
        }
    }
    stage_slots(0, c_poff);
    int stage_buf = 0;                     // buffer holding the current tile's slots

    // hub variable of the tile's first record (the same load in every thread: block-uniform)
    constexpr int HUBI = HUB >= 0 ? HUB : 0;
    int c_tkey = -1, n_tkey = -1;
    if constexpr (kHubTab) {
        if (lo < hi) c_tkey = __ldg(g.poff + HUBI * g.n + lo);
    }

    for (long long base = lo; base < hi; base += blockDim.x) {
        const long long r = base + threadIdx.x;
        const bool active = r < hi;
        if (r + blockDim.x < hi) {
            const long long r1 = r + blockDim.x;
            LHVI_FETCH(r1, n_pot, n_poff, n_egv, n_egs, n_ec, n_wf, n_gam);
        }
        if constexpr (kHubTab) {
            if (base + blockDim.x < hi) n_tkey = __ldg(g.poff + HUBI * g.n + base + blockDim.x);
            if (c_tkey != tab_key) {
                __syncthreads();                     // every thread is done with the old table
                const int idx = threadIdx.x;
                if (idx < K * K * T) {
                    const int k = idx / (K * T), k2 = (idx / T) % K, t = idx % T;
                    real q = s_eq[t];                // own component: exp(-xi^2)
                    if (k2 != k) {
                        const real mu_k = g.eta[c_tkey + 2 * k], var_k = g.eta[c_tkey + 2 * k + 1];
                        const real mu_2 = g.eta[c_tkey + 2 * k2], var_2 = g.eta[c_tkey + 2 * k2 + 1];
                        const real u = F::sqrt(real(2) * var_k) * s_quad[t] + (mu_k - mu_2);
                        q = F::exp_scaled(real(-0.5) * F::kExpScale * F::rcp(var_2) * (u * u));
                    }
                    s_hq[k][k2][t] = q;
                }
                if (idx < K) {
                    const real mu_k = g.eta[c_tkey + 2 * idx], var_k = g.eta[c_tkey + 2 * idx + 1];
                    const real inv = F::rcp(var_k);
                    s_hpar[0][idx] = mu_k;
                    s_hpar[1][idx] = var_k;
                    s_hpar[2][idx] = real(-0.5) * F::kExpScale * inv;
                    s_hpar[3][idx] = inv * real(1.0 / kSqrt2Pi);
                    s_hpar[4][idx] = F::sqrt(real(2) * var_k);
                }
                __syncthreads();
                tab_key = c_tkey;
            }
        }
        if constexpr (kSlotPrefetch) {
            // f_poff holds the offsets of tile +1 (loaded one tile ago): warm L1 with those slots
            if constexpr (kStage) {
                stage_slots(stage_buf ^ 1, f_poff);
            } else {
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    if (a != HUB && f_poff[a] >= 0)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(g.eta + f_poff[a]));
                }
            }
            const long long r2 = r + 2 * (long long)blockDim.x;
#pragma unroll
            for (int a = 0; a < NC; ++a) f_poff[a] = r2 < hi ? __ldcs(g.poff + a * g.n + r2) : -1;
        }
        int key[NCS];
        real gv[NCS][NV];
#pragma unroll
        for (int a = 0; a < NCS; ++a) {
            key[a] = -1;
#pragma unroll
            for (int i = 0; i < NV; ++i) gv[a][i] = real(0);
        }

        if constexpr (kUnaryPure) {
          if (active) {
            const real wf = c_wf;
            key[0] = c_poff[0];
            if (key[0] != node_key) {
                node_key = key[0];
                real slot[NV];
                load_vec<NV>(g.eta + key[0], slot);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    node_s[k] = F::sqrt(real(2) * slot[2 * k + 1]);
                    node_inv[k] = F::rcp(slot[2 * k + 1]);
#pragma unroll
                    for (int t = 0; t < T; ++t) node_x[k][t] = node_s[k] * xi[t] + slot[2 * k];
                }
            }
            // log psi(x) = c0 + l0 x + a0 x^2 after folding the point evidence (walk units)
            const real* cf = g.ptab + c_pot;
            constexpr real to_unit = real(1) / F::kUnit;
            constexpr int NCOEF = (NCT + 1) * (NCT + 2) / 2;
            real cq[NCOEF];
#pragma unroll
            for (int i = 0; i < NCOEF; ++i) cq[i] = __ldg(cf + i);
            real c0 = cq[0], l0 = cq[1], a0 = cq[1 + NCT];
            {
                // coefficient layout: c, b[NCT], then upper-triangular A row-major
                real ev[NE > 0 ? NE : 1];
#pragma unroll
                for (int e = 0; e < NE; ++e) ev[e] = c_ec[e];
                int p = 1 + NCT;
#pragma unroll
                for (int i = 0; i < NCT; ++i) {
#pragma unroll
                    for (int j = i; j < NCT; ++j) {
                        const real coef = cq[p++];
                        if (i == 0 && j > 0) l0 += coef * ev[j - 1];
                        else if (i > 0) c0 += coef * ev[i - 1] * ev[j - 1];
                    }
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) c0 += cq[2 + e] * ev[e];
            }
            c0 *= to_unit; l0 *= to_unit; a0 *= to_unit;

            real Eks[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                real e0 = real(0), e1 = real(0), e2 = real(0), qmin = real(0);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const real x = node_x[k][t];
                    const real q = c0 + x * (l0 + a0 * x);
                    qmin = F::min(qmin, q);
                    e0 += c.w0[t] * q;
                    e1 += c.w1[t] * q;
                    e2 += c.w2[t] * q;
                }
                if (qmin < F::kQFloor) {
                    e0 = e1 = e2 = real(0);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real x = node_x[k][t];
                        const real q = checked_log_psi<real>(c0 + x * (l0 + a0 * x));
                        e0 += c.w0[t] * q;
                        e1 += c.w1[t] * q;
                        e2 += c.w2[t] * q;
                    }
                }
                const real Ek = e0 * F::kUnit;
                Eks[k] = Ek;
                gv[0][2 * k] = -(node_s[k] * F::kUnit * e1) * node_inv[k];
                gv[0][2 * k + 1] = -(e2 * F::kUnit - real(0.5) * Ek) * node_inv[k];
            }
            real e_sum = real(0);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                acc[k] -= (double)(wf * Eks[k]);
                e_sum += wk[k] * Eks[k];
            }
            acc[K] -= (double)(wf * e_sum);
            if constexpr (WEIGHTED) {
#pragma unroll
                for (int i = 0; i < NV; ++i) gv[0][i] *= c_gam;
            }
          }
        } else if (active) {
            // weights are consumed at the end of the record: issue their loads now
            const real wf = (WEIGHTED || FL == kNode) ? __ldcs(g.wf + r) : real(1);
            real gam_r[NCS];
#pragma unroll
            for (int a = 0; a < NCS; ++a) gam_r[a] = real(1);
            if constexpr (FL == kNode) {
                gam_r[0] = __ldcs(g.nscale + r);
            } else if constexpr (WEIGHTED) {
#pragma unroll
                for (int a = 0; a < NC; ++a) gam_r[a] = __ldcs(g.gam + a * g.n + r);
            }
            real mu[NCS][K], var[NCS][K], hvar[NCS][K], nrm[NCS][K];
            // the record's hub variable is the one tabulated in shared memory (always, except for
            // the records of a tile in which the hub changes)
            const bool use_tab = kHubTab && c_poff[HUBI] == tab_key;
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                key[a] = c_poff[a];
                c.pt.poff[a] = key[a];
                if (kHubTab && a == HUB && use_tab) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        mu[a][k] = s_hpar[0][k];
                        var[a][k] = s_hpar[1][k];
                        hvar[a][k] = s_hpar[2][k];
                        nrm[a][k] = s_hpar[3][k];
                    }
                    continue;
                }
                real slot[kSlotElems];
                if constexpr (kStage) {
                    if (a != HUB) {
                        // this tile's group of copies was committed one tile ago (one newer group may be pending)
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                        const int sa = a - (HUB >= 0 && a > HUB ? 1 : 0);
#pragma unroll
                        for (int ch = 0; ch < kChunks; ++ch) {
                            const unsigned char* src =
                                s_stage + ((((stage_buf * kStageArgs + sa) * kChunks + ch) * kSpecThreads + threadIdx.x) * kChunkBytes);
                            constexpr int EPC = kChunkBytes / (int)sizeof(real);     // elements per chunk
                            if constexpr (kChunkBytes == 16) {
                                const float4 t = *reinterpret_cast<const float4*>(src);
                                if constexpr (sizeof(real) == 4) {
                                    slot[ch * EPC + 0] = t.x; slot[ch * EPC + 1] = t.y; slot[ch * EPC + 2] = t.z; slot[ch * EPC + 3] = t.w;
                                } else {
                                    const double2 d = *reinterpret_cast<const double2*>(&t);
                                    slot[ch * EPC + 0] = d.x; slot[ch * EPC + 1] = d.y;
                                }
                            } else {
                                const float2 t = *reinterpret_cast<const float2*>(src);
                                slot[0] = t.x; slot[1] = t.y;
                            }
                        }
                    } else {
                        real tmp[NV];
                        load_vec<NV>(g.eta + key[a], tmp);
#pragma unroll
                        for (int i = 0; i < NV; ++i) slot[i] = tmp[i];
                    }
                } else {
                    real tmp[NV];
                    load_vec<NV>(g.eta + key[a], tmp);
#pragma unroll
                    for (int i = 0; i < NV; ++i) slot[i] = tmp[i];
                }
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
                egval[j] = c_egv[j];
                const real v = c_egs[j];
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
                const real* cf = g.ptab + c_pot;
                constexpr real to_unit = real(1) / F::kUnit;         // natural log -> walk units
                cst0 = __ldg(cf) * to_unit;
                if constexpr (NCT > 0) {
#pragma unroll
                    for (int i = 0; i < NCT; ++i) lin0[i] = __ldg(cf + 1 + i) * to_unit;
                    int p = 1 + NCT;
#pragma unroll
                    for (int i = 0; i < NCT; ++i)
#pragma unroll
                        for (int j = i; j < NCT; ++j) c.A[i][j] = __ldg(cf + p++) * to_unit;
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const int j = NA + e;
                        const real xv = c_ec[e];
                        cst0 += xv * (lin0[j] + c.A[j][j] * xv);
#pragma unroll
                        for (int i = 0; i < j; ++i) lin0[i] += c.A[i][j] * xv;
#pragma unroll
                        for (int i = j + 1; i < NCT; ++i) lin0[i] += c.A[j][i] * xv;
                    }
                }
            }

            // w_k' times the normalisers of every axis: the starting value of the belief products
            real pk0[K];
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) {
                pk0[k2] = wk[k2];
                if constexpr (FL != kPure) {
#pragma unroll
                    for (int a = 0; a < NC; ++a) pk0[k2] *= nrm[a][k2];
#pragma unroll
                    for (int j = 0; j < NG; ++j) pk0[k2] *= egn[j];
                }
            }

            real Eks[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // tabulate the axes under component k
                real sdev[NCS];
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    if (kHubTab && a == HUB && use_tab) {
                        sdev[a] = s_hpar[4][k];
#pragma unroll
                        for (int k2 = 0; k2 < K; ++k2)
#pragma unroll
                            for (int t = 0; t < T; ++t) c.q[a][k2][t] = s_hq[k][k2][t];
                        continue;
                    }
                    sdev[a] = F::sqrt(real(2) * var[a][k]);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real dx = sdev[a] * xi[t];
                        if constexpr (FL == kPure) c.x[a][t] = dx + mu[a][k];
                        if constexpr (FL != kPure) {
#pragma unroll
                            // densities without their normalisers 1 / (sqrt(2pi) var): those depend
                            // on (argument, component) only and are folded into pk once per record
                            for (int k2 = 0; k2 < K; ++k2) {
                                if (k2 == k) {
                                    c.q[a][k2][t] = eq[t];                   // exp(-xi^2)
                                } else {
                                    const real u = dx + (mu[a][k] - mu[a][k2]);
                                    c.q[a][k2][t] = F::exp_scaled(hvar[a][k2] * (u * u));
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) {
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        if constexpr (FL == kPure) c.x[NC + j][t] = egs[j] * xi[t] + egval[j];
#pragma unroll
                        for (int k2 = 0; k2 < K; ++k2) c.q[NC + j][k2][t] = eq[t];
                    }
                }
                real pk[K];
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2) pk[k2] = pk0[k2];

#pragma unroll
                for (int a = 0; a < NC; ++a) { c.m1[a] = real(0); c.m2[a] = real(0); }
                c.qmin = real(0);
                // every belief on this grid is at least the own-component term, so float products
                // cannot have underflowed if that term is comfortably representable
                real own = pk0[k];
                if constexpr (FL != kPure) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) own *= eq_min;
                }
                real Ek = Walk<real, K, T, NC, NG, NE, FL, false, 0>::run(c, pk, real(1), cst0, lin0);
                bool redo = own < F::kBFloor;
                if constexpr (FL == kFull) {
                    // closed-form quadrature sums of log psi over the grid centred at cen with
                    // half-widths sc: log psi = P + sum_a Q_a xi_a + sum_a R_aa xi_a^2 + cross terms
                    real cen[NA > 0 ? NA : 1], sc[NA > 0 ? NA : 1], Dv[NA > 0 ? NA : 1];
#pragma unroll
                    for (int a = 0; a < NC; ++a) { cen[a] = mu[a][k]; sc[a] = sdev[a]; }
#pragma unroll
                    for (int j = 0; j < NG; ++j) { cen[NC + j] = egval[j]; sc[NC + j] = egs[j]; }
                    real P = cst0;
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        const real h = c.A[a][a] * cen[a];
                        Dv[a] = lin0[a] + (h + h);
                        P += cen[a] * (lin0[a] + h);
                    }
#pragma unroll
                    for (int a = 0; a < NA; ++a)
#pragma unroll
                        for (int b = a + 1; b < NA; ++b) {
                            const real ab = c.A[a][b] * cen[b];
                            Dv[a] += ab;
                            Dv[b] += c.A[a][b] * cen[a];
                            P += ab * cen[a];
                        }
                    real Raa[NA > 0 ? NA : 1], Rsum = real(0), absQ = real(0), absR = real(0);
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        Raa[a] = c.A[a][a] * sc[a] * sc[a];
                        Rsum += Raa[a];
                        absQ += fabs(sc[a] * Dv[a]);
                        absR += fabs(Raa[a]);
                    }
#pragma unroll
                    for (int a = 0; a < NA; ++a)
#pragma unroll
                        for (int b = a + 1; b < NA; ++b) absR += fabs(c.A[a][b]) * sc[a] * sc[b];
                    // lower bound of log psi on the grid: can the 1e-100 floor be active?
                    const real xi_max = v_mom[4];
                    redo = redo || (P - xi_max * (absQ + xi_max * absR)) < F::kQFloor;
                    const real cm0 = v_mom[0], cm2 = v_mom[1], cm4 = v_mom[2], cm22 = v_mom[3];
                    Ek += cm0 * P + cm2 * Rsum;
#pragma unroll
                    for (int a = 0; a < NC; ++a) {
                        c.m1[a] += cm2 * (sc[a] * Dv[a]);
                        c.m2[a] += cm2 * P + cm4 * Raa[a] + cm22 * (Rsum - Raa[a]);
                    }
                } else {
                    redo = redo || c.qmin < F::kQFloor;
                }
                if (redo) {
                    // a floor may be active on this grid: redo it with the literal formulas
#pragma unroll
                    for (int a = 0; a < NC; ++a) {
                        c.m1[a] = real(0); c.m2[a] = real(0);
#pragma unroll
                        for (int t = 0; t < T; ++t) c.x[a][t] = sdev[a] * xi[t] + mu[a][k];
                    }
#pragma unroll
                    for (int j = 0; j < NG; ++j)
#pragma unroll
                        for (int t = 0; t < T; ++t) c.x[NC + j][t] = egs[j] * xi[t] + egval[j];
                    Ek = Walk<real, K, T, NC, NG, NE, FL, true, 0>::run(c, pk, real(1), cst0, lin0);
                }
                Ek *= F::kUnit;
                Eks[k] = Ek;

#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    // g_mu = -sum W F (x-mu) / var ; g_var = -sum W F ((x-mu)^2 - var) / (2 var^2)
                    const real inv = FL != kPure ? hvar[a][k] * (real(-2) / F::kExpScale) : F::rcp(var[a][k]);
                    gv[a][2 * k] = -(sdev[a] * F::kUnit * c.m1[a]) * inv;
                    gv[a][2 * k + 1] = -(c.m2[a] * F::kUnit - real(0.5) * Ek) * inv;
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
#pragma unroll
                for (int a = 0; a < NC; ++a)
#pragma unroll
                    for (int i = 0; i < NV; ++i) gv[a][i] *= gam_r[0];
            } else if constexpr (WEIGHTED) {
#pragma unroll
                for (int a = 0; a < NC; ++a) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) gv[a][i] *= gam_r[a];
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
                        // every thread of the block passes this boundary within one tile: straight
                        // to global REDs (pipelined in L2) instead of ~1500 contended shared atomics
                        if (hub_key >= 0) red_vec<NV>(g.grad + hub_key, hub_acc);
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
        c_pot = n_pot;
        c_wf = n_wf;
        c_gam = n_gam;
        c_tkey = n_tkey;
        stage_buf ^= 1;
#pragma unroll
        for (int a = 0; a < NCS; ++a) c_poff[a] = n_poff[a];
#pragma unroll
        for (int j = 0; j < (NG > 0 ? NG : 1); ++j) { c_egv[j] = n_egv[j]; c_egs[j] = n_egs[j]; }
#pragma unroll
        for (int e = 0; e < (NE > 0 ? NE : 1); ++e) c_ec[e] = n_ec[e];
    }
#undef LHVI_FETCH

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

    publish_partials(acc, K + 1, s_scratch, g.partials, bs);

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

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool WEIGHTED, int HUB>
__global__ void __launch_bounds__(kSpecThreads, 2)
factor_spec_kernel(const GroupView<real> g, const SpecLaunch L) {
    __shared__ SpecShared<real, K, T, NC, FL, HUB> sh;
    factor_spec_body<real, K, T, NC, NG, NE, FL, WEIGHTED, HUB>(g, L, BlockSlice{(int)blockIdx.x, (int)gridDim.x, nullptr}, sh);
}

template <typename KernelT>
static int resident_blocks(KernelT kernel) {
    int per_sm = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSpecThreads, 0);
    return (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
}

// ---- pure unary records: a streaming kernel ------------------------------------------------------
//
// F = log psi only, one hidden continuous argument: per record a handful of FMAs, so the kernel
// is bound by how many bytes it keeps in flight.  Each thread owns QUAD consecutive records per
// iteration and fetches every column with one 16-byte load, one iteration ahead (double-buffered
// in registers): 5 columns x 512 B per warp in flight.  Quadrature nodes depend on the variable
// only and are cached across records; gradients of a run of records on the same variable are
// summed in registers and flushed when the variable changes -- through the block's shared cache
// for a hub (USE_CACHE), straight to global REDs otherwise.

constexpr int kQuad = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<int> { using type = int4; };

template <typename T>
__device__ __forceinline__ void load_quad(const T* __restrict__ p, T (&v)[kQuad]) {
    if constexpr (sizeof(T) == 4) {
        const int4 t = __ldcs(reinterpret_cast<const int4*>(p));
        v[0] = *reinterpret_cast<const T*>(&t.x);
        v[1] = *reinterpret_cast<const T*>(&t.y);
        v[2] = *reinterpret_cast<const T*>(&t.z);
        v[3] = *reinterpret_cast<const T*>(&t.w);
    } else {
        const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
        const double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}

template <typename real, int K, int T>
struct PureUnaryShared {
    double scratch[(kSpecThreads / 32) * (K + 1)];
    real quad[2 * T];
    real w[K];
    real val[kCacheSlots][2 * K];
    int tag[kCacheSlots];
};

template <typename real, int K, int T, int NE, bool WEIGHTED, bool USE_CACHE>
__device__ __forceinline__ void
pure_unary_body(const GroupView<real>& g, const SpecLaunch& L, const BlockSlice bs, PureUnaryShared<real, K, T>& sh) {
    using F = Fast<real>;
    constexpr int NV = 2 * K, NCT = 1 + NE, NCOEF = (NCT + 1) * (NCT + 2) / 2;
    constexpr int kSlotElems = NV <= 2 ? 2 : ((NV + 3) / 4) * 4;
    constexpr int NES = NE > 0 ? NE : 1;

    auto& s_quad = sh.quad;
    auto& s_w = sh.w;
    auto& s_tag = sh.tag;
    auto& s_val = sh.val;
    auto& s_scratch = sh.scratch;
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    for (int i = threadIdx.x; i < kCacheSlots; i += blockDim.x) s_tag[i] = -1;
    for (int i = threadIdx.x; i < kCacheSlots * NV; i += blockDim.x) (&s_val[0][0])[i] = real(0);
    __syncthreads();

    real xi[T], w0[T], w1[T], w2[T], wk[K];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        xi[t] = s_quad[t];
        w0[t] = s_quad[T + t];
        w1[t] = w0[t] * xi[t];
        w2[t] = w1[t] * xi[t];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = s_w[k];

    double acc[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;

    int run_key = -1;                    // variable of the current run of records
    real run_acc[NV];                    // its gradient so far
    real node_x[K][T], node_s[K], node_inv[K];
#pragma unroll
    for (int i = 0; i < NV; ++i) run_acc[i] = real(0);

    auto flush = [&](int key, const real (&v)[NV]) {
        if constexpr (USE_CACHE) {
            const int slot = (key / kSlotElems) & (kCacheSlots - 1);
            const int old = atomicCAS(&s_tag[slot], -1, key);
            if (old == -1 || old == key) {
#pragma unroll
                for (int i = 0; i < NV; ++i) atomicAdd(&s_val[slot][i], v[i]);
                return;
            }
        }
        red_vec<NV>(g.grad + key, v);
    };

    const long long lo = (long long)bs.bid * L.chunk;
    const long long hi = lo + L.chunk < g.n ? lo + L.chunk : g.n;
    const long long stride = (long long)blockDim.x * kQuad;

    // double-buffered column quads (all in registers)
    int c_pot[kQuad], n_pot[kQuad], c_off[kQuad], n_off[kQuad];
    real c_ec[NES][kQuad], n_ec[NES][kQuad], c_wf[kQuad], n_wf[kQuad], c_gam[kQuad], n_gam[kQuad];
#define LHVI_FETCH_QUAD(RR, POT, OFF, EC, WF, GAM)                                         \
    do {                                                                                    \
        if ((RR) + kQuad <= hi) {                                                           \
            load_quad<int>(g.pot + (RR), POT);                                              \
            load_quad<int>(g.poff + (RR), OFF);                                             \
            _Pragma("unroll") for (int e = 0; e < NE; ++e) load_quad<real>(g.ecval + e * g.n + (RR), EC[e]); \
            if constexpr (WEIGHTED) { load_quad<real>(g.wf + (RR), WF); load_quad<real>(g.gam + (RR), GAM); } \
        } else {                                                                            \
            _Pragma("unroll") for (int j = 0; j < kQuad; ++j) {                             \
                const long long q_ = (RR) + j;                                              \
                const bool ok_ = q_ < hi;                                                   \
                POT[j] = ok_ ? g.pot[q_] : 0;                                               \
                OFF[j] = ok_ ? g.poff[q_] : -1;                                             \
                _Pragma("unroll") for (int e = 0; e < NE; ++e) EC[e][j] = ok_ ? g.ecval[e * g.n + q_] : real(0); \
                if constexpr (WEIGHTED) { WF[j] = ok_ ? g.wf[q_] : real(0); GAM[j] = ok_ ? g.gam[q_] : real(0); } \
            }                                                                               \
        }                                                                                   \
    } while (0)

    long long r = lo + (long long)threadIdx.x * kQuad;
    if (r < hi) LHVI_FETCH_QUAD(r, c_pot, c_off, c_ec, c_wf, c_gam);
    // the parameter slots of the next tile are prefetched into L1 while this tile is worked on, which
    // needs the offsets two tiles ahead (f_off): without it every record waits for its own gather
    // (offset -> slot: two dependent round trips to memory per record, ~2 us, nothing to hide them
    // behind at 8 warps per block)
    int f_off[kQuad];
#pragma unroll
    for (int j = 0; j < kQuad; ++j) f_off[j] = -1;
    if (r + stride + kQuad <= hi) load_quad<int>(g.poff + r + stride, f_off);

    for (; r < hi; r += stride) {
        const long long rn = r + stride;
#pragma unroll
        for (int j = 0; j < kQuad; ++j)
            if (f_off[j] >= 0 && f_off[j] != run_key) asm volatile("prefetch.global.L1 [%0];" ::"l"(g.eta + f_off[j]));
        if (rn < hi) LHVI_FETCH_QUAD(rn, n_pot, n_off, n_ec, n_wf, n_gam);
        if (rn + stride + kQuad <= hi) {
            load_quad<int>(g.poff + rn + stride, f_off);
        } else {
#pragma unroll
            for (int j = 0; j < kQuad; ++j) f_off[j] = -1;
        }

#pragma unroll
        for (int j = 0; j < kQuad; ++j) {
            const int key = c_off[j];
            if (key < 0) continue;                       // past the end of the chunk
            if (key != run_key) {
                if (run_key >= 0) red_vec<NV>(g.grad + run_key, run_acc);     // see factor_spec_kernel
                run_key = key;
#pragma unroll
                for (int i = 0; i < NV; ++i) run_acc[i] = real(0);
                real slot[NV];
                load_vec<NV>(g.eta + key, slot);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    node_s[k] = F::sqrt(real(2) * slot[2 * k + 1]);
                    node_inv[k] = F::rcp(slot[2 * k + 1]);
#pragma unroll
                    for (int t = 0; t < T; ++t) node_x[k][t] = node_s[k] * xi[t] + slot[2 * k];
                }
            }
            // log psi(x) = c0 + l0 x + a0 x^2 after folding the point evidence (walk units);
            // coefficient layout: c, b[NCT], then upper-triangular A row-major
            const real* cf = g.ptab + c_pot[j];
            real cq[NCOEF];
#pragma unroll
            for (int i = 0; i < NCOEF; ++i) cq[i] = __ldg(cf + i);
            real c0 = cq[0], l0 = cq[1], a0 = cq[1 + NCT];
            {
                int p = 1 + NCT;
#pragma unroll
                for (int i = 0; i < NCT; ++i) {
#pragma unroll
                    for (int jj = i; jj < NCT; ++jj) {
                        const real coef = cq[p++];
                        if (i == 0 && jj > 0) l0 += coef * c_ec[jj - 1 < NES ? jj - 1 : 0][j];
                        else if (i > 0) c0 += coef * c_ec[i - 1 < NES ? i - 1 : 0][j] * c_ec[jj - 1 < NES ? jj - 1 : 0][j];
                    }
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) c0 += cq[2 + e] * c_ec[e][j];
            }
            constexpr real to_unit = real(1) / F::kUnit;
            c0 *= to_unit; l0 *= to_unit; a0 *= to_unit;

            const real wf = WEIGHTED ? c_wf[j] : real(1);
            const real gam = WEIGHTED ? c_gam[j] : real(1);
            real e_sum = real(0);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                real e0 = real(0), e1 = real(0), e2 = real(0), qmin = real(0);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const real x = node_x[k][t];
                    const real q = c0 + x * (l0 + a0 * x);
                    qmin = F::min(qmin, q);
                    e0 += w0[t] * q;
                    e1 += w1[t] * q;
                    e2 += w2[t] * q;
                }
                if (qmin < F::kQFloor) {                 // the 1e-100 floor is active: literal formula
                    e0 = e1 = e2 = real(0);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real x = node_x[k][t];
                        const real q = checked_log_psi<real>(c0 + x * (l0 + a0 * x));
                        e0 += w0[t] * q;
                        e1 += w1[t] * q;
                        e2 += w2[t] * q;
                    }
                }
                const real Ek = e0 * F::kUnit;
                // g_mu = -sum W F (x-mu) / var ; g_var = -sum W F ((x-mu)^2 - var) / (2 var^2)
                run_acc[2 * k] -= gam * (node_s[k] * F::kUnit * e1) * node_inv[k];
                run_acc[2 * k + 1] -= gam * (e2 * F::kUnit - real(0.5) * Ek) * node_inv[k];
                acc[k] -= (double)(wf * Ek);
                e_sum += wk[k] * Ek;
            }
            acc[K] -= (double)(wf * e_sum);
        }

#pragma unroll
        for (int j = 0; j < kQuad; ++j) {
            c_pot[j] = n_pot[j];
            c_off[j] = n_off[j];
            c_wf[j] = n_wf[j];
            c_gam[j] = n_gam[j];
#pragma unroll
            for (int e = 0; e < NES; ++e) c_ec[e][j] = n_ec[e][j];
        }
    }
#undef LHVI_FETCH_QUAD

    // end of the block's chunk: flush the open runs (warp-combined when the whole warp agrees)
    {
        const int lane = threadIdx.x & 31;
        const int k0 = __shfl_sync(0xffffffffu, run_key, 0);
        if (__all_sync(0xffffffffu, run_key == k0)) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                real v = run_acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                run_acc[i] = v;
            }
            if (lane == 0 && run_key >= 0) flush(run_key, run_acc);
        } else if (run_key >= 0) {
            flush(run_key, run_acc);
        }
    }

    publish_partials(acc, K + 1, s_scratch, g.partials, bs);

    if constexpr (USE_CACHE) {
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

template <typename real, int K, int T, int NE, bool WEIGHTED, bool USE_CACHE>
__global__ void __launch_bounds__(kSpecThreads, 2)
pure_unary_kernel(const GroupView<real> g, const SpecLaunch L) {
    __shared__ PureUnaryShared<real, K, T> sh;
    pure_unary_body<real, K, T, NE, WEIGHTED, USE_CACHE>(g, L, BlockSlice{(int)blockIdx.x, (int)gridDim.x, nullptr}, sh);
}

template <typename real, int K, int T, int NE>
static int launch_pure_unary(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const GroupView<real> v = make_view<real>(m, g, row0);
    const bool weighted = g->weighted != 0;
    const bool hub = ((g->hub_mask >> g->nd) & 1) != 0;
    // 16-byte column loads need every column start aligned: single-column arrays always are
    // (allocation granularity), a second evidence column only if n is a multiple of 4
    if (NE > 1 && (g->n % kQuad) != 0) return 1;
    auto go = [&](auto kernel) {
        static int resident = 0;
        if (resident == 0) resident = resident_blocks(kernel);
        const long long tile = (long long)kSpecThreads * kQuad;
        long long blocks = (v.n + tile - 1) / tile;
        if (blocks > resident) blocks = resident;
        if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
        SpecLaunch L;
        L.chunk = ((v.n + blocks - 1) / blocks + tile - 1) / tile * tile;
        blocks = (v.n + L.chunk - 1) / L.chunk;
        kernel<<<(unsigned)blocks, kSpecThreads, 0, s>>>(v, L);
        return check_launch("pure_unary_kernel");
    };
    if (weighted) return hub ? go(pure_unary_kernel<real, K, T, NE, true, true>)
                             : go(pure_unary_kernel<real, K, T, NE, true, false>);
    return hub ? go(pure_unary_kernel<real, K, T, NE, false, true>)
               : go(pure_unary_kernel<real, K, T, NE, false, false>);
}

// ---- folded unary records: a streaming reduction ---------------------------------------------------
//
// Input: the fold columns of lhvi_group (log psi(x) = c0 + l0 x + a0 x^2 per record, evidence already
// folded in).  E_k[log psi], E_k[log psi (x-mu)] and E_k[log psi ((x-mu)^2 - var)] are *linear* in
// (c0, l0, a0) with coefficients that depend on the variable only, so inside a run of records on
// the same variable a thread just accumulates the gamma- and W_f-weighted sums of the coefficients
// (re-expanded around the belief mean so that the sums stay small: 9 FMAs per record) and applies
// the per-component formulas once per run, in double.  Every record is still read and folded in
// every iteration -- nothing is cached across iterations.
//
// The reference integrates log(psi + 1e-100); a record whose quadratic could reach the floor on
// the run's node interval (lower bound A - |B| R - |C| R^2 below the threshold) is excluded from
// the sums and evaluated node by node with the literal formula.
//
// The quadrature rule enters through its even moments M0, M2, M4 (the rule is symmetric, the host
// checks it), so T is a run-time value here.

// Resident blocks per SM of the streaming kernel.  Two (16 warps, up to 128 registers: no spills, the compiler
// keeps a tile's six 16-byte loads per thread in flight) measured better than three and four on a B200 --
// 42.4 / 47.2 / 47.5 us launched alone at 7.0 M records, 121.3 / 123.3 / 125.7 us per iteration (one block: 56.2 /
// 134.6) -- in line with
// the six-column read micro-benchmark (profiles/r2_stream6_microbench.txt: 8 blocks per SM 42 us, 4 blocks 35 us).
#ifndef LHVI_FOLD_BLOCKS
#define LHVI_FOLD_BLOCKS 2
#endif
#ifndef LHVI_FOLD_WAVES
#define LHVI_FOLD_WAVES 1
#endif
constexpr int kFoldThreads = 256;
constexpr int kFoldTile = kFoldThreads * kQuad;
static_assert(kFoldTile == 1024, "lhvi.h documents n_pad as a multiple of 1024");

// literal evaluation of one record (floor possibly active); adds into acc / rg
template <typename real, int K>
__device__ __noinline__ void fold_slow_record(const real* slot, const real* s_quad, int T,
                                              const real* s_w, double c0, double l0, double a0,
                                              double wf, double gam, double* acc, double* rg) {
    double esum = 0.0;
    for (int k = 0; k < K; ++k) {
        const double mu = (double)slot[2 * k], var = (double)slot[2 * k + 1];
        const double sd = ::sqrt(2.0 * var);
        double e0 = 0.0, e1 = 0.0, e2 = 0.0;
        for (int t = 0; t < T; ++t) {
            const double xi = (double)s_quad[t], om = (double)s_quad[T + t];
            const double x = sd * xi + mu;
            const double lq = ::log(::exp(c0 + x * (l0 + a0 * x)) + kEps);
            e0 += om * lq;
            e1 += om * xi * lq;
            e2 += om * xi * xi * lq;
        }
        acc[k] -= wf * e0;
        esum += (double)s_w[k] * e0;
        rg[2 * k] -= gam * (sd * e1) / var;
        rg[2 * k + 1] -= gam * (e2 - 0.5 * e0) / var;
    }
    acc[K] -= wf * esum;
}

template <typename real, int K>
struct FoldShared {
    real quad[2 * LHVI_MAX_T];
    double mom[4];                     // M0, M2, M4, max |xi|
    real w[K];
    int tag[kCacheSlots];
    real val[kCacheSlots][2 * K];
};

template <typename real> struct RunStart { real xbar, R; };

// expansion point (belief mean) and node radius of a variable's run
template <typename real, int K>
__device__ __noinline__ RunStart<real> fold_begin_run(const real* eta, int key,
                                                      const FoldShared<real, K>* sh) {
    constexpr int NV = 2 * K;
    real slot[NV];
    load_vec<NV>(eta + key, slot);
    real m = real(0);
#pragma unroll
    for (int k = 0; k < K; ++k) m += sh->w[k] * slot[2 * k];
    const real xi_max = (real)sh->mom[3];
    real rad = real(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const real d = fabs(slot[2 * k] - m) + Fast<real>::sqrt(real(2) * slot[2 * k + 1]) * xi_max;
        rad = d > rad ? d : rad;
    }
    RunStart<real> out;
    out.xbar = m;
    out.R = rad * real(1.0001);
    return out;
}

// close a run: per-component formulas on the accumulated sums (in double), gradient out through
// the block's shared cache (hubs) or vector REDs
template <typename real, int K, bool USE_CACHE>
__device__ __forceinline__ void fold_close_run(const real* eta, real* grad, int key, real xbar,
                                            real sg0, real sg1, real sg2, real sw0, real sw1, real sw2,
                                            FoldShared<real, K>* sh, double* acc, double* rg) {
    constexpr int NV = 2 * K;
    constexpr int kSlotElems = NV <= 2 ? 2 : ((NV + 3) / 4) * 4;
    real slot[NV];
    load_vec<NV>(eta + key, slot);
    const double M0 = sh->mom[0], M2 = sh->mom[1], M4 = sh->mom[2];
    const double G0 = (double)sg0, G1 = (double)sg1, G2 = (double)sg2;
    const double W0 = (double)sw0, W1 = (double)sw1, W2 = (double)sw2;
    real v[NV];
    double esum = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double mu = (double)slot[2 * k], var = (double)slot[2 * k + 1];
        const double d = mu - (double)xbar, s2 = 2.0 * var;
        // q(mu + s xi) = A + B xi + C xi^2 (s = sqrt(2 var)); the W_f-weighted sum needs E[q] only
        const double Aw = W0 + d * (W1 + W2 * d), Cw = W2 * s2;
        const double Ek = M0 * Aw + M2 * Cw;
        acc[k] -= Ek;
        esum += (double)sh->w[k] * Ek;
        const double Ag = G0 + d * (G1 + G2 * d), Bg = G1 + 2.0 * G2 * d, Cg = G2 * s2;
        const double e0 = M0 * Ag + M2 * Cg;
        const double e2 = M2 * Ag + M4 * Cg;
        // sum w xi q = s M2 Bg, so g_mu = -(s * s M2 Bg) / var = -2 M2 Bg ; g_var = -(e2 - e0 / 2) / var
        double gm = -2.0 * M2 * Bg, gv = -(e2 - 0.5 * e0) / var;
        if (rg != nullptr) {           // records of this run that took the literal path
            gm += rg[2 * k];
            gv += rg[2 * k + 1];
            rg[2 * k] = 0.0;
            rg[2 * k + 1] = 0.0;
        }
        v[2 * k] = (real)gm;
        v[2 * k + 1] = (real)gv;
    }
    acc[K] -= esum;
    if constexpr (USE_CACHE) {
        const int slot_i = (key / kSlotElems) & (kCacheSlots - 1);
        const int old = atomicCAS(&sh->tag[slot_i], -1, key);
        if (old == -1 || old == key) {
#pragma unroll
            for (int i = 0; i < NV; ++i) atomicAdd(&sh->val[slot_i][i], v[i]);
            return;
        }
    }
    red_vec<NV>(grad + key, v);
}

// per-thread state of the current run: expansion point, node radius, weighted coefficient sums
template <typename real>
struct FoldState {
    int run_key;
    real xbar, R;
    real sg0, sg1, sg2;      // gamma-weighted sums of (A', B', C')
    real sw0, sw1, sw2;      // W_f-weighted sums
};

// what the uncommon path needs besides the state: the tile's records (already loaded by the
// caller, handed over through memory) and the model pointers
template <typename real>
struct FoldCols {
    real c0[kQuad], l0[kQuad], a0[kQuad], wf[kQuad], gam[kQuad];
    int off[kQuad];
    const real* eta;
    real* grad;
    int T;
};

// natural-log units: below these the floor is visible at the working precision
template <typename real> __device__ __forceinline__ constexpr real fold_floor() {
    // log(e^q + 1e-100) - q = log1p(e^(-230.26 - q)) reaches half an ulp of q at q = -219 (float)
    // and q = -198 (double)
    return sizeof(real) == 4 ? real(-216.0) : real(-195.0);
}

// accumulators of the uncommon path; they live in local memory and are touched only there
template <int K>
struct FoldSlowAcc {
    double acc[K + 1];
    double rg[2 * K];
};

// The uncommon tile: a run ends inside it, or a record may touch the floor.  Processes the four
// records of this thread one by one from global memory; kept out of line (and its state passed
// through memory) so that the streaming loop keeps everything in registers.
template <typename real, int K, bool WEIGHTED, bool USE_CACHE>
__device__ __noinline__ void fold_slow_tile(const FoldCols<real>* c, FoldState<real>* st,
                                            FoldShared<real, K>* sh, FoldSlowAcc<K>* sa, bool first) {
    if (first) {
        for (int i = 0; i <= K; ++i) sa->acc[i] = 0.0;
        for (int i = 0; i < 2 * K; ++i) sa->rg[i] = 0.0;
    }
#pragma unroll 1
    for (int j = 0; j < kQuad; ++j) {
        const int key = c->off[j];
        if (key != st->run_key) {
            // a run ending mid-chunk ends for every thread of the block within two tiles: these
            // closes go straight to global REDs (pipelined in L2) -- through the shared cache they
            // would be ~1500 same-address shared atomics in a row and stall the block for ~50 us
            if (st->run_key >= 0)
                fold_close_run<real, K, false>(c->eta, c->grad, st->run_key, st->xbar, st->sg0, st->sg1, st->sg2,
                                               st->sw0, st->sw1, st->sw2, sh, sa->acc, sa->rg);
            st->sg0 = st->sg1 = st->sg2 = st->sw0 = st->sw1 = st->sw2 = real(0);
            const RunStart<real> b = fold_begin_run<real, K>(c->eta, key, sh);
            st->run_key = key;
            st->xbar = b.xbar;
            st->R = b.R;
        }
        const real c0 = c->c0[j], l0 = c->l0[j], a0 = c->a0[j];
        const real wf = WEIGHTED ? c->wf[j] : real(1);
        const real gam = WEIGHTED ? c->gam[j] : real(1);
        const real Bp = l0 + real(2) * a0 * st->xbar;
        const real Ap = c0 + st->xbar * (l0 + a0 * st->xbar);
        const real low = Ap - (fabs(Bp) + fabs(a0) * st->R) * st->R;
        if (low < fold_floor<real>()) {
            fold_slow_record<real, K>(c->eta + key, sh->quad, c->T, sh->w, (double)c0, (double)l0, (double)a0,
                                      (double)wf, (double)gam, sa->acc, sa->rg);
        } else {
            st->sg0 += gam * Ap; st->sg1 += gam * Bp; st->sg2 += gam * a0;
            st->sw0 += wf * Ap;  st->sw1 += wf * Bp;  st->sw2 += wf * a0;
        }
    }
}

// Close the open runs of a whole warp.  The formulas are linear in the sums, so a warp whose lanes
// all sit in the same run (and whose expansion point is therefore the same) adds the sums up first
// and lets one lane apply them.  Out of line, state through memory: the streaming loop keeps its
// registers.  G_w / energy go to sa->acc; the sums and sa->rg are cleared.
template <typename real, int K, bool USE_CACHE>
__device__ __noinline__ void fold_close_warp(const real* eta, real* grad, FoldState<real>* st,
                                             FoldShared<real, K>* sh, FoldSlowAcc<K>* sa, bool slow_used) {
    constexpr int NV = 2 * K;
    const int lane = threadIdx.x & 31;
    const int run_key = st->run_key;
    real sg0 = st->sg0, sg1 = st->sg1, sg2 = st->sg2, sw0 = st->sw0, sw1 = st->sw1, sw2 = st->sw2;
    const int k0 = __shfl_sync(0xffffffffu, run_key, 0);
    const bool uniform = __all_sync(0xffffffffu, run_key == k0);
    if (uniform) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sg0 += __shfl_xor_sync(0xffffffffu, sg0, o);
            sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
            sg2 += __shfl_xor_sync(0xffffffffu, sg2, o);
            sw0 += __shfl_xor_sync(0xffffffffu, sw0, o);
            sw1 += __shfl_xor_sync(0xffffffffu, sw1, o);
            sw2 += __shfl_xor_sync(0xffffffffu, sw2, o);
        }
        double rgs[NV];
        const bool any_slow = __any_sync(0xffffffffu, slow_used);
        if (any_slow) {               // literal-path gradients of this run, summed over the warp
#pragma unroll
            for (int i = 0; i < NV; ++i) rgs[i] = warp_sum(slow_used ? sa->rg[i] : 0.0);
        }
        if (lane == 0 && run_key >= 0)
            fold_close_run<real, K, USE_CACHE>(eta, grad, run_key, st->xbar, sg0, sg1, sg2, sw0, sw1, sw2, sh, sa->acc,
                                               any_slow ? rgs : nullptr);
        if (slow_used) {
#pragma unroll
            for (int i = 0; i < NV; ++i) sa->rg[i] = 0.0;
        }
    } else if (run_key >= 0) {
        fold_close_run<real, K, USE_CACHE>(eta, grad, run_key, st->xbar, sg0, sg1, sg2, sw0, sw1, sw2, sh, sa->acc,
                                           slow_used ? sa->rg : nullptr);
    }
    st->sg0 = st->sg1 = st->sg2 = st->sw0 = st->sw1 = st->sw2 = real(0);
}

// ---- TMA ring of the streaming kernel -----------------------------------------------------------------
// The record columns of a tile (1024 records: slot offsets, c0, l0, a0 and -- lifted models -- W_f,
// gamma) are brought into shared memory by the TMA unit (cp.async.bulk, one 4 KB / 8 KB copy per
// column, issued by one thread, completion counted in bytes on an mbarrier) kFoldStages tiles ahead
// of the arithmetic.  What keeps HBM busy is then the depth of the ring (2 x 24 KB per block in
// flight or landed), not the number of resident warps: the LDG version kept 6 x 512 B per warp in
// flight and needed 32 warps per SM for 3.8 TB/s, which the persistent iteration kernel (16 warps per
// SM next to the 128-register run-major records) cannot give it.  Two stages, not more: with four
// (96 KB per block, 192 KB per SM) the persistent kernel's other record groups lost the L1 cache
// they gather through and ran 1.5x slower (profiles/r2_iter_plan.md).
//
// Used by the persistent iteration kernel only (TMA = true).  Alone on the GPU the streaming kernel is
// NOT faster with it -- 44.1 us with four stages at 2 blocks per SM, 56.3 us with two stages at 4
// blocks per SM, against 44.4 us for the LDG loop at 4 blocks per SM (7.0 M records, B200) -- because
// bytes in flight are not what limits it there: the per-tile dependency chain is (IPC 0.8 per SM).
// Inside the persistent kernel, at 16 warps per SM, the ring makes the streamed group 20 % faster
// (97 -> 79 block-microseconds per block at 67 blocks).
template <typename real> struct FoldRingShape {
    static constexpr int kStages = 2;
    static constexpr int kColBytes = kFoldTile * (int)sizeof(real);
};

template <typename real, int K, bool TMA = false>
struct FoldBlockShared {
    // [stage][column][bytes of one column tile]; columns: poff, c0, l0, a0, wf, gam
    alignas(128) unsigned char ring[TMA ? FoldRingShape<real>::kStages : 1][TMA ? 6 : 1]
                                   [TMA ? FoldRingShape<real>::kColBytes : 16];
    alignas(8) unsigned long long full[FoldRingShape<real>::kStages];     // mbarriers: bytes of a stage have landed
    FoldShared<real, K> sh;
    double scratch[(kFoldThreads / 32) * (K + 1)];
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LHVI_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LHVI_MBAR_DONE;\n"
        "bra LHVI_MBAR_WAIT;\n"
        "LHVI_MBAR_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// one column tile, global -> shared, completion on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <typename real, int K, bool WEIGHTED, bool USE_CACHE, bool TMA>
__device__ __forceinline__ void
unary_fold_body(const GroupView<real>& g, const SpecLaunch& L, const BlockSlice bs, FoldBlockShared<real, K, TMA>& shb) {
    constexpr int NV = 2 * K;

    FoldShared<real, K>& sh = shb.sh;
    auto& s_scratch = shb.scratch;
    const int T = g.T;
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) sh.quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) sh.w[i] = g.w[i];
    for (int i = threadIdx.x; i < kCacheSlots; i += blockDim.x) sh.tag[i] = -1;
    for (int i = threadIdx.x; i < kCacheSlots * NV; i += blockDim.x) (&sh.val[0][0])[i] = real(0);
    if (threadIdx.x == 0) {
        double m0 = 0.0, m2 = 0.0, m4 = 0.0, xm = 0.0;
        for (int t = 0; t < T; ++t) {
            const double x = (double)g.quad[t], om = (double)g.quad[T + t];
            m0 += om; m2 += om * x * x; m4 += om * x * x * x * x;
            xm = ::fmax(xm, ::fabs(x));
        }
        sh.mom[0] = m0; sh.mom[1] = m2; sh.mom[2] = m4; sh.mom[3] = xm;
    }
    __syncthreads();

    // 32-bit record indices (n_pad < 2^31 is checked at launch).  L.chunk holds the number of
    // tiles: block b takes tiles [tiles b / B, tiles (b + 1) / B), so that block sizes differ by
    // at most one tile and every SM (LHVI_FOLD_BLOCKS resident blocks) streams the same number of bytes.
    const unsigned lo = (unsigned)(L.chunk * bs.bid / bs.nblocks) * kFoldTile;
    const unsigned hi = (unsigned)(L.chunk * (bs.bid + 1) / bs.nblocks) * kFoldTile;
    const real* __restrict__ col0 = g.fold;
    const real* __restrict__ col1 = g.fold + g.n_pad;
    const real* __restrict__ col2 = g.fold + 2 * g.n_pad;

    if constexpr (TMA) {
    // first tiles in flight before anything else happens in this block
    if (threadIdx.x == 0) {
        for (int st = 0; st < FoldRingShape<real>::kStages; ++st) mbar_init(&shb.full[st], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int tiles0 = (int)((hi - lo) / kFoldTile);
        constexpr unsigned cb = (unsigned)FoldRingShape<real>::kColBytes;
        for (int t = 0; t < tiles0 && t < FoldRingShape<real>::kStages; ++t) {
            const unsigned r0 = lo + (unsigned)t * kFoldTile;
            mbar_expect_tx(&shb.full[t], kFoldTile * 4u + (WEIGHTED ? 5u : 3u) * cb);
            tma_load_1d(shb.ring[t][0], g.poff + r0, kFoldTile * 4u, &shb.full[t]);
            tma_load_1d(shb.ring[t][1], col0 + r0, cb, &shb.full[t]);
            tma_load_1d(shb.ring[t][2], col1 + r0, cb, &shb.full[t]);
            tma_load_1d(shb.ring[t][3], col2 + r0, cb, &shb.full[t]);
            if constexpr (WEIGHTED) {
                tma_load_1d(shb.ring[t][4], g.wf + r0, cb, &shb.full[t]);
                tma_load_1d(shb.ring[t][5], g.gam + r0, cb, &shb.full[t]);
            }
        }
    }
    __syncthreads();                     // the barriers are initialised before anybody waits on them
    }
    FoldSlowAcc<K> sa;                   // written by the uncommon paths only (local memory)
    for (int i = 0; i <= K; ++i) sa.acc[i] = 0.0;
    for (int i = 0; i < 2 * K; ++i) sa.rg[i] = 0.0;
    bool slow_used = false;              // sa.rg may be non-zero
    bool sa_used = false;                // sa.acc may be non-zero

    // open the run of this thread's first record, so that the first tile is an ordinary one
    unsigned r = lo + threadIdx.x * kQuad;
    int run_key = -1;
    real xbar = real(0), R = real(0);
    if (r < hi) {
        run_key = g.poff[r];
        const RunStart<real> b = fold_begin_run<real, K>(g.eta, run_key, &sh);
        xbar = b.xbar;
        R = b.R;
    }
    real sg0 = real(0), sg1 = real(0), sg2 = real(0);
    real sw0 = real(0), sw1 = real(0), sw2 = real(0);

    // producer side of the ring (TMA): one thread arms a stage's mbarrier with the bytes to expect and
    // issues the column copies of one tile
    constexpr int kStages = FoldRingShape<real>::kStages;
    constexpr unsigned kColBytes = (unsigned)FoldRingShape<real>::kColBytes;
    constexpr unsigned kOffBytes = kFoldTile * 4u;
    constexpr unsigned kStageBytes = kOffBytes + (WEIGHTED ? 5u : 3u) * kColBytes;
    const int n_tiles = (int)((hi - lo) / kFoldTile);
    auto issue = [&](int tile) {
        if constexpr (TMA) {
            const int st = tile % kStages;
            const unsigned r0 = lo + (unsigned)tile * kFoldTile;
            mbar_expect_tx(&shb.full[st], kStageBytes);
            tma_load_1d(shb.ring[st][0], g.poff + r0, kOffBytes, &shb.full[st]);
            tma_load_1d(shb.ring[st][1], col0 + r0, kColBytes, &shb.full[st]);
            tma_load_1d(shb.ring[st][2], col1 + r0, kColBytes, &shb.full[st]);
            tma_load_1d(shb.ring[st][3], col2 + r0, kColBytes, &shb.full[st]);
            if constexpr (WEIGHTED) {
                tma_load_1d(shb.ring[st][4], g.wf + r0, kColBytes, &shb.full[st]);
                tma_load_1d(shb.ring[st][5], g.gam + r0, kColBytes, &shb.full[st]);
            }
        }
    };
    // (TMA: the barriers were initialised and the first kStages tiles issued at the top of the kernel.
    // LDG: the loads are kept in flight by occupancy -- LHVI_FOLD_BLOCKS blocks x 8 warps per SM, six 512-byte requests
    // per warp and tile -- rather than by a second set of column registers.)
    // constant records riding with this group (lhvi_group::cst_*): this block's slice, one record per thread
    // and tile (the loads are independent of everything else and hide behind the tile's)
    // (32-bit indices and a float partial sum -- a handful of records per thread -- keep the streaming loop
    // within its 64 registers; cst_n < 2^31 is checked at launch)
    unsigned cst_i = 0, cst_hi = 0;
    real cst_sum = real(0);
    if (g.cst_n > 0) {
        const long long per = (g.cst_n + bs.nblocks - 1) / bs.nblocks;
        const long long c_hi = per * (bs.bid + 1) < g.cst_n ? per * (bs.bid + 1) : g.cst_n;
        cst_i = (unsigned)(per * bs.bid + threadIdx.x);
        cst_hi = (unsigned)(c_hi > 0 ? c_hi : 0);
    }
    int tile_i = 0;
#pragma unroll 1
    for (; r < hi; r += kFoldTile, ++tile_i) {
        real q_c0[kQuad], q_l0[kQuad], q_a0[kQuad], q_wf[kQuad], q_gam[kQuad];
        int q_off[kQuad];
        if (cst_i < cst_hi) {
            const real wf = g.cst_wf != nullptr ? __ldg(g.cst_wf + cst_i) : real(1);
            cst_sum += wf * Math<real>::log_psi(__ldg(g.cst_q + cst_i));
            cst_i += blockDim.x;
        }
        if constexpr (TMA) {
            const int st = tile_i % kStages;
            mbar_wait(&shb.full[st], (unsigned)((tile_i / kStages) & 1));
            const unsigned char* base = shb.ring[st][0];
            auto quad_of = [&](int col, auto (&dst)[kQuad]) {
                using E = typename std::remove_reference<decltype(dst[0])>::type;
                const unsigned char* p = base + (size_t)col * kColBytes + (size_t)threadIdx.x * kQuad * sizeof(E);
                if constexpr (sizeof(E) == 4) {
                    const int4 t = *reinterpret_cast<const int4*>(p);
                    dst[0] = *reinterpret_cast<const E*>(&t.x); dst[1] = *reinterpret_cast<const E*>(&t.y);
                    dst[2] = *reinterpret_cast<const E*>(&t.z); dst[3] = *reinterpret_cast<const E*>(&t.w);
                } else {
                    const double2 a = *reinterpret_cast<const double2*>(p);
                    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
                    dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y;
                }
            };
            quad_of(0, q_off);
            quad_of(1, q_c0);
            quad_of(2, q_l0);
            quad_of(3, q_a0);
            if constexpr (WEIGHTED) { quad_of(4, q_wf); quad_of(5, q_gam); }
            // every thread has its quad in registers: the stage can take the tile kStages ahead
            __syncthreads();
            if (threadIdx.x == 0 && tile_i + kStages < n_tiles) issue(tile_i + kStages);
        } else {
            load_quad<int>(g.poff + r, q_off);
            load_quad<real>(col0 + r, q_c0);
            load_quad<real>(col1 + r, q_l0);
            load_quad<real>(col2 + r, q_a0);
            if constexpr (WEIGHTED) { load_quad<real>(g.wf + r, q_wf); load_quad<real>(g.gam + r, q_gam); }
        }

        // a whole warp leaving its runs at this quad (the block's range crosses into the next hub's
        // records) closes them together: one close and one flush per warp instead of 32 closes and
        // 64 same-address REDs (the per-thread path made the boundary blocks ~12 us late)
        if (__any_sync(0xffffffffu, q_off[0] != run_key)) {
            if (__all_sync(0xffffffffu, q_off[0] != run_key)) {
                FoldState<real> st;
                st.run_key = run_key; st.xbar = xbar; st.R = R;
                st.sg0 = sg0; st.sg1 = sg1; st.sg2 = sg2; st.sw0 = sw0; st.sw1 = sw1; st.sw2 = sw2;
                fold_close_warp<real, K, USE_CACHE>(g.eta, g.grad, &st, &sh, &sa, slow_used);
                sa_used = true;
                sg0 = sg1 = sg2 = sw0 = sw1 = sw2 = real(0);
                run_key = q_off[0];
                const RunStart<real> b = fold_begin_run<real, K>(g.eta, run_key, &sh);
                xbar = b.xbar;
                R = b.R;
            }
        }

        // accumulate speculatively (which also keeps every load ahead of the branch), commit if
        // the whole quad belongs to the open run and stays clear of the floor
        bool fast = true;
        real t0 = sg0, t1 = sg1, t2 = sg2, u0 = sw0, u1 = sw1, u2 = sw2;
#pragma unroll
        for (int j = 0; j < kQuad; ++j) {
            const real a0 = q_a0[j];
            const real Bp = q_l0[j] + (a0 + a0) * xbar;
            const real Ap = q_c0[j] + xbar * (q_l0[j] + a0 * xbar);
            const real low = Ap - (fabs(Bp) + fabs(a0) * R) * R;
            fast = fast && q_off[j] == run_key && !(low < fold_floor<real>());
            const real wf = WEIGHTED ? q_wf[j] : real(1);
            const real gam = WEIGHTED ? q_gam[j] : real(1);
            t0 += gam * Ap; t1 += gam * Bp; t2 += gam * a0;
            u0 += wf * Ap;  u1 += wf * Bp;  u2 += wf * a0;
        }
        if (fast) {
            sg0 = t0; sg1 = t1; sg2 = t2; sw0 = u0; sw1 = u1; sw2 = u2;
        } else {
            FoldCols<real> cols;
#pragma unroll
            for (int j = 0; j < kQuad; ++j) {
                cols.c0[j] = q_c0[j]; cols.l0[j] = q_l0[j]; cols.a0[j] = q_a0[j]; cols.off[j] = q_off[j];
                cols.wf[j] = WEIGHTED ? q_wf[j] : real(1);
                cols.gam[j] = WEIGHTED ? q_gam[j] : real(1);
            }
            cols.eta = g.eta; cols.grad = g.grad; cols.T = T;
            FoldState<real> st;
            st.run_key = run_key; st.xbar = xbar; st.R = R;
            st.sg0 = sg0; st.sg1 = sg1; st.sg2 = sg2; st.sw0 = sw0; st.sw1 = sw1; st.sw2 = sw2;
            fold_slow_tile<real, K, WEIGHTED, USE_CACHE>(&cols, &st, &sh, &sa, false);
            slow_used = true;
            sa_used = true;
            run_key = st.run_key; xbar = st.xbar; R = st.R;
            sg0 = st.sg0; sg1 = st.sg1; sg2 = st.sg2; sw0 = st.sw0; sw1 = st.sw1; sw2 = st.sw2;
        }
    }

    // ---- close the open runs.  The formulas are linear in the sums, so a warp whose lanes all sit
    // in the same run (and whose expansion point is therefore the same) adds the sums up first
    // and lets one lane apply them.
    double acc[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;
    {
        const int lane = threadIdx.x & 31;
        const int k0 = __shfl_sync(0xffffffffu, run_key, 0);
        const bool uniform = __all_sync(0xffffffffu, run_key == k0);
        if (uniform) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sg0 += __shfl_xor_sync(0xffffffffu, sg0, o);
                sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
                sg2 += __shfl_xor_sync(0xffffffffu, sg2, o);
                sw0 += __shfl_xor_sync(0xffffffffu, sw0, o);
                sw1 += __shfl_xor_sync(0xffffffffu, sw1, o);
                sw2 += __shfl_xor_sync(0xffffffffu, sw2, o);
            }
            double rgs[NV];
            const bool any_slow = __any_sync(0xffffffffu, slow_used);
            if (any_slow) {               // literal-path gradients of this run, summed over the warp
#pragma unroll
                for (int i = 0; i < NV; ++i) rgs[i] = warp_sum(slow_used ? sa.rg[i] : 0.0);
            }
            if (lane == 0 && run_key >= 0)
                fold_close_run<real, K, USE_CACHE>(g.eta, g.grad, run_key, xbar, sg0, sg1, sg2, sw0, sw1, sw2, &sh, acc,
                                                   any_slow ? rgs : nullptr);
        } else if (run_key >= 0) {
            fold_close_run<real, K, USE_CACHE>(g.eta, g.grad, run_key, xbar, sg0, sg1, sg2, sw0, sw1, sw2, &sh, acc,
                                               slow_used ? sa.rg : nullptr);
        }
        if (sa_used) {
#pragma unroll
            for (int i = 0; i <= K; ++i) acc[i] += sa.acc[i];
        }
    }

    // ---- constant records fused into this group (lhvi_group::cst_*): what is left of this block's slice
    // after the tiles (a block with few tiles), then the sums: E_k[F] = log(psi + 1e-100) for every k
    if (g.cst_n > 0) {
        for (; cst_i < cst_hi; cst_i += blockDim.x) {
            const real wf = g.cst_wf != nullptr ? __ldg(g.cst_wf + cst_i) : real(1);
            cst_sum += wf * Math<real>::log_psi(__ldg(g.cst_q + cst_i));
        }
        double wsum = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { acc[k] -= (double)cst_sum; wsum += (double)sh.w[k]; }
        acc[K] -= (double)cst_sum * wsum;
    }

    publish_partials(acc, K + 1, s_scratch, g.partials, bs);
    // (publish_partials ends with a barrier: nobody waits on the ring any more; the shared memory may
    // be another record group's next)
    if constexpr (TMA) {
        if (threadIdx.x == 0)
            for (int st = 0; st < FoldRingShape<real>::kStages; ++st) mbar_inval(&shb.full[st]);
    }

    if constexpr (USE_CACHE) {
        for (int slot = threadIdx.x; slot < kCacheSlots; slot += blockDim.x) {
            const int tag = sh.tag[slot];
            if (tag >= 0) {
                real v[NV];
#pragma unroll
                for (int j = 0; j < NV; ++j) v[j] = sh.val[slot][j];
                red_vec<NV>(g.grad + tag, v);
            }
        }
    }
}

template <typename real, int K, bool WEIGHTED, bool USE_CACHE>
__global__ void __launch_bounds__(kFoldThreads, LHVI_FOLD_BLOCKS)
unary_fold_kernel(const GroupView<real> g, const SpecLaunch L) {
    __shared__ FoldBlockShared<real, K, false> shb;
    unary_fold_body<real, K, WEIGHTED, USE_CACHE, false>(g, L, BlockSlice{(int)blockIdx.x, (int)gridDim.x, nullptr}, shb);
}

template <typename real, int K>
static int launch_unary_fold(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const GroupView<real> v = make_view<real>(m, g, row0);
    const bool weighted = g->weighted != 0;
    const bool hub = ((g->hub_mask >> g->nd) & 1) != 0;
    auto go = [&](auto kernel) {
        static int resident = 0;
        if (resident == 0) {
            int per_sm = 0, sms = 0, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kFoldThreads, 0);
            resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
        }
        if (v.n_pad >= (1ll << 31)) { set_error("unary_fold_kernel: more than 2^31 records in one group"); return (int)LHVI_ELIMIT; }
        const long long tiles = v.n_pad / kFoldTile;
        // LHVI_FOLD_WAVES > 1 oversubscribes the SMs so that the hardware block scheduler evens out
        // slow SMs; measured neutral (44-46 us either way on the bench workload), so one wave
        long long blocks = tiles < LHVI_FOLD_WAVES * resident ? tiles : LHVI_FOLD_WAVES * resident;
        // a small group (a rank's shard of a strong-scaled model): whole chunks of ceil(tiles /
        // resident) tiles per block -- the host pads hub runs to that chunk (engine.align_runs), so
        // that no block crosses a hub boundary at all
        if (tiles <= 4 * (long long)resident) {
            const long long c = (tiles + resident - 1) / resident;
            blocks = (tiles + c - 1) / c;
        }
        if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
        SpecLaunch L;
        L.chunk = tiles;
        kernel<<<(unsigned)blocks, kFoldThreads, 0, s>>>(v, L);
        return check_launch("unary_fold_kernel");
    };
    if (weighted) return hub ? go(unary_fold_kernel<real, K, true, true>) : go(unary_fold_kernel<real, K, true, false>);
    return hub ? go(unary_fold_kernel<real, K, false, true>) : go(unary_fold_kernel<real, K, false, false>);
}

}  // namespace lhvi
#include "lhvi_run_impl.cuh"
namespace lhvi {

// ---- dispatch ----------------------------------------------------------------------------------

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool WEIGHTED, int HUB>
static int launch_final(const GroupView<real>& v, cudaStream_t s) {
    auto kernel = factor_spec_kernel<real, K, T, NC, NG, NE, FL, WEIGHTED, HUB>;
    static int resident = 0;                 // one persistent wave; each block owns a contiguous chunk
    if (resident == 0) resident = resident_blocks(kernel);
    long long blocks = (v.n + kSpecThreads - 1) / kSpecThreads;
    if (blocks > resident) blocks = resident;
    if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
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
        // the streaming kernel pays off for long runs of records on one variable (hub groups); a
        // group of distinct variables closes a run per record and is better served below
        if (g->fold != nullptr && g->nc == 1 && (((g->hub_mask >> g->nd) & 1) != 0 || m->T != T))
            return launch_unary_fold<real, K>(m, g, row0, s);
        switch (code) {
            case 10: { const int rc = launch_pure_unary<real, K, T, 0>(m, g, row0, s); if (rc <= 0) return rc; break; }
            case 11: { const int rc = launch_pure_unary<real, K, T, 1>(m, g, row0, s); if (rc <= 0) return rc; break; }
            case 12: { const int rc = launch_pure_unary<real, K, T, 2>(m, g, row0, s); if (rc <= 0) return rc; break; }
            default: break;
        }
        switch (code) {
#define X(NC_, NE_) case NC_ * 10 + NE_: return launch_one<real, K, T, NC_, 0, NE_, kPure>(m, g, row0, s);
            LHVI_SPEC_PURE(X)
#undef X
            default: return 1;
        }
    }
    const int code = g->nc * 100 + g->ng * 10 + g->ne;
    if (g->run_start != nullptr && (code == 200 || code == 201)) {
        const int rc = code == 200 ? launch_run<real, K, T, 0>(m, g, row0, s) : launch_run<real, K, T, 1>(m, g, row0, s);
        if (rc <= 0) return rc;          // 1: not eligible (too many hubs), use the record-major kernel
    }
    switch (code) {
#define X(NC_, NG_, NE_) case NC_ * 100 + NG_ * 10 + NE_: return launch_one<real, K, T, NC_, NG_, NE_, kFull>(m, g, row0, s);
        LHVI_SPEC_FULL(X)
#undef X
        default: return 1;
    }
}

}  // namespace lhvi
