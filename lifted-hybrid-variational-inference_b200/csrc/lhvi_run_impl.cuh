// Run-major kernel for full records with two hidden continuous arguments of which one takes only
// a few distinct values in the group (the "hub" argument: the group-level variables of a
// relational model).  Included by lhvi_spec_impl.cuh.
//
// Layout (lhvi_group::run_*): the records are sorted by their *other* hidden argument (the run
// argument), so that all records of one run variable are contiguous; run_start / run_key give the
// runs, run_hid the record's hub as an index into hub_keys.  One thread owns one run at a time:
//   - the run variable's axis tables (nodes, cross densities, normalisers: 6 of the 9 exponentials
//     per component pair) are built once per run and kept in registers,
//   - the axis tables of *every* hub of the group sit in shared memory, built once per block,
//   - the run variable's gradient is summed in registers and leaves with one vector RED per run,
//   - hub gradients are summed in thread-private shared-memory slots ([hub][value][thread]: no
//     conflicts, no atomics, no shuffles) and reduced once per block.
// What remains per record and component is the 3 x 3 grid of log-beliefs (Walk) and the closed-form
// moments of the quadratic log-potential -- about half of the instructions of the record-major
// kernel.  Floors are handled as there: optimistic walk, literal redo when a bound trips.
#pragma once

namespace lhvi {

constexpr int kRunMaxHubs = LHVI_RUN_MAX_HUBS;
#ifndef LHVI_RUN_THREADS
#define LHVI_RUN_THREADS 256
#define LHVI_RUN_BLOCKS 2
#endif
constexpr int kRunThreads = LHVI_RUN_THREADS;     // threads per block / resident blocks per SM: register budget
constexpr int kRunBlocks = LHVI_RUN_BLOCKS;

// The quadrature rule travels in the launch parameters (constant bank): the compiler then uses the
// values as constant operands instead of holding ~25 registers of them.
template <typename real, int T>
struct RunLaunch {
    int n_hubs;
    real xi[T], w0[T], w1[T], w2[T], eq[T];     // nodes, omega, omega xi, omega xi^2, exp(-xi^2)
    real eq_min;
    real cm0, cm2, cm22, cmd, xm;               // sum W, sum W xi^2, sum W xi^2 xi'^2, sum W xi^4 - cm22, max |xi|
};

// Record columns are read once per iteration: load them with an L2 evict-first policy (they still
// live in L1 for the ten accesses of a 128-byte line) so that the 84 MB of records compete less with
// the parameters, gradients and Adam moments for L2 (measured: ~1 us per iteration, within noise).
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_stream(const int* p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p, unsigned long long pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_stream(const double* p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

template <typename real, int K, int T>
struct RunShared {
    static constexpr int TP = (T + 3) / 4 * 4;
    alignas(16) real hq[kRunMaxHubs][K][K][TP];    // q_{k2}(x_{k,t}) of every hub
    alignas(16) real hnrm[kRunMaxHubs][4];         // 1 / (sqrt(2 pi) var_k2)
    alignas(16) real hms[kRunMaxHubs][K][4];       // mu_k, sqrt(2 var_k), 2 var_k, 1 / var_k
    double scratch[(kRunThreads / 32) * (K + 1)];
    double gw[K][kRunThreads];
    real w[K];
    real quad[2 * T];                              // for the literal (checked) path only
    int hkey[kRunMaxHubs];
};

// s_acc: [n_hubs][2K][kRunThreads] thread-private hub sums (dynamic shared memory)
template <typename real, int K, int T, int NE, bool WEIGHTED, int HUBPOS>
__device__ __forceinline__ void
factor_run_body(const GroupView<real>& g, const RunLaunch<real, T>& L, const int H, const BlockSlice bs,
                RunShared<real, K, T>& sh, real* s_acc) {
    using F = Fast<real>;
    constexpr int NC = 2, NG = 0, NCT = 2 + NE, NV = 2 * K;
    using C = Ctx<real, K, T, NC, NG, NE>;
    constexpr int NQ = C::NQ;
    constexpr int RA = 1 - HUBPOS;                  // canonical position of the run argument
    static_assert(K <= 4, "hub normalisers are stored four to a row");

    auto& s_w = sh.w;
    auto& s_quad = sh.quad;
    auto& s_hq = sh.hq;
    auto& s_hnrm = sh.hnrm;
    auto& s_hms = sh.hms;
    auto& s_hkey = sh.hkey;
    auto& s_scratch = sh.scratch;
    auto& s_gw = sh.gw;

    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = tid; i < K; i += blockDim.x) s_w[i] = g.w[i];
    for (int k = 0; k < K; ++k) s_gw[k][tid] = 0.0;
    for (int i = tid; i < H; i += blockDim.x) s_hkey[i] = g.hub_keys[i];
    for (int i = tid; i < H * NV * kRunThreads; i += blockDim.x) s_acc[i] = real(0);
    __syncthreads();
    // axis tables of every hub
    for (int idx = tid; idx < H * K * K * T; idx += blockDim.x) {
        const int h = idx / (K * K * T), k = (idx / (K * T)) % K, k2 = (idx / T) % K, t = idx % T;
        const int key = s_hkey[h];
        real q = L.eq[t];
        if (k2 != k) {
            const real mu_k = g.eta[key + 2 * k], var_k = g.eta[key + 2 * k + 1];
            const real mu_2 = g.eta[key + 2 * k2], var_2 = g.eta[key + 2 * k2 + 1];
            const real u = F::sqrt(real(2) * var_k) * L.xi[t] + (mu_k - mu_2);
            q = F::exp_scaled(real(-0.5) * F::kExpScale * F::rcp(var_2) * (u * u));
        }
        s_hq[h][k][k2][t] = q;
    }
    for (int idx = tid; idx < H * K; idx += blockDim.x) {
        const int h = idx / K, k = idx % K;
        const int key = s_hkey[h];
        const real mu_k = g.eta[key + 2 * k], var_k = g.eta[key + 2 * k + 1];
        const real inv = F::rcp(var_k);
        s_hnrm[h][k] = inv * real(1.0 / kSqrt2Pi);
        s_hms[h][k][0] = mu_k;
        s_hms[h][k][1] = F::sqrt(real(2) * var_k);
        s_hms[h][k][2] = real(2) * var_k;
        s_hms[h][k][3] = inv;
    }
    __syncthreads();

    C c;
    c.eta = g.eta;
    c.s_w = s_w;
    auto own_floor = [](real own) { return own < F::kBFloor; };

    const unsigned long long pol = l2_evict_first_policy();
    const long long stride = (long long)bs.nblocks * blockDim.x;
    for (long long run = (long long)bs.bid * blockDim.x + tid; run < g.n_runs; run += stride) {
        const int keyE = __ldg(g.run_key + run);
        const int r0 = __ldg(g.run_start + run), r1 = __ldg(g.run_start + run + 1);
        // the next run's first record and slot are fetched now and used to warm L1 once this
        // run's records are done (by then the two loads have long arrived)
        const long long nxt = run + stride;
        int rn = -1, keyn = 0;
        if (nxt < g.n_runs) {
            rn = __ldg(g.run_start + nxt);
            keyn = __ldg(g.run_key + nxt);
        }

        // ---- the run variable: parameters and axis tables, once per run
        real muE[K], sdE[K], wnE[K], hvE[K];     // mean, sqrt(2 var), w_k / (sqrt(2 pi) var), -log2(e) / (2 var)
        {
            real slot[NV];
            load_vec<NV>(g.eta + keyE, slot);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                muE[k] = slot[2 * k];
                const real inv = F::rcp(slot[2 * k + 1]);
                hvE[k] = real(-0.5) * F::kExpScale * inv;
                wnE[k] = s_w[k] * (inv * real(1.0 / kSqrt2Pi));
                sdE[k] = F::sqrt(real(2) * slot[2 * k + 1]);
            }
        }
        real qE[K][K][T];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const real dx = sdE[k] * L.xi[t];
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2) {
                    if (k2 == k) {
                        qE[k][k2][t] = L.eq[t];
                    } else {
                        const real u = dx + (muE[k] - muE[k2]);
                        qE[k][k2][t] = F::exp_scaled(hvE[k2] * (u * u));
                    }
                }
            }
        real G1[K], G2[K], af[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { G1[k] = real(0); G2[k] = real(0); }
#pragma unroll
        for (int i = 0; i < K; ++i) af[i] = real(0);
        c.pt.poff[0] = keyE;

        // ---- the run variable's own records, fused (lhvi_group::run_node / run_una_*): its node-entropy
        // record -- F = log b on the variable's own nodes, b from the cross densities that are in
        // registers already -- and one pure unary quadratic factor (F = log psi, evaluated at the nodes
        // like pure_unary_body does).  Same accumulators as the records below; F enters with the sign of
        // `lpsi - lb`, so the node term is a record with lpsi = 0 and the belief term negated.
        if (g.run_node != nullptr) {
            const real wfN = g.run_node[run], nsN = g.run_node[g.n_runs + run];
            if (wfN != real(0) || nsN != real(0)) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    real Lg[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        real b = wnE[0] * qE[k][0][t];
#pragma unroll
                        for (int k2 = 1; k2 < K; ++k2) b += wnE[k2] * qE[k][k2][t];
                        Lg[t] = F::log_belief(b);
                    }
                    if (own_floor(wnE[k] * L.eq_min)) {          // float only: the belief may have underflowed
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            PointCtx<real, 1, 0> pc;
                            pc.poff[0] = keyE;
                            pc.x[0] = sdE[k] * L.xi[t] + muE[k];
                            Lg[t] = checked_log_belief<real, K, 1, 0>(g.eta, s_w, pc);
                        }
                    }
                    real e0 = real(0), e1 = real(0), e2 = real(0);
#pragma unroll
                    for (int t = 0; t < T; ++t) { e0 += L.w0[t] * Lg[t]; e1 += L.w1[t] * Lg[t]; e2 += L.w2[t] * Lg[t]; }
                    const real Ek = e0 * F::kUnit;
                    G1[k] += nsN * e1;
                    G2[k] += nsN * (e2 * F::kUnit - real(0.5) * Ek);
                    af[k] += wfN * Ek;
                }
            }
        }
        if (g.run_una_pot != nullptr) {
            const int up = g.run_una_pot[run];
            if (up >= 0) {
                const real wfU = g.run_una_w[run], gmU = g.run_una_w[g.n_runs + run];
                constexpr real to_unit = real(1) / F::kUnit;
                const real* cf = g.ptab + up;
                const real c0 = __ldg(cf) * to_unit, l0 = __ldg(cf + 1) * to_unit, a0 = __ldg(cf + 2) * to_unit;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    real e0 = real(0), e1 = real(0), e2 = real(0), qmin = real(0);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const real x = sdE[k] * L.xi[t] + muE[k];
                        const real q = c0 + x * (l0 + a0 * x);
                        qmin = F::min(qmin, q);
                        e0 += L.w0[t] * q; e1 += L.w1[t] * q; e2 += L.w2[t] * q;
                    }
                    if (qmin < F::kQFloor) {                     // the 1e-100 floor is active: literal formula
                        e0 = e1 = e2 = real(0);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            const real x = sdE[k] * L.xi[t] + muE[k];
                            const real q = checked_log_psi<real>(c0 + x * (l0 + a0 * x));
                            e0 += L.w0[t] * q; e1 += L.w1[t] * q; e2 += L.w2[t] * q;
                        }
                    }
                    const real Ek = e0 * F::kUnit;
                    G1[k] += gmU * e1;
                    G2[k] += gmU * (e2 * F::kUnit - real(0.5) * Ek);
                    af[k] += wfU * Ek;
                }
            }
        }

        // ---- its records (columns fetched one record ahead)
        int n_pot = 0, n_h = 0;
        real n_wf = real(1), n_gE = real(1), n_gH = real(1);
#define LHVI_RUN_FETCH(RR)                                                        \
        do {                                                                      \
            n_pot = ld_stream(g.pot + (RR), pol);                                        \
            n_h = ld_stream(g.run_hid + (RR), pol);                                        \
            if constexpr (WEIGHTED) {                                             \
                n_wf = ld_stream(g.wf + (RR), pol);                                       \
                n_gE = ld_stream(g.gam + (long long)RA * g.n + (RR), pol);                \
                n_gH = ld_stream(g.gam + (long long)HUBPOS * g.n + (RR), pol);            \
            }                                                                     \
        } while (0)
        if (r0 < r1) LHVI_RUN_FETCH(r0);
#pragma unroll 1
        for (int r = r0; r < r1; ++r) {
            const int pot = n_pot, h = n_h;
            const real wf = n_wf, gamE = n_gE, gamH = n_gH;
            if (r + 1 < r1) LHVI_RUN_FETCH(r + 1);
            // quadratic log-potential reduced by the point evidence, in walk order (axis 0 = run
            // argument, axis 1 = hub argument): c0 + l0 x0 + l1 x1 + a00 x0^2 + a01 x0 x1 + a11 x1^2
            real c0, l0, l1, a00, a01, a11;
            {
                const real* cf = g.ptab + pot;
                constexpr real to_unit = real(1) / F::kUnit;
                real lc[NCT], Ac[NCT][NCT];
                c0 = __ldg(cf) * to_unit;
#pragma unroll
                for (int i = 0; i < NCT; ++i) lc[i] = __ldg(cf + 1 + i) * to_unit;
                int p = 1 + NCT;
#pragma unroll
                for (int i = 0; i < NCT; ++i)
#pragma unroll
                    for (int j = i; j < NCT; ++j) Ac[i][j] = __ldg(cf + p++) * to_unit;
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int j = 2 + e;
                    const real xv = ld_stream(g.ecval + (long long)e * g.n + r, pol);
                    c0 += xv * (lc[j] + Ac[j][j] * xv);
#pragma unroll
                    for (int i = 0; i < j; ++i) lc[i] += Ac[i][j] * xv;
#pragma unroll
                    for (int i = j + 1; i < NCT; ++i) lc[i] += Ac[j][i] * xv;
                }
                l0 = lc[RA]; l1 = lc[HUBPOS];
                a00 = Ac[RA][RA]; a11 = Ac[HUBPOS][HUBPOS]; a01 = Ac[0][1];
            }

            real pk0[K];
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) pk0[k2] = wnE[k2] * s_hnrm[h][k2];

#pragma unroll
            for (int k = 0; k < K; ++k) {
                real qH[K][T];
#pragma unroll
                for (int k2 = 0; k2 < K; ++k2)
#pragma unroll
                    for (int t = 0; t < T; ++t) qH[k2][t] = s_hq[h][k][k2][t];
                const real muH = s_hms[h][k][0], sdH = s_hms[h][k][1], tvH = s_hms[h][k][2];
                const real e = muE[k], sE = sdE[k];

                // ---- log b on the T x T grid; mirror nodes are paired on both axes.  Per outer node:
                // S = sum_t2 w0 L, HS / HD = the weighted pair sums / differences of the inner axis
                real S[T], HS[T], HD[T];
#pragma unroll
                for (int t1 = 0; t1 < T; ++t1) {
                    real pk2[K], Lg[T];
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk2[k2] = pk0[k2] * qE[k][k2][t1];
#pragma unroll
                    for (int t2 = 0; t2 < T; ++t2) {
                        real b = pk2[0] * qH[0][t2];
#pragma unroll
                        for (int k2 = 1; k2 < K; ++k2) b += pk2[k2] * qH[k2][t2];
                        Lg[t2] = F::log_belief(b);
                    }
                    real s0 = real(0), hs = real(0), hd = real(0);
                    if constexpr (T == 3) {              // one pair: its weights are applied at the end
                        hs = Lg[0] + Lg[2];
                        hd = Lg[2] - Lg[0];
                        s0 = L.w0[0] * hs + L.w0[1] * Lg[1];
                    } else {
#pragma unroll
                        for (int t2 = 0; t2 < T / 2; ++t2) {
                            const real sm = Lg[t2] + Lg[T - 1 - t2];
                            s0 += L.w0[t2] * sm;
                            hs += L.w2[t2] * sm;
                            hd += L.w1[T - 1 - t2] * (Lg[T - 1 - t2] - Lg[t2]);
                        }
                        if constexpr ((T & 1) != 0) s0 += L.w0[T / 2] * Lg[T / 2];
                    }
                    S[t1] = s0; HS[t1] = hs; HD[t1] = hd;
                }
                real Elb = real(0), e1 = real(0), e2 = real(0), g1 = real(0), g2 = real(0);
#pragma unroll
                for (int t1 = 0; t1 < T / 2; ++t1) {
                    const real sm = S[t1] + S[T - 1 - t1];
                    Elb += L.w0[t1] * sm;
                    e2 += L.w2[t1] * sm;
                    e1 += L.w1[T - 1 - t1] * (S[T - 1 - t1] - S[t1]);
                    g1 += L.w0[t1] * (HD[t1] + HD[T - 1 - t1]);
                    g2 += L.w0[t1] * (HS[t1] + HS[T - 1 - t1]);
                }
                if constexpr ((T & 1) != 0) {
                    Elb += L.w0[T / 2] * S[T / 2];
                    g1 += L.w0[T / 2] * HD[T / 2];
                    g2 += L.w0[T / 2] * HS[T / 2];
                }
                if constexpr (T == 3) { g1 *= L.w1[2]; g2 *= L.w2[0]; }

                // ---- closed-form quadrature sums of log psi (see factor_spec_kernel)
                const real h0 = a00 * e, h1 = a11 * muH, x01 = a01 * muH;
                const real t0 = (l0 + h0) + x01, t1 = l1 + h1;
                const real P = e * t0 + (muH * t1 + c0);
                const real Q0 = sE * (t0 + h0), Q1 = sdH * (a01 * e + (t1 + h1));
                const real R0 = a00 * (sE * sE), R1 = a11 * tvH, Rs = R0 + R1;
                const real absR = fabs(a01) * (sE * sdH) + (fabs(R0) + fabs(R1));
                const bool redo = own_floor(pk0[k] * (L.eq_min * L.eq_min)) ||
                                  (P - L.xm * ((fabs(Q0) + fabs(Q1)) + L.xm * absR)) < F::kQFloor;
                const real base2 = L.cm2 * P + L.cm22 * Rs;
                real Ek = (L.cm0 * P + L.cm2 * Rs) - Elb;
                real m1E = L.cm2 * Q0 - e1, m1H = L.cm2 * Q1 - g1;
                real m2E = (L.cmd * R0 + base2) - e2, m2H = (L.cmd * R1 + base2) - g2;
                if (redo) {
                    // a floor may be active on this grid: redo it with the literal formulas
                    real pk[K], lin0[NQ];
#pragma unroll
                    for (int k2 = 0; k2 < K; ++k2) pk[k2] = pk0[k2];
#pragma unroll
                    for (int j = 0; j < NQ; ++j) lin0[j] = real(0);
                    lin0[0] = l0; lin0[1] = l1;
#pragma unroll
                    for (int i = 0; i < NQ; ++i)
#pragma unroll
                        for (int j = 0; j < NQ; ++j) c.A[i][j] = real(0);
                    c.A[0][0] = a00; c.A[1][1] = a11; c.A[0][1] = a01;
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        c.w0[t] = L.w0[t]; c.w1[t] = L.w1[t]; c.w2[t] = L.w2[t];
                        c.x[0][t] = sE * L.xi[t] + e;
                        c.x[1][t] = sdH * L.xi[t] + muH;
#pragma unroll
                        for (int k2 = 0; k2 < K; ++k2) { c.q[0][k2][t] = real(0); c.q[1][k2][t] = real(0); }
                    }
                    c.m1[0] = c.m1[1] = c.m2[0] = c.m2[1] = real(0);
                    c.pt.poff[1] = s_hkey[h];
                    Ek = Walk<real, K, T, NC, NG, NE, kFull, true, 0>::run(c, pk, real(1), c0, lin0);
                    m1E = c.m1[0]; m1H = c.m1[1]; m2E = c.m2[0]; m2H = c.m2[1];
                }
                Ek *= F::kUnit;
                // raw sums; the factors -sdev kUnit / var and -1 / var are applied once at the end
                G1[k] += gamE * m1E;
                G2[k] += gamE * (m2E * F::kUnit - real(0.5) * Ek);
                real* ap = s_acc + ((h * NV + 2 * k) * kRunThreads + tid);
                ap[0] += gamH * m1H;
                ap[kRunThreads] += gamH * (m2H * F::kUnit - real(0.5) * Ek);
                af[k] += wf * Ek;
            }
        }
#undef LHVI_RUN_FETCH
        if (rn >= 0) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(g.pot + rn));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(g.run_hid + rn));
            if constexpr (WEIGHTED) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(g.wf + rn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(g.gam + rn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(g.gam + g.n + rn));
            }
            asm volatile("prefetch.global.L1 [%0];" ::"l"(g.eta + keyn));
        }

        // ---- end of the run: the run variable's gradient, one vector RED
        real gvE[NV];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const real inv = real(2) * F::rcp(sdE[k] * sdE[k]);      // 1 / var
            gvE[2 * k] = -(sdE[k] * F::kUnit * G1[k]) * inv;
            gvE[2 * k + 1] = -G2[k] * inv;
        }
        if (r1 > r0) red_vec<NV>(g.grad + keyE, gvE);
        // G_w sums of this thread, in double, kept in shared memory ([k][thread]: private slots)
#pragma unroll
        for (int k = 0; k < K; ++k) s_gw[k][tid] -= (double)af[k];
    }
    // the energy is the same sum weighted by w_k: E = sum_k w_k G_w[k]
    double acc[K + 1];
    acc[K] = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[k] = s_gw[k][tid];
        acc[K] += (double)s_w[k] * acc[k];
    }

    publish_partials(acc, K + 1, s_scratch, g.partials, bs);      // ends with a barrier

    // ---- hub gradients: sum the thread-private slots, one warp per hub at a time
    {
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        for (int h = warp; h < H; h += nwarps) {
            real v[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                real s = real(0);
                const real* src = s_acc + (h * NV + i) * kRunThreads;
#pragma unroll
                for (int j = 0; j < kRunThreads / 32; ++j) s += src[lane + 32 * j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                v[i] = s;
            }
            if (lane == 0) {
                bool any = false;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    v[2 * k] = -(s_hms[h][k][1] * F::kUnit * v[2 * k]) * s_hms[h][k][3];
                    v[2 * k + 1] = -v[2 * k + 1] * s_hms[h][k][3];
                    any = any || v[2 * k] != real(0) || v[2 * k + 1] != real(0);
                }
                if (any) red_vec<NV>(g.grad + s_hkey[h], v);
            }
        }
    }
}

template <typename real, int K, int T, int NE, bool WEIGHTED, int HUBPOS>
__global__ void __launch_bounds__(kRunThreads, kRunBlocks)
factor_run_kernel(const GroupView<real> g, const RunLaunch<real, T> L) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ RunShared<real, K, T> sh;
    factor_run_body<real, K, T, NE, WEIGHTED, HUBPOS>(g, L, L.n_hubs, BlockSlice{(int)blockIdx.x, (int)gridDim.x, nullptr}, sh,
                                                      reinterpret_cast<real*>(s_dyn));
}

template <typename real, int K, int T, int NE>
static int launch_run(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const GroupView<real> v = make_view<real>(m, g, row0);
    if (g->n_hubs < 1 || g->n_hubs > kRunMaxHubs) return 1;
    const size_t dyn = (size_t)g->n_hubs * 2 * K * kRunThreads * sizeof(real);
    if (dyn > 96 * 1024) return 1;
    if (g->n >= (1ll << 31)) { set_error("factor_run_kernel: more than 2^31 records in one group"); return (int)LHVI_ELIMIT; }
    auto go = [&](auto kernel) {
        // (every instantiation has the same function-pointer type, so nothing here may be cached
        // in a static of this generic lambda; the calls are host-side and the launches are replayed
        // from a CUDA graph anyway)
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024));
        if (e != cudaSuccess) { set_error("factor_run_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)LHVI_ECUDA; }
        int dev = 0, per_sm = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRunThreads, dyn);
        long long resident = (long long)(per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
        long long blocks = (g->n_runs + kRunThreads - 1) / kRunThreads;
        if (blocks > resident) blocks = resident;
        if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
        if (blocks < 1) blocks = 1;
        RunLaunch<real, T> L;
        L.n_hubs = g->n_hubs;
        double M0 = 0.0, M2 = 0.0, M4 = 0.0, xm = 0.0, eqm = 1.0;
        for (int t = 0; t < T; ++t) {
            const double x = m->quad_host[t], om = m->quad_host[T + t];
            L.xi[t] = (real)x; L.w0[t] = (real)om; L.w1[t] = (real)(om * x); L.w2[t] = (real)(om * x * x);
            L.eq[t] = (real)::exp(-x * x);
            eqm = ::exp(-x * x) < eqm ? ::exp(-x * x) : eqm;
            M0 += om; M2 += om * x * x; M4 += om * x * x * x * x;
            xm = ::fabs(x) > xm ? ::fabs(x) : xm;
        }
        if (!(M0 > 0.5)) { set_error("factor_run_kernel: lhvi_model::quad_host is not filled in"); return (int)LHVI_EINVAL; }
        L.eq_min = (real)eqm;
        L.cm0 = (real)(M0 * M0); L.cm2 = (real)(M0 * M2); L.cm22 = (real)(M2 * M2);
        L.cmd = (real)(M0 * M4 - M2 * M2); L.xm = (real)xm;
        kernel<<<(unsigned)blocks, kRunThreads, dyn, s>>>(v, L);
        return check_launch("factor_run_kernel");
    };
    const bool weighted = g->weighted != 0;
    if (g->run_hub_arg == 0)
        return weighted ? go(factor_run_kernel<real, K, T, NE, true, 0>) : go(factor_run_kernel<real, K, T, NE, false, 0>);
    return weighted ? go(factor_run_kernel<real, K, T, NE, true, 1>) : go(factor_run_kernel<real, K, T, NE, false, 1>);
}

}  // namespace lhvi
