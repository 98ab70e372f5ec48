// Generic factor kernel: any signature within the LHVI_MAX_* limits, runtime K / T / arity.
//
// One thread per factor record.  Per component k the thread tabulates, for every axis
// (hidden discrete, hidden continuous, Gaussian evidence), the nodes, the quadrature /
// categorical weights and the K cross-densities q_{k'}(x_{k,t}); it then walks the tensor
// product once, evaluating F = log(psi+1e-100) - log(b+1e-100) a single time per point and
// feeding every accumulator from it ("fused single-pass formulation";
// replaces the 1 + 2*#cont + sum D walks of VarInference.py:57-160).
//
// This kernel keeps its tables in local memory and is the correctness fallback; the
// signatures that matter for throughput have register-resident specialisations in
// lhvi_spec.cu.
#include "lhvi_common.cuh"

namespace lhvi {

constexpr int kGenericThreads = 128;

template <typename real>
__global__ void __launch_bounds__(kGenericThreads)
factor_generic_kernel(const GroupView<real> g) {
    using M = Math<real>;
    const int K = g.K, T = g.T;
    const int nh = g.nd + g.nc;
    const int n_ax = nh + g.ng;
    const int nct = g.nc + g.ng + g.ne;
    const int ncoef = nct == 0 ? 1 : (nct + 1) * (nct + 2) / 2;

    __shared__ real s_quad[2 * LHVI_MAX_T];
    __shared__ real s_w[LHVI_MAX_K];
    __shared__ double s_scratch[(kGenericThreads / 32) * (LHVI_MAX_K + 1)];
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) s_quad[i] = g.quad[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_w[i] = g.w[i];
    __syncthreads();
    const real* qx = s_quad;
    const real* qw = s_quad + T;

    // axis geometry (identical for every record of the group)
    int size[LHVI_MAX_AXES], noff[LHVI_MAX_AXES + 1], goff[LHVI_MAX_AXES + 1], cstride[LHVI_MAX_AXES];
    noff[0] = 0;
    goff[0] = 0;
    for (int a = 0; a < n_ax; ++a) {
        size[a] = a < g.nd ? g.dims[a] : T;
        noff[a + 1] = noff[a] + size[a];
    }
    for (int a = 0; a < nh; ++a) goff[a + 1] = goff[a] + K * (a < g.nd ? g.dims[a] : 2);
    {
        int st = 1;
        for (int a = g.nd - 1; a >= 0; --a) { cstride[a] = st; st *= g.dims[a]; }
    }

    double acc[LHVI_MAX_K + 1];     // G_w[k] ..., energy  (thread-local, block-reduced at the end)
    for (int i = 0; i <= K; ++i) acc[i] = 0.0;

    real xs[LHVI_MAX_NODES];                    // node positions (continuous axes)
    real wt[LHVI_MAX_NODES];                    // node weights under component k
    real qd[LHVI_MAX_NODES * LHVI_MAX_K];       // qd[(noff[a]+t)*K + k'] cross densities
    real gacc[LHVI_MAX_GACC];                   // parameter-gradient accumulators

    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < g.n;
         r += (long long)gridDim.x * blockDim.x) {
        const real wf = (g.weighted || g.node) ? g.wf[r] : real(1);
        const real nscale = g.node ? g.nscale[r] : real(1);
        const real* coef0 = g.node ? nullptr : g.ptab + g.pot[r];
        for (int i = 0; i < goff[nh]; ++i) gacc[i] = real(0);

        for (int k = 0; k < K; ++k) {
            // ---- tabulate the axes under component k
            for (int a = 0; a < g.nd; ++a) {
                const int D = g.dims[a];
                const real* p = g.eta + g.poff[a * g.n + r];
                for (int d = 0; d < D; ++d) {
                    wt[noff[a] + d] = p[k * D + d];
                    for (int k2 = 0; k2 < K; ++k2) qd[(noff[a] + d) * K + k2] = p[k2 * D + d];
                }
            }
            real mu_k[LHVI_MAX_AXES], var_k[LHVI_MAX_AXES];
            for (int c = 0; c < g.nc; ++c) {
                const int a = g.nd + c;
                const real* p = g.eta + g.poff[a * g.n + r];
                const real mu = p[2 * k], var = p[2 * k + 1];
                mu_k[c] = mu;
                var_k[c] = var;
                const real s = M::sqrt(real(2) * var);
                for (int t = 0; t < T; ++t) {
                    const real x = s * qx[t] + mu;
                    xs[noff[a] + t] = x;
                    wt[noff[a] + t] = qw[t];
                    for (int k2 = 0; k2 < K; ++k2)
                        qd[(noff[a] + t) * K + k2] = norm_pdf<real>(x, p[2 * k2], p[2 * k2 + 1]);
                }
            }
            for (int j = 0; j < g.ng; ++j) {
                const int a = nh + j;
                const real val = g.egval[j * g.n + r], var = g.egvar[j * g.n + r];
                const real s = M::sqrt(real(2) * var);
                for (int t = 0; t < T; ++t) {
                    const real x = s * qx[t] + val;
                    xs[noff[a] + t] = x;
                    wt[noff[a] + t] = qw[t];
                    const real q = norm_pdf<real>(x, val, var);
                    for (int k2 = 0; k2 < K; ++k2) qd[(noff[a] + t) * K + k2] = q;
                }
            }

            // ---- one walk over the tensor-product grid
            int idx[LHVI_MAX_AXES];
            for (int a = 0; a < n_ax; ++a) idx[a] = 0;
            real Ek = real(0);
            real am[LHVI_MAX_AXES], av[LHVI_MAX_AXES];
            for (int c = 0; c < g.nc; ++c) { am[c] = real(0); av[c] = real(0); }

            while (true) {
                real b = real(0);
                for (int k2 = 0; k2 < K; ++k2) {
                    real p = s_w[k2];
                    for (int a = 0; a < n_ax; ++a) p *= qd[(noff[a] + idx[a]) * K + k2];
                    b += p;
                }
                real lb;
                if (g.pure) {
                    lb = real(0);          // unary split: -log b is carried by the node record
                } else if (M::belief_underflow(b)) {
                    // float only: redo this point's belief in double from the parameters
                    double bd = 0.0;
                    for (int k2 = 0; k2 < K; ++k2) {
                        double p = (double)s_w[k2];
                        for (int a = 0; a < g.nd; ++a)
                            p *= (double)g.eta[g.poff[a * g.n + r] + k2 * g.dims[a] + idx[a]];
                        for (int c = 0; c < g.nc; ++c) {
                            const real* pp = g.eta + g.poff[(g.nd + c) * g.n + r];
                            p *= norm_pdf_d((double)xs[noff[g.nd + c] + idx[g.nd + c]],
                                            (double)pp[2 * k2], (double)pp[2 * k2 + 1]);
                        }
                        for (int j = 0; j < g.ng; ++j)
                            p *= norm_pdf_d((double)xs[noff[nh + j] + idx[nh + j]],
                                            (double)g.egval[j * g.n + r], (double)g.egvar[j * g.n + r]);
                        bd += p;
                    }
                    lb = (real)::log(bd + kEps);
                } else {
                    lb = M::log_belief(b);
                }

                real F;
                if (g.node) {
                    F = lb;
                } else {
                    int cfg = 0;
                    for (int a = 0; a < g.nd; ++a) cfg += idx[a] * cstride[a];
                    const real* cf = coef0 + cfg * ncoef;
                    real lpsi;
                    if (nct == 0) {
                        lpsi = cf[0];
                    } else {
                        real xv[LHVI_MAX_AXES + 4];
                        for (int i = 0; i < g.nc + g.ng; ++i) xv[i] = xs[noff[g.nd + i] + idx[g.nd + i]];
                        for (int j = 0; j < g.ne; ++j) xv[g.nc + g.ng + j] = g.ecval[j * g.n + r];
                        if (g.pot_kind == LHVI_POT_IMAGE_EDGE) {
                            // psi = d distant_cof + (d > max_threshold ? v : exp(-d / scaling_cof)), d = |x0 - x1|
                            // (Potential.py:419-424); evaluated in double: psi is not an exponential, so the
                            // float shortcut of log_psi does not apply
                            const double d = fabs((double)xv[0] - (double)xv[1]);
                            const double psi = d * (double)cf[0] + (d > (double)cf[2] ? (double)cf[3] : ::exp(-d / (double)cf[1]));
                            lpsi = (real)::log(psi + kEps);
                        } else {
                            real q = cf[0];
                            for (int i = 0; i < nct; ++i) q += cf[1 + i] * xv[i];
                            int p = 1 + nct;
                            for (int i = 0; i < nct; ++i)
                                for (int j = i; j < nct; ++j) q += cf[p++] * xv[i] * xv[j];
                            // hard formula: psi = 1 if formula(x) > 0 else 0 (MLNPotential.py:48-49)
                            lpsi = g.pot_kind == LHVI_POT_HARD ? (q > real(0) ? real(0) : (real)kLogEps) : M::log_psi(q);
                        }
                    }
                    F = lpsi - lb;
                }

                real W = real(1);
                for (int a = 0; a < n_ax; ++a) W *= wt[noff[a] + idx[a]];
                const real S = W * F;
                Ek += S;
                for (int c = 0; c < g.nc; ++c) {
                    const real dx = xs[noff[g.nd + c] + idx[g.nd + c]] - mu_k[c];
                    am[c] += S * dx;
                    av[c] += S * (dx * dx - var_k[c]);
                }
                for (int a = 0; a < g.nd; ++a) {
                    real Wo = real(1);
                    for (int a2 = 0; a2 < n_ax; ++a2)
                        if (a2 != a) Wo *= wt[noff[a2] + idx[a2]];
                    gacc[goff[a] + k * g.dims[a] + idx[a]] -= Wo * F;
                }

                // mixed-radix increment, last axis fastest
                int a = n_ax - 1;
                while (a >= 0 && ++idx[a] == size[a]) { idx[a] = 0; --a; }
                if (a < 0) break;
            }

            for (int c = 0; c < g.nc; ++c) {
                const real inv = M::rcp(var_k[c]);
                gacc[goff[g.nd + c] + 2 * k] = -am[c] * inv;
                gacc[goff[g.nd + c] + 2 * k + 1] = -av[c] * (real(0.5) * inv * inv);
            }
            acc[k] -= (double)(wf * Ek);
            acc[K] -= (double)(wf * s_w[k] * Ek);
        }

        // ---- scatter the parameter gradients
        for (int a = 0; a < nh; ++a) {
            if (a < g.nd && g.no_cat) continue;          // compat="reference": lhvi_category_grad_reference
            const real gam = (g.weighted && !g.node ? g.gam[a * g.n + r] : real(1)) * nscale;
            if (gam == real(0)) continue;
            real* dst = g.grad + g.poff[a * g.n + r];
            const int cnt = goff[a + 1] - goff[a];
            for (int i = 0; i < cnt; ++i) atomicAdd(dst + i, gam * gacc[goff[a] + i]);
        }
    }

    publish_partials(acc, K + 1, s_scratch, g.partials);
}

template <typename real>
static int launch_generic_t(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    GroupView<real> v = make_view<real>(m, g, row0);
    long long blocks = (g->n + kGenericThreads - 1) / kGenericThreads;
    if (blocks > LHVI_PARTIAL_ROWS - 1) blocks = LHVI_PARTIAL_ROWS - 1;
    if (blocks < 1) blocks = 1;
    factor_generic_kernel<real><<<(unsigned)blocks, kGenericThreads, 0, s>>>(v);
    return check_launch("factor_generic_kernel");
}

int launch_generic(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    const int K = m->K, T = m->T;
    const int nh = g->nd + g->nc, n_ax = nh + g->ng;
    if (n_ax > LHVI_MAX_AXES) { set_error("factor with %d integrated arguments (max %d)", n_ax, LHVI_MAX_AXES); return LHVI_ELIMIT; }
    if (g->ne > 4) { set_error("factor with %d point-evidence continuous arguments (max 4)", g->ne); return LHVI_ELIMIT; }
    int nodes = 0, gsz = 0;
    for (int a = 0; a < n_ax; ++a) {
        int sz = a < g->nd ? g->dims[a] : T;
        if (a < g->nd && (sz < 1 || sz > LHVI_MAX_DSTATES)) { set_error("discrete argument with %d states (max %d)", sz, LHVI_MAX_DSTATES); return LHVI_ELIMIT; }
        nodes += sz;
        if (a < nh) gsz += K * (a < g->nd ? sz : 2);
    }
    if (nodes > LHVI_MAX_NODES) { set_error("sum of axis sizes %d exceeds %d", nodes, LHVI_MAX_NODES); return LHVI_ELIMIT; }
    if (gsz > LHVI_MAX_GACC) { set_error("gradient accumulator size %d exceeds %d", gsz, LHVI_MAX_GACC); return LHVI_ELIMIT; }
    return m->dtype == LHVI_F64 ? launch_generic_t<double>(m, g, row0, s)
                                : launch_generic_t<float>(m, g, row0, s);
}

}  // namespace lhvi
