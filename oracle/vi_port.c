/*
 * ORACLE (test infrastructure, not product code) -- plain C / OpenMP restatement of one pass of
 * the reference's variational-inference loop over the lowered record tables.
 *
 * What it restates (leodd/Lifted-Hybrid-Variational-Inference, all in float64 like the reference):
 *   expectation()          VarInference.py:40-55    tensor-product quadrature / enumeration
 *   rvs_belief()           VarInference.py:336-353  b(x) = sum_k w_k prod_i q_ik(x_i)
 *   norm_pdf()             VarInference.py:26-30    note the 1/(2.506628274631 * var) normaliser
 *   gradient_w_tau()       VarInference.py:57-90    G_w[k]   -= W_f E_k[F]
 *   gradient_mu_var()      VarInference.py:92-131   g_mu, g_var
 *   gradient_category_tau  VarInference.py:133-160  G_c[k,d] (with the other arguments' own domains)
 *   free_energy()          VarInference.py:162-195
 *   lifted weights         LiftedVarInference.py:74,90,131-132,162
 *   Gaussian evidence      C2FVarInference.py:110-113,266-267
 * with F = log(psi + 1e-100) - log(b + 1e-100) evaluated once per grid point (SURVEY section 8,
 * "fused single-pass formulation").  The record layout is lowering.RecordGroup's.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load
 * the library built from this file (oracle/Makefile -> oracle/_build/libvi_oracle.so).  It is
 * pinned to the reference through tests/test_c_port.py (same goldens as the numpy port).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VO_MAX_AXES 6
#define VO_MAX_K 8
#define VO_MAX_NODES 96
#define VO_EPS 1e-100
#define VO_SQRT_2PI 2.506628274631

typedef struct vo_group {
    int nd, nc, ng, ne;
    int dims[VO_MAX_AXES];
    int node, weighted, pure;
    int kind;              /* 0: quadratic log-potential; 1: MLNHardPotential (MLNPotential.py:48-49);
                              2: ImageEdgePotential (Potential.py:419-424), block = [distant, scaling, threshold, v] */
    long long n;
    const int* pot;        /* [n] */
    const int* poff;       /* [(nd+nc)*n] */
    const double* egval;   /* [ng*n] */
    const double* egvar;   /* [ng*n] */
    const double* ecval;   /* [ne*n] */
    const double* wf;      /* [n] */
    const double* gam;     /* [(nd+nc)*n] */
    const double* nscale;  /* [n] */
} vo_group;

static double vo_pdf(double x, double mu, double var) {
    const double u = x - mu;
    return exp(-u * u * 0.5 / var) / (VO_SQRT_2PI * var);
}

/* one record: adds into grad (thread-private), gw[K], *energy */
static void vo_record(const vo_group* g, long long r, int K, int T, const double* qx, const double* qw,
                      const double* ptab, const double* eta, const double* w, double* grad, double* gw,
                      double* energy) {
    const int nh = g->nd + g->nc, n_ax = nh + g->ng, nct = g->nc + g->ng + g->ne;
    const int ncoef = nct == 0 ? 1 : (nct + 1) * (nct + 2) / 2;
    int size[VO_MAX_AXES], noff[VO_MAX_AXES + 1], cstride[VO_MAX_AXES];
    noff[0] = 0;
    for (int a = 0; a < n_ax; ++a) {
        size[a] = a < g->nd ? g->dims[a] : T;
        noff[a + 1] = noff[a] + size[a];
    }
    {
        int st = 1;
        for (int a = g->nd - 1; a >= 0; --a) { cstride[a] = st; st *= g->dims[a]; }
    }
    const double wf = g->wf[r];
    const double nscale = g->node ? g->nscale[r] : 1.0;
    double xs[VO_MAX_NODES], wt[VO_MAX_NODES], qd[VO_MAX_NODES * VO_MAX_K];

    for (int k = 0; k < K; ++k) {
        double mu_k[VO_MAX_AXES], var_k[VO_MAX_AXES];
        for (int a = 0; a < g->nd; ++a) {                    /* enumeration over the domain */
            const int D = g->dims[a];
            const double* p = eta + g->poff[a * g->n + r];
            for (int d = 0; d < D; ++d) {
                wt[noff[a] + d] = p[k * D + d];
                for (int k2 = 0; k2 < K; ++k2) qd[(noff[a] + d) * K + k2] = p[k2 * D + d];
            }
        }
        for (int c = 0; c < g->nc; ++c) {                    /* Gauss-Hermite nodes under component k */
            const int a = g->nd + c;
            const double* p = eta + g->poff[a * g->n + r];
            const double mu = p[2 * k], var = p[2 * k + 1], s = sqrt(2.0 * var);
            mu_k[c] = mu;
            var_k[c] = var;
            for (int t = 0; t < T; ++t) {
                const double x = s * qx[t] + mu;
                xs[noff[a] + t] = x;
                wt[noff[a] + t] = qw[t];
                for (int k2 = 0; k2 < K; ++k2) qd[(noff[a] + t) * K + k2] = vo_pdf(x, p[2 * k2], p[2 * k2 + 1]);
            }
        }
        for (int j = 0; j < g->ng; ++j) {                    /* fixed Gaussian evidence argument */
            const int a = nh + j;
            const double val = g->egval[j * g->n + r], var = g->egvar[j * g->n + r], s = sqrt(2.0 * var);
            for (int t = 0; t < T; ++t) {
                const double x = s * qx[t] + val;
                xs[noff[a] + t] = x;
                wt[noff[a] + t] = qw[t];
                const double q = vo_pdf(x, val, var);
                for (int k2 = 0; k2 < K; ++k2) qd[(noff[a] + t) * K + k2] = q;
            }
        }

        int idx[VO_MAX_AXES];
        for (int a = 0; a < n_ax; ++a) idx[a] = 0;
        double Ek = 0.0, am[VO_MAX_AXES], av[VO_MAX_AXES];
        for (int c = 0; c < g->nc; ++c) { am[c] = 0.0; av[c] = 0.0; }
        for (;;) {
            double lb = 0.0;
            if (!g->pure) {
                double b = 0.0;
                for (int k2 = 0; k2 < K; ++k2) {
                    double p = w[k2];
                    for (int a = 0; a < n_ax; ++a) p *= qd[(noff[a] + idx[a]) * K + k2];
                    b += p;
                }
                lb = log(b + VO_EPS);
            }
            double F;
            if (g->node) {
                F = lb;
            } else {
                int cfg = 0;
                for (int a = 0; a < g->nd; ++a) cfg += idx[a] * cstride[a];
                const double* cf = ptab + g->pot[r] + (long long)cfg * ncoef;
                double lpsi;
                if (nct == 0) {
                    lpsi = cf[0];                             /* table entries hold log(psi + 1e-100) */
                } else {
                    double xv[VO_MAX_AXES + 8];
                    for (int i = 0; i < g->nc + g->ng; ++i) xv[i] = xs[noff[g->nd + i] + idx[g->nd + i]];
                    for (int j = 0; j < g->ne; ++j) xv[g->nc + g->ng + j] = g->ecval[j * g->n + r];
                    if (g->kind == 2) {
                        const double d = fabs(xv[0] - xv[1]);
                        lpsi = log(d * cf[0] + (d > cf[2] ? cf[3] : exp(-d / cf[1])) + VO_EPS);
                    } else {
                        double q = cf[0];
                        for (int i = 0; i < nct; ++i) q += cf[1 + i] * xv[i];
                        int p = 1 + nct;
                        for (int i = 0; i < nct; ++i)
                            for (int j = i; j < nct; ++j) q += cf[p++] * xv[i] * xv[j];
                        lpsi = g->kind == 1 ? log((q > 0 ? 1.0 : 0.0) + VO_EPS) : log(exp(q) + VO_EPS);
                    }
                }
                F = lpsi - lb;
            }
            double W = 1.0;
            for (int a = 0; a < n_ax; ++a) W *= wt[noff[a] + idx[a]];
            const double S = W * F;
            Ek += S;
            for (int c = 0; c < g->nc; ++c) {
                const double dx = xs[noff[g->nd + c] + idx[g->nd + c]] - mu_k[c];
                am[c] += S * dx;
                av[c] += S * (dx * dx - var_k[c]);
            }
            for (int a = 0; a < g->nd; ++a) {
                double Wo = 1.0;
                for (int a2 = 0; a2 < n_ax; ++a2)
                    if (a2 != a) Wo *= wt[noff[a2] + idx[a2]];
                const double gam = g->gam[a * g->n + r] * nscale;
                grad[g->poff[a * g->n + r] + k * g->dims[a] + idx[a]] -= gam * Wo * F;
            }
            int a = n_ax - 1;
            while (a >= 0 && ++idx[a] == size[a]) { idx[a] = 0; --a; }
            if (a < 0) break;
        }
        for (int c = 0; c < g->nc; ++c) {
            const int a = g->nd + c;
            const double gam = g->gam[a * g->n + r] * nscale;
            double* dst = grad + g->poff[a * g->n + r];
            dst[2 * k] -= gam * am[c] / var_k[c];
            dst[2 * k + 1] -= gam * av[c] / (2.0 * var_k[c] * var_k[c]);
        }
        gw[k] -= wf * Ek;
        *energy -= wf * w[k] * Ek;
    }
}

int vo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* grad[n_param], g_w[K], energy[1] are overwritten.  Returns 0, or -1 on a limit / allocation error. */
int vo_grad_pass(int K, int T, const double* quad, const double* ptab, const double* eta, const double* w,
                 long long n_param, const vo_group* groups, int n_groups, double* grad, double* g_w,
                 double* energy, int threads) {
    if (K < 1 || K > VO_MAX_K || T < 1) return -1;
    for (int gi = 0; gi < n_groups; ++gi) {
        const vo_group* g = &groups[gi];
        int nodes = 0;
        if (g->nd + g->nc + g->ng > VO_MAX_AXES) return -1;
        for (int a = 0; a < g->nd + g->nc + g->ng; ++a) nodes += a < g->nd ? g->dims[a] : T;
        if (nodes > VO_MAX_NODES) return -1;
    }
    if (threads < 1) threads = vo_max_threads();
    double* priv = (double*)calloc((size_t)threads * (size_t)(n_param + VO_MAX_K + 1), sizeof(double));
    if (!priv) return -1;
    const double* qx = quad;
    const double* qw = quad + T;
    const long long stride = n_param + VO_MAX_K + 1;
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double* mygrad = priv + (size_t)tid * stride;
        double* mygw = mygrad + n_param;
        double* myen = mygw + VO_MAX_K;
        for (int gi = 0; gi < n_groups; ++gi) {
            const vo_group* g = &groups[gi];
#pragma omp for schedule(static) nowait
            for (long long r = 0; r < g->n; ++r)
                vo_record(g, r, K, T, qx, qw, ptab, eta, w, mygrad, mygw, myen);
        }
    }
    /* fixed-order reduction over the thread-private buffers */
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long long i = 0; i < n_param; ++i) {
        double s = 0.0;
        for (int t = 0; t < threads; ++t) s += priv[(size_t)t * stride + i];
        grad[i] = s;
    }
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int t = 0; t < threads; ++t) s += priv[(size_t)t * stride + n_param + k];
        g_w[k] = s;
    }
    double e = 0.0;
    for (int t = 0; t < threads; ++t) e += priv[(size_t)t * stride + n_param + VO_MAX_K];
    *energy = e;
    free(priv);
    return 0;
}
