// compat="reference": the categorical gradients as the UNMODIFIED reference computes them.
//
// gradient_category_tau (VarInference.py:133-160, LiftedVarInference.py:136-164) builds the axes of the
// OTHER hidden arguments of a factor from `rv.domain` -- the domain of the discrete variable being
// differentiated -- instead of `rv_.domain` (:147-150): another hidden argument b is enumerated over the
// values of a's domain with row k of b's own parameter table as weights (a categorical row, or (mu, var)
// when b is continuous).  The integrand is the usual log(psi + 1e-100) - log(b + 1e-100), evaluated at
// those values.  include/lhvi.h (lhvi_category_grad_reference) states when this is reproduced.
//
// One thread per (record, hidden discrete argument); demo-sized groups, plain loops, double arithmetic
// whatever the model's element type (only the final atomic is in `real`).
#include "lhvi_common.cuh"

namespace lhvi {

struct H2Args {
    double dvals[LHVI_MAX_AXES][LHVI_MAX_DSTATES];
    int xmap[LHVI_MAX_AXES][LHVI_MAX_AXES][LHVI_MAX_DSTATES];
};

template <typename real>
__global__ void __launch_bounds__(128)
category_reference_kernel(const GroupView<real> g, const H2Args h) {
    const int K = g.K, nd = g.nd, nc = g.nc, nh = nd + nc;
    const int nct = nc + g.ne;
    const int ncoef = nct == 0 ? 1 : (nct + 1) * (nct + 2) / 2;
    int cstride[LHVI_MAX_AXES];
    {
        int st = 1;
        for (int a = nd - 1; a >= 0; --a) { cstride[a] = st; st *= g.dims[a]; }
    }
    const long long total = g.n * nd;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / nd;
        const int a = (int)(idx % nd);
        const int D = g.dims[a];
        const double gam = g.weighted ? (double)g.gam[a * g.n + r] : 1.0;
        if (gam == 0.0) continue;
        const real* p[LHVI_MAX_AXES];
        for (int b = 0; b < nh; ++b) p[b] = g.eta + g.poff[b * g.n + r];
        const real* coef0 = g.ptab + g.pot[r];
        double xv[LHVI_MAX_AXES + 4];
        for (int e = 0; e < g.ne; ++e) xv[nc + e] = (double)g.ecval[e * g.n + r];

        int others[LHVI_MAX_AXES];
        int m = 0;
        for (int b = 0; b < nh; ++b)
            if (b != a) others[m++] = b;
        int combos = 1;
        for (int i = 0; i < m; ++i) combos *= D;      // every other argument takes D nodes (the host checked it)

        for (int k = 0; k < K; ++k) {
            for (int d = 0; d < D; ++d) {
                double sum = 0.0;
                for (int c = 0; c < combos; ++c) {
                    int j[LHVI_MAX_AXES], st[LHVI_MAX_AXES];
                    int rest = c;
                    for (int i = m - 1; i >= 0; --i) { j[others[i]] = rest % D; rest /= D; }
                    double W = 1.0;
                    st[a] = d;
                    for (int i = 0; i < m; ++i) {
                        const int b = others[i];
                        if (b < nd) {
                            W *= (double)p[b][k * g.dims[b] + j[b]];
                            st[b] = h.xmap[a][b][j[b]];
                        } else {
                            W *= (double)p[b][2 * k + j[b]];              // j = 0: mu, j = 1: var
                            xv[b - nd] = h.dvals[a][j[b]];
                        }
                    }
                    int cfg = 0;
                    for (int b = 0; b < nd; ++b) cfg += st[b] * cstride[b];
                    const real* cf = coef0 + (long long)cfg * ncoef;
                    double lpsi;
                    if (nct == 0) {
                        lpsi = (double)cf[0];
                    } else {
                        double q = (double)cf[0];
                        for (int i = 0; i < nct; ++i) q += (double)cf[1 + i] * xv[i];
                        int pp = 1 + nct;
                        for (int i = 0; i < nct; ++i)
                            for (int i2 = i; i2 < nct; ++i2) q += (double)cf[pp++] * xv[i] * xv[i2];
                        lpsi = ::log(::exp(q) + kEps);
                    }
                    double bel = 0.0;
                    for (int k2 = 0; k2 < K; ++k2) {
                        double v = (double)g.w[k2] * (double)p[a][k2 * D + d];
                        for (int i = 0; i < m; ++i) {
                            const int b = others[i];
                            if (b < nd) v *= (double)p[b][k2 * g.dims[b] + st[b]];
                            else v *= norm_pdf_d(xv[b - nd], (double)p[b][2 * k2], (double)p[b][2 * k2 + 1]);
                        }
                        bel += v;
                    }
                    sum += W * (lpsi - ::log(bel + kEps));
                }
                atomicAdd(g.grad + g.poff[a * g.n + r] + k * D + d, (real)(-gam * sum));
            }
        }
    }
}

}  // namespace lhvi

using namespace lhvi;

extern "C" int lhvi_category_grad_reference(const lhvi_model* m, const lhvi_group* g, const lhvi_h2* h, void* stream) {
    if (!m || !g || !h) { set_error("lhvi_category_grad_reference: null descriptor"); return LHVI_EINVAL; }
    if (m->dtype != LHVI_F32 && m->dtype != LHVI_F64) { set_error("dtype %d is neither LHVI_F32 nor LHVI_F64", m->dtype); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (g->node || g->pure || g->nd < 1 || g->ng != 0) { set_error("lhvi_category_grad_reference: only full groups with a hidden discrete argument and no Gaussian evidence"); return LHVI_EINVAL; }
    if (g->nd + g->nc > LHVI_MAX_AXES || g->ne > 4) { set_error("lhvi_category_grad_reference: too many arguments"); return LHVI_ELIMIT; }
    if (g->n == 0) return LHVI_OK;
    if (!m->eta || !m->w || !m->grad || !m->ptab || !g->pot || !g->poff || (g->weighted && !g->gam) || (g->ne > 0 && !g->ecval)) {
        set_error("lhvi_category_grad_reference: null buffer");
        return LHVI_EINVAL;
    }
    for (int a = 0; a < g->nd; ++a) {
        const int D = g->dims[a];
        if (D < 1 || D > LHVI_MAX_DSTATES) { set_error("discrete argument with %d states (max %d)", D, LHVI_MAX_DSTATES); return LHVI_ELIMIT; }
        if (g->nc > 0 && D != 2) { set_error("lhvi_category_grad_reference: a %d-state argument next to a continuous one (the reference pairs values and (mu, var) by position: only 2 states line up)", D); return LHVI_EINVAL; }
        for (int b = 0; b < g->nd; ++b) {
            if (b == a) continue;
            if (g->dims[b] != D) { set_error("lhvi_category_grad_reference: discrete arguments with %d and %d states in one factor", D, g->dims[b]); return LHVI_EINVAL; }
            for (int j = 0; j < D; ++j)
                if (h->xmap[a][b][j] < 0 || h->xmap[a][b][j] >= g->dims[b]) { set_error("lhvi_category_grad_reference: value %d of argument %d is not in argument %d's domain (the reference raises here)", j, a, b); return LHVI_EINVAL; }
        }
    }
    H2Args args;
    for (int a = 0; a < LHVI_MAX_AXES; ++a) {
        for (int j = 0; j < LHVI_MAX_DSTATES; ++j) args.dvals[a][j] = h->dvals[a][j];
        for (int b = 0; b < LHVI_MAX_AXES; ++b)
            for (int j = 0; j < LHVI_MAX_DSTATES; ++j) args.xmap[a][b][j] = h->xmap[a][b][j];
    }
    cudaStream_t s = (cudaStream_t)stream;
    long long blocks = (g->n * g->nd + 127) / 128;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (m->dtype == LHVI_F64)
        category_reference_kernel<double><<<(unsigned)blocks, 128, 0, s>>>(make_view<double>(m, g, 0), args);
    else
        category_reference_kernel<float><<<(unsigned)blocks, 128, 0, s>>>(make_view<float>(m, g, 0), args);
    return check_launch("category_reference_kernel");
}
