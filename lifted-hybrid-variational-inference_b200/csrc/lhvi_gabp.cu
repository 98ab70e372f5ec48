// Gaussian belief propagation on the device (the config-4 cross-check at full
// size).  The reference's GaBP (GaBP.py:7-216) keeps a dict of (mu, sig) messages between Python
// objects and sweeps them synchronously: every variable -> factor message, then every factor ->
// variable message (run, :139-165).  Here the same sweep runs in information form on index arrays:
//
//   every pairwise factor f over hidden variables (i, j) with log psi_f = -1/2 x' Jf x + hf' x gives
//   two directed message slots e = (f: i -> j) and rev[e] = (f: j -> i);  with the variable totals
//   SP[i] = sum of incoming precisions, SH[i] = sum of incoming potentials (unary factors and
//   evidence-reduced factors are constants jd / hd of the variable),
//
//       a = SP[src] - P[rev],  b = SH[src] - H[rev]                 variable -> factor  (message_rv_to_f, :19-35)
//       P'[e] = Jdd - Jc^2 / (Jss + a),  H'[e] = hd - Jc (hs + b) / (Jss + a)   factor -> variable (:37-136)
//
//   one thread per directed edge; the new totals are accumulated with REDs into a second buffer
//   (Jacobi / flooding schedule, like the reference's two loops), which the launcher seeds with the
//   variables' constants.  Marginals: mean = SH / SP, variance = 1 / SP (get_belief_params, :183-195).
//
// HBM-bound streaming over the edge list: per edge 2 x int32 indices + 5 constants + 2 messages read,
// 2 written, plus the gathers of the source totals and the reverse message (L2 hits on a grid).
#include "lhvi_common.cuh"

namespace lhvi {

template <typename real>
__global__ void __launch_bounds__(256)
gabp_edge_kernel(long long n_edges, const int* __restrict__ src, const int* __restrict__ dst,
                 const int* __restrict__ rev, const real* __restrict__ coef,     // [5][n_edges]: Jss, Jdd, Jc, hs, hd
                 const real* __restrict__ P, const real* __restrict__ H,
                 const real* __restrict__ SP, const real* __restrict__ SH,
                 real* __restrict__ P2, real* __restrict__ H2, real* __restrict__ SP2, real* __restrict__ SH2) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_edges;
         e += (long long)gridDim.x * blockDim.x) {
        const int s = src[e], d = dst[e], r = rev[e];
        const real a = SP[s] - P[r];
        const real b = SH[s] - H[r];
        const real jss = coef[e], jdd = coef[n_edges + e], jc = coef[2 * n_edges + e];
        const real hs = coef[3 * n_edges + e], hd = coef[4 * n_edges + e];
        const real inv = real(1) / (jss + a);
        const real p = jdd - jc * jc * inv;
        const real h = hd - jc * (hs + b) * inv;
        P2[e] = p;
        H2[e] = h;
        atomicAdd(SP2 + d, p);
        atomicAdd(SH2 + d, h);
    }
}

template <typename real>
__global__ void __launch_bounds__(256)
gabp_seed_kernel(long long n, const real* __restrict__ jd, const real* __restrict__ hd,
                 real* __restrict__ SP2, real* __restrict__ SH2) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        SP2[i] = jd[i];
        SH2[i] = hd[i];
    }
}

template <typename real>
__global__ void __launch_bounds__(256)
gabp_marginal_kernel(long long n, const real* __restrict__ SP, const real* __restrict__ SH,
                     real* __restrict__ mean, real* __restrict__ var) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const real v = real(1) / SP[i];
        var[i] = v;
        mean[i] = v * SH[i];
    }
}

static unsigned grid_1d(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = 148ll * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

template <typename real>
static int sweeps_t(const lhvi_gabp* g, int n_sweeps, cudaStream_t s) {
    const long long V = g->n_vars, E = g->n_edges;
    real* P[2] = {(real*)g->P, (real*)g->P + E};
    real* H[2] = {(real*)g->H, (real*)g->H + E};
    real* SP[2] = {(real*)g->SP, (real*)g->SP + V};
    real* SH[2] = {(real*)g->SH, (real*)g->SH + V};
    int cur = g->parity & 1;
    for (int it = 0; it < n_sweeps; ++it) {
        const int nxt = cur ^ 1;
        gabp_seed_kernel<real><<<grid_1d(V), 256, 0, s>>>(V, (const real*)g->jd, (const real*)g->hd, SP[nxt], SH[nxt]);
        if (E > 0)
            gabp_edge_kernel<real><<<grid_1d(E), 256, 0, s>>>(E, g->src, g->dst, g->rev, (const real*)g->coef,
                                                             P[cur], H[cur], SP[cur], SH[cur], P[nxt], H[nxt], SP[nxt], SH[nxt]);
        cur = nxt;
    }
    return check_launch("gabp_edge_kernel");
}

}  // namespace lhvi

using namespace lhvi;

static int gabp_validate(const lhvi_gabp* g, const char* who) {
    if (!g) { set_error("%s: null descriptor", who); return LHVI_EINVAL; }
    if (g->dtype != LHVI_F32 && g->dtype != LHVI_F64) { set_error("%s: dtype %d is neither LHVI_F32 nor LHVI_F64", who, g->dtype); return LHVI_EINVAL; }
    if (g->n_vars < 0 || g->n_edges < 0 || g->n_vars >= (1ll << 31) || g->n_edges >= (1ll << 31)) { set_error("%s: n_vars=%lld n_edges=%lld out of range", who, (long long)g->n_vars, (long long)g->n_edges); return LHVI_ELIMIT; }
    if (g->n_vars > 0 && (!g->jd || !g->hd || !g->SP || !g->SH)) { set_error("%s: null variable buffer (jd/hd/SP/SH)", who); return LHVI_EINVAL; }
    if (g->n_edges > 0 && (!g->src || !g->dst || !g->rev || !g->coef || !g->P || !g->H)) { set_error("%s: null edge buffer (src/dst/rev/coef/P/H)", who); return LHVI_EINVAL; }
    if (g->parity != 0 && g->parity != 1) { set_error("%s: parity=%d must be 0 or 1", who, g->parity); return LHVI_EINVAL; }
    return LHVI_OK;
}

extern "C" int lhvi_gabp_sweeps(const lhvi_gabp* g, int32_t n_sweeps, void* stream) {
    int rc = gabp_validate(g, "lhvi_gabp_sweeps");
    if (rc != LHVI_OK) return rc;
    if (n_sweeps < 0) { set_error("lhvi_gabp_sweeps: negative sweep count"); return LHVI_EINVAL; }
    if (n_sweeps == 0 || g->n_vars == 0) return LHVI_OK;
    return g->dtype == LHVI_F64 ? sweeps_t<double>(g, n_sweeps, (cudaStream_t)stream)
                                : sweeps_t<float>(g, n_sweeps, (cudaStream_t)stream);
}

extern "C" int lhvi_gabp_marginals(const lhvi_gabp* g, void* mean, void* var, void* stream) {
    int rc = gabp_validate(g, "lhvi_gabp_marginals");
    if (rc != LHVI_OK) return rc;
    if (g->n_vars == 0) return LHVI_OK;
    if (!mean || !var) { set_error("lhvi_gabp_marginals: null output"); return LHVI_EINVAL; }
    const long long V = g->n_vars;
    cudaStream_t s = (cudaStream_t)stream;
    if (g->dtype == LHVI_F64)
        gabp_marginal_kernel<double><<<grid_1d(V), 256, 0, s>>>(V, (const double*)g->SP + (g->parity & 1) * V,
                                                               (const double*)g->SH + (g->parity & 1) * V, (double*)mean, (double*)var);
    else
        gabp_marginal_kernel<float><<<grid_1d(V), 256, 0, s>>>(V, (const float*)g->SP + (g->parity & 1) * V,
                                                              (const float*)g->SH + (g->parity & 1) * V, (float*)mean, (float*)var);
    return check_launch("gabp_marginal_kernel");
}
