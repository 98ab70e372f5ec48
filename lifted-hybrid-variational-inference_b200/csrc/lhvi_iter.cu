// extern "C" entry point of the persistent iteration kernel (lhvi_iter_impl.cuh): validation and the
// dispatch over (dtype, K).  Each pair is instantiated in its own translation unit.
#include "lhvi_common.cuh"

namespace lhvi {
#define LHVI_DECL(NAME) \
    int NAME(const lhvi_model*, const lhvi_group*, int, const lhvi_exchange*, const lhvi_optim*, int, int, cudaStream_t);
LHVI_DECL(iter_f32_k1) LHVI_DECL(iter_f32_k2) LHVI_DECL(iter_f32_k3)
LHVI_DECL(iter_f64_k1) LHVI_DECL(iter_f64_k2) LHVI_DECL(iter_f64_k3)
#undef LHVI_DECL
}  // namespace lhvi

using namespace lhvi;

static int iterate_impl(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups, const lhvi_exchange* x,
                        const lhvi_optim* o, int32_t n_iter, int probe_only, void* stream) {
    if (!m || (n_groups > 0 && !groups) || !o) { set_error("lhvi_iterate: null model, group table or optimiser descriptor"); return LHVI_EINVAL; }
    if (m->dtype != LHVI_F32 && m->dtype != LHVI_F64) { set_error("dtype %d is neither LHVI_F32 nor LHVI_F64", m->dtype); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (n_groups < 0 || n_iter < 0) { set_error("lhvi_iterate: negative group or iteration count"); return LHVI_EINVAL; }
    if (probe_only == 0) {
        if (!m->quad || !m->eta || !m->w || !m->grad || !m->partials) { set_error("null model buffer (quad/eta/w/grad/partials)"); return LHVI_EINVAL; }
        if (!o->tau || !o->mom1 || !o->mom2 || !o->wstate || !o->step) { set_error("lhvi_iterate: null optimiser buffer"); return LHVI_EINVAL; }
        if (o->n_vars > 0 && (!o->var_kind || !o->var_dim || !o->var_off)) { set_error("lhvi_iterate: null variable table"); return LHVI_EINVAL; }
        if (o->n_owned < 0 || o->n_owned > o->n_vars) { set_error("lhvi_iterate: n_owned=%lld outside 0..n_vars=%lld", (long long)o->n_owned, (long long)o->n_vars); return LHVI_EINVAL; }
        if (x != nullptr && x->world > 1) {
            if (x->world > LHVI_MAX_PEERS) { set_error("lhvi_iterate: world=%d exceeds LHVI_MAX_PEERS=%d", x->world, LHVI_MAX_PEERS); return LHVI_ELIMIT; }
            if (x->rank < 0 || x->rank >= x->world) { set_error("lhvi_iterate: rank %d outside world %d", x->rank, x->world); return LHVI_EINVAL; }
            if (x->n_idx < 0 || (x->n_idx > 0 && !x->idx) || !x->seq || !x->status) { set_error("lhvi_iterate: incomplete exchange descriptor"); return LHVI_EINVAL; }
            for (int p = 0; p < x->world; ++p)
                if (!x->recv[p] || !x->flags[p]) { set_error("lhvi_iterate: peer %d has no mapped buffer", p); return LHVI_EINVAL; }
        }
        if (n_iter == 0) return LHVI_OK;
    }
    if (m->K > 3) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f64 = m->dtype == LHVI_F64;
    switch (m->K) {
        case 1: return f64 ? iter_f64_k1(m, groups, n_groups, x, o, n_iter, probe_only, s) : iter_f32_k1(m, groups, n_groups, x, o, n_iter, probe_only, s);
        case 2: return f64 ? iter_f64_k2(m, groups, n_groups, x, o, n_iter, probe_only, s) : iter_f32_k2(m, groups, n_groups, x, o, n_iter, probe_only, s);
        default: return f64 ? iter_f64_k3(m, groups, n_groups, x, o, n_iter, probe_only, s) : iter_f32_k3(m, groups, n_groups, x, o, n_iter, probe_only, s);
    }
}

extern "C" int lhvi_iterate(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups, const lhvi_exchange* x,
                            const lhvi_optim* opt, int32_t n_iter, void* stream) {
    return iterate_impl(m, groups, n_groups, x, opt, n_iter, 0, stream);
}

extern "C" int lhvi_iterate_blocks(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups,
                                   const lhvi_exchange* x) {
    lhvi_optim none = {};
    const int rc = iterate_impl(m, groups, n_groups, x, &none, 1, 2, nullptr);
    return rc == 1 ? 0 : rc;      // "unsupported" (1) would read as one block
}

extern "C" int lhvi_iterate_supported(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups,
                                      const lhvi_exchange* x) {
    lhvi_optim none = {};
    return iterate_impl(m, groups, n_groups, x, &none, 1, 1, nullptr) == 0 ? 1 : 0;
}
