// Dispatcher of the template-specialised factor kernels (see lhvi_spec_impl.cuh).  Each
// (dtype, K) pair is instantiated in its own translation unit (lhvi_spec_<dtype>_k<K>.cu) so the
// build parallelises.
#include "lhvi_common.cuh"
#include "lhvi_spec_sigs.h"
#include "lhvi_hyb_sigs.h"

namespace lhvi {

int spec_f32_k1(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int spec_f32_k2(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int spec_f32_k3(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int spec_f64_k1(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int spec_f64_k2(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int spec_f64_k3(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
// records with hidden discrete arguments, quadrature degree 10 (lhvi_hyb_impl.cuh)
int hyb_f32_k1(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int hyb_f32_k2(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int hyb_f32_k3(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int hyb_f64_k1(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int hyb_f64_k2(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);
int hyb_f64_k3(const lhvi_model*, const lhvi_group*, int64_t, cudaStream_t);

static bool walk_available(const lhvi_model* m, const lhvi_group* g);

bool spec_available(const lhvi_model* m, const lhvi_group* g) {
    if (g->pot_kind != LHVI_POT_QUADRATIC) return false;     // evaluated point by point: generic kernel
    return walk_available(m, g) || hyb_available(m, g);
}

static bool walk_available(const lhvi_model* m, const lhvi_group* g) {
    if (g->nd != 0 || m->K < 1 || m->K > 3) return false;
    // the streaming unary kernel takes T at run time (it only needs the rule's even moments)
    if (g->fold && g->pure && !g->node && g->nc == 1 && g->ng == 0) return true;
    if (m->T != 3 || !m->rule_symmetric) return false;
    if (g->node) return (g->nc == 1 && g->ng == 0 && g->ne == 0) || (g->nc == 0 && g->ng == 1 && g->ne == 0);
    if (g->pure) {
        if (g->ng != 0) return false;
        switch (g->nc * 10 + g->ne) {
#define X(NC_, NE_) case NC_ * 10 + NE_: return true;
            LHVI_SPEC_PURE(X)
#undef X
            default: return false;
        }
    }
    switch (g->nc * 100 + g->ng * 10 + g->ne) {
#define X(NC_, NG_, NE_) case NC_ * 100 + NG_ * 10 + NE_: return true;
        LHVI_SPEC_FULL(X)
#undef X
        default: return false;
    }
}

int launch_spec(const lhvi_model* m, const lhvi_group* g, int64_t row0, cudaStream_t s) {
    if (!spec_available(m, g)) return 1;
    const bool f64 = m->dtype == LHVI_F64;
    if (!walk_available(m, g)) {
        switch (m->K) {
            case 1: return f64 ? hyb_f64_k1(m, g, row0, s) : hyb_f32_k1(m, g, row0, s);
            case 2: return f64 ? hyb_f64_k2(m, g, row0, s) : hyb_f32_k2(m, g, row0, s);
            case 3: return f64 ? hyb_f64_k3(m, g, row0, s) : hyb_f32_k3(m, g, row0, s);
            default: return 1;
        }
    }
    switch (m->K) {
        case 1: return f64 ? spec_f64_k1(m, g, row0, s) : spec_f32_k1(m, g, row0, s);
        case 2: return f64 ? spec_f64_k2(m, g, row0, s) : spec_f32_k2(m, g, row0, s);
        case 3: return f64 ? spec_f64_k3(m, g, row0, s) : spec_f32_k3(m, g, row0, s);
        default: return 1;
    }
}

}  // namespace lhvi
