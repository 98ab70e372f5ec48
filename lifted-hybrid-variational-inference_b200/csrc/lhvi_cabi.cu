// extern "C" entry points of the factor pass, argument validation and error reporting.
// See include/lhvi.h for the contract of every function.
#include <cstdarg>
#include <cstdio>

#include "lhvi_common.cuh"

namespace lhvi {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return LHVI_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LHVI_ECUDA;
}

static int validate(const lhvi_model* m, const lhvi_group* g, int64_t row0) {
    if (!m || !g) { set_error("null model or group descriptor"); return LHVI_EINVAL; }
    if (m->dtype != LHVI_F32 && m->dtype != LHVI_F64) { set_error("dtype %d is neither LHVI_F32 nor LHVI_F64", m->dtype); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (m->T < 1 || m->T > LHVI_MAX_T) { set_error("T=%d out of range 1..%d", m->T, LHVI_MAX_T); return LHVI_ELIMIT; }
    if (!m->quad || !m->eta || !m->w || !m->grad || !m->partials) { set_error("null model buffer (quad/eta/w/grad/partials)"); return LHVI_EINVAL; }
    if (g->n < 0 || row0 < 0) { set_error("negative record count or partial row"); return LHVI_EINVAL; }
    if (g->nd < 0 || g->nc < 0 || g->ng < 0 || g->ne < 0) { set_error("negative argument count"); return LHVI_EINVAL; }
    if (g->n == 0) return LHVI_OK;
    const int nh = g->nd + g->nc;
    if (nh > 0 && !g->poff) { set_error("group has hidden arguments but poff is null"); return LHVI_EINVAL; }
    if (g->ng > 0 && (!g->egval || !g->egvar)) { set_error("group has Gaussian evidence but egval/egvar is null"); return LHVI_EINVAL; }
    if (g->ne > 0 && !g->ecval) { set_error("group has point evidence but ecval is null"); return LHVI_EINVAL; }
    if (g->weighted && (!g->wf || (nh > 0 && !g->gam))) { set_error("weighted group without wf/gam"); return LHVI_EINVAL; }
    if (g->node) {
        if (!g->nscale || !g->wf) { set_error("node group without nscale/wf"); return LHVI_EINVAL; }
        if (g->pure) { set_error("a node group cannot be pure"); return LHVI_EINVAL; }
        if (nh + g->ng != 1 || g->ne != 0) { set_error("node group must have exactly one integrated argument"); return LHVI_EINVAL; }
    } else if (!g->pot || !m->ptab) {
        set_error("factor group without pot/ptab");
        return LHVI_EINVAL;
    }
    if (g->pot_kind != LHVI_POT_QUADRATIC) {
        if (g->pot_kind != LHVI_POT_HARD && g->pot_kind != LHVI_POT_IMAGE_EDGE) { set_error("pot_kind=%d is not a LHVI_POT_* code", g->pot_kind); return LHVI_EINVAL; }
        if (g->node || g->fold || g->run_start) { set_error("pot_kind=%d: node groups, fold and run-major columns are defined for quadratic log-potentials only", g->pot_kind); return LHVI_EINVAL; }
        if (g->pot_kind == LHVI_POT_IMAGE_EDGE && (g->nd != 0 || g->nc + g->ng + g->ne != 2)) { set_error("LHVI_POT_IMAGE_EDGE takes exactly two continuous arguments"); return LHVI_EINVAL; }
        if (g->pot_kind == LHVI_POT_HARD && g->nc + g->ng + g->ne == 0) { set_error("LHVI_POT_HARD without a continuous argument: tabulate the potential instead"); return LHVI_EINVAL; }
    }
    if (g->fold) {
        if (g->node || !g->pure || g->nd != 0 || g->nc != 1 || g->ng != 0) { set_error("fold columns are only defined for pure groups with one hidden continuous argument"); return LHVI_EINVAL; }
        if (g->n_pad < g->n || g->n_pad % 1024 != 0) { set_error("fold columns: n_pad=%lld must be a multiple of 1024 and >= n=%lld", (long long)g->n_pad, (long long)g->n); return LHVI_EINVAL; }
    }
    if ((g->run_node || g->run_una_pot || g->run_una_w) && !g->run_start) { set_error("run_node / run_una_* are columns of a run-major group (run_start is null)"); return LHVI_EINVAL; }
    if ((g->run_una_pot != nullptr) != (g->run_una_w != nullptr)) { set_error("run_una_pot and run_una_w come together"); return LHVI_EINVAL; }
    if (g->cst_n >= (1ll << 31)) { set_error("cst_n=%lld: at most 2^31 - 1 constant records per group", (long long)g->cst_n); return LHVI_ELIMIT; }
    if (g->cst_n < 0 || (g->cst_n > 0 && (!g->fold || !g->cst_q))) { set_error("cst_* columns belong to a streamed group (fold columns) and need cst_q"); return LHVI_EINVAL; }
    if (g->run_start) {
        if (g->node || g->pure || g->nd != 0 || g->nc != 2 || g->ng != 0) { set_error("run-major columns are only defined for full groups with two hidden continuous arguments"); return LHVI_EINVAL; }
        if (!g->run_key || !g->run_hid || !g->hub_keys) { set_error("run-major group without run_key/run_hid/hub_keys"); return LHVI_EINVAL; }
        if (g->n_runs < 1 || g->n_hubs < 1 || g->run_hub_arg < 0 || g->run_hub_arg > 1) { set_error("run-major group: n_runs=%lld n_hubs=%d run_hub_arg=%d", (long long)g->n_runs, g->n_hubs, g->run_hub_arg); return LHVI_EINVAL; }
    }
    return LHVI_OK;
}

}  // namespace lhvi

using namespace lhvi;

extern "C" const char* lhvi_last_error(void) { return g_error; }

extern "C" int lhvi_abi_version(void) { return LHVI_ABI_VERSION; }

extern "C" int lhvi_has_specialisation(const lhvi_model* m, const lhvi_group* g) {
    if (!m || !g) return 0;
    return spec_available(m, g) ? 1 : 0;
}

extern "C" int lhvi_factor_expect_grad(const lhvi_model* m, const lhvi_group* g, int64_t row0,
                                       int force_generic, void* stream) {
    int rc = validate(m, g, row0);
    if (rc != LHVI_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (g->n == 0) {
        // empty group: publish "0 valid rows" in the region's header row
        cudaError_t e = cudaMemsetAsync(m->partials + row0 * (m->K + 1), 0, sizeof(double), s);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync(partials): %s", cudaGetErrorString(e)); return LHVI_ECUDA; }
        return LHVI_OK;
    }
    if (!force_generic) {
        rc = launch_spec(m, g, row0, s);
        if (rc <= 0) return rc;      // launched (0) or failed (<0); 1 means "no specialisation"
    }
    return launch_generic(m, g, row0, s);
}
