// Elementwise kernels around the factor pass: ELBO / G_w reduction, the optimiser step
// (softmax Jacobians + Adam or SGD + variance clip + re-normalisation) and batched belief
// queries.  All are HBM-streaming kernels over the flat parameter vector.
#include <cstring>

#include "lhvi_opt_impl.cuh"


using namespace lhvi;

extern "C" int lhvi_elbo_reduce(const lhvi_model* m, int64_t rows, void* stream) {
    if (!m || !m->partials || !m->grad || rows < 0) { set_error("lhvi_elbo_reduce: null buffer or negative rows"); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_elbo_reduce: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    if (m->dtype == LHVI_F64)
        elbo_reduce_kernel<double><<<1, 1024, 0, s>>>(m->partials, regions, m->K, (double*)m->grad + m->n_param);
    else
        elbo_reduce_kernel<float><<<1, 1024, 0, s>>>(m->partials, regions, m->K, (float*)m->grad + m->n_param);
    return check_launch("elbo_reduce_kernel");
}

template <typename real>
static int finish_t(const lhvi_model* m, long long regions, double* step, double b1, double b2,
                    const lhvi_exchange* x, cudaStream_t s) {
    FinishArgs<real> a;
    a.partials = m->partials; a.regions = regions; a.K = m->K;
    a.grad = (real*)m->grad; a.n_param = m->n_param;
    a.step = step; a.b1 = b1; a.b2 = b2;
    a.world = 1; a.rank = 0; a.n_idx = 0; a.idx = nullptr; a.seq = nullptr; a.status = nullptr;
    for (int p = 0; p < LHVI_MAX_PEERS; ++p) { a.recv[p] = nullptr; a.flags[p] = nullptr; }
    unsigned blocks = 1;
    if (x != nullptr && x->world > 1) {
        a.world = x->world; a.rank = x->rank; a.n_idx = x->n_idx; a.idx = x->idx;
        a.seq = (unsigned long long*)x->seq; a.status = x->status;
        for (int p = 0; p < x->world; ++p) {
            a.recv[p] = (real*)x->recv[p];
            a.flags[p] = (unsigned long long*)x->flags[p];
        }
        blocks = (unsigned)x->blocks;
    }
    finish_kernel<real><<<blocks, kFinishThreads, 0, s>>>(a);
    return check_launch("finish_kernel");
}

extern "C" int lhvi_finish(const lhvi_model* m, int64_t rows, double* step, double b1, double b2,
                           const lhvi_exchange* x, void* stream) {
    if (!m || !m->partials || !m->grad || rows < 0) { set_error("lhvi_finish: null buffer or negative rows"); return LHVI_EINVAL; }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_finish: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    if (x != nullptr && x->world > 1) {
        if (x->world > LHVI_MAX_PEERS) { set_error("lhvi_finish: world=%d exceeds LHVI_MAX_PEERS=%d", x->world, LHVI_MAX_PEERS); return LHVI_ELIMIT; }
        if (x->rank < 0 || x->rank >= x->world) { set_error("lhvi_finish: rank %d outside world %d", x->rank, x->world); return LHVI_EINVAL; }
        if (x->blocks < 1 || x->blocks > 32) { set_error("lhvi_finish: blocks=%d out of range 1..32", x->blocks); return LHVI_EINVAL; }
        if (x->n_idx < 0 || (x->n_idx > 0 && !x->idx) || !x->seq || !x->status) { set_error("lhvi_finish: incomplete exchange descriptor"); return LHVI_EINVAL; }
        for (int p = 0; p < x->world; ++p)
            if (!x->recv[p] || !x->flags[p]) { set_error("lhvi_finish: peer %d has no mapped buffer", p); return LHVI_EINVAL; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    return m->dtype == LHVI_F64 ? finish_t<double>(m, regions, step, b1, b2, x, s)
                                : finish_t<float>(m, regions, step, b1, b2, x, s);
}

// ---- peer-visible buffers (CUDA IPC) ---------------------------------------------------------

static int cuda_fail(const char* what, cudaError_t e) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return LHVI_ECUDA;
}

extern "C" int lhvi_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
    if (bytes <= 0 || !ptr || !handle) { set_error("lhvi_peer_alloc: bad argument"); return LHVI_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == LHVI_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return cuda_fail("lhvi_peer_alloc: cudaMalloc", e);
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail("lhvi_peer_alloc", e); }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return LHVI_OK;
}

extern "C" int lhvi_peer_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) { set_error("lhvi_peer_open: bad argument"); return LHVI_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail("lhvi_peer_open: cudaIpcOpenMemHandle", e);
    *ptr = p;
    return LHVI_OK;
}

extern "C" int lhvi_peer_close(void* ptr) {
    if (!ptr) return LHVI_OK;
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? LHVI_OK : cuda_fail("lhvi_peer_close", e);
}

extern "C" int lhvi_peer_free(void* ptr) {
    if (!ptr) return LHVI_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? LHVI_OK : cuda_fail("lhvi_peer_free", e);
}

extern "C" int lhvi_step_tick(double* step, double b1, double b2, void* stream) {
    if (!step) { set_error("lhvi_step_tick: null step buffer"); return LHVI_EINVAL; }
    step_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, b1, b2);
    return check_launch("step_tick_kernel");
}

template <typename real>
static int param_step_t(int K, int64_t n_vars, const uint8_t* kind, const int32_t* dim, const int32_t* off,
                        void* eta, void* tau, void* grad, int64_t n_param, void* m1, void* m2,
                        void* wstate, const double* step, double lr, double b1, double b2, double eps,
                        double var_floor, int sgd, int zero_grad, cudaStream_t s) {
    StepArgs<real> a;
    a.K = K; a.n_vars = n_vars; a.n_param = n_param;
    a.kind = kind; a.dim = dim; a.off = off;
    a.eta = (real*)eta; a.tau = (real*)tau; a.grad = (real*)grad;
    a.m1 = (real*)m1; a.m2 = (real*)m2; a.wstate = (real*)wstate; a.step = step;
    a.lr = (real)lr; a.b1 = (real)b1; a.b2 = (real)b2; a.eps = (real)eps; a.var_floor = (real)var_floor;
    a.sgd = sgd; a.zero_grad = zero_grad;
    param_step_kernel<real><<<grid_for(n_vars, 256), 256, 0, s>>>(a);
    return check_launch("param_step_kernel");
}

extern "C" int lhvi_param_step(int dtype, int K, int64_t n_vars, const uint8_t* var_kind,
                               const int32_t* var_dim, const int32_t* var_off, void* eta, void* tau,
                               void* grad, int64_t n_param, void* mom1, void* mom2, void* wstate,
                               const double* step, double lr, double b1, double b2, double eps,
                               double var_threshold, int sgd, int zero_grad, void* stream) {
    if (!eta || !tau || !grad || !mom1 || !mom2 || !wstate || !step) { set_error("lhvi_param_step: null buffer"); return LHVI_EINVAL; }
    if (n_vars > 0 && (!var_kind || !var_dim || !var_off)) { set_error("lhvi_param_step: null variable table"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    return dtype == LHVI_F64
        ? param_step_t<double>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, zero_grad, s)
        : param_step_t<float>(K, n_vars, var_kind, var_dim, var_off, eta, tau, grad, n_param, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, zero_grad, s);
}

template <typename real>
static int finish_step_t(const lhvi_model* m, long long regions, const lhvi_exchange* x, int K, int64_t n_vars,
                         int64_t n_owned, const uint8_t* kind, const int32_t* dim, const int32_t* off, void* eta,
                         void* tau, void* m1, void* m2, void* wstate, const double* step, double lr, double b1,
                         double b2, double eps, double var_floor, int sgd, cudaStream_t s) {
    FinishArgs<real> f;
    f.partials = m->partials; f.regions = regions; f.K = m->K;
    f.grad = (real*)m->grad; f.n_param = m->n_param;
    f.step = nullptr; f.b1 = b1; f.b2 = b2;
    f.world = 1; f.rank = 0; f.n_idx = 0; f.idx = nullptr; f.seq = nullptr; f.status = nullptr;
    for (int p = 0; p < LHVI_MAX_PEERS; ++p) { f.recv[p] = nullptr; f.flags[p] = nullptr; }
    if (x != nullptr && x->world > 1) {
        f.world = x->world; f.rank = x->rank; f.n_idx = x->n_idx; f.idx = x->idx;
        f.seq = (unsigned long long*)x->seq; f.status = x->status;
        for (int p = 0; p < x->world; ++p) {
            f.recv[p] = (real*)x->recv[p];
            f.flags[p] = (unsigned long long*)x->flags[p];
        }
    }
    StepArgs<real> a;
    a.K = K; a.n_vars = n_vars; a.n_param = m->n_param;
    a.kind = kind; a.dim = dim; a.off = off;
    a.eta = (real*)eta; a.tau = (real*)tau; a.grad = (real*)m->grad;
    a.m1 = (real*)m1; a.m2 = (real*)m2; a.wstate = (real*)wstate; a.step = step;
    a.lr = (real)lr; a.b1 = (real)b1; a.b2 = (real)b2; a.eps = (real)eps; a.var_floor = (real)var_floor;
    a.sgd = sgd; a.zero_grad = 1;
    if (kind == nullptr) {
        // uniform continuous slots (validated by the caller): one thread per 16-byte vector
        constexpr int per = PairVec<real>::pairs;
        const int slot = 2 * K <= 2 ? 2 : (2 * K + 3) / 4 * 4;
        const int slot_vecs = slot / (2 * per);
        const long long n_vec = (long long)n_vars * slot_vecs;
        finish_step_flat_kernel<real><<<1 + grid_for(n_vec > 0 ? n_vec : 1, 256), 256, 0, s>>>(f, a, n_vec, slot_vecs);
        return check_launch("finish_step_flat_kernel");
    }
    finish_step_kernel<real><<<1 + grid_for(n_owned > 0 ? n_owned : 1, 256), 256, 0, s>>>(f, a, (long long)n_owned);
    return check_launch("finish_step_kernel");
}

extern "C" int lhvi_finish_step(const lhvi_model* m, int64_t rows, const lhvi_exchange* x, int64_t n_vars,
                                int64_t n_owned, const uint8_t* var_kind, const int32_t* var_dim,
                                const int32_t* var_off, void* tau, void* mom1, void* mom2, void* wstate,
                                const double* step, double lr, double b1, double b2, double eps,
                                double var_threshold, int sgd, void* stream) {
    if (!m || !m->partials || !m->grad || !m->eta || rows < 0) { set_error("lhvi_finish_step: null buffer or negative rows"); return LHVI_EINVAL; }
    if (!tau || !mom1 || !mom2 || !wstate || !step) { set_error("lhvi_finish_step: null buffer"); return LHVI_EINVAL; }
    const bool uniform = n_vars > 0 && !var_kind && !var_dim && !var_off;      // see lhvi.h: uniform continuous slots
    if (n_vars > 0 && !uniform && (!var_kind || !var_dim || !var_off)) { set_error("lhvi_finish_step: null variable table"); return LHVI_EINVAL; }
    if (n_owned < 0 || n_owned > n_vars) { set_error("lhvi_finish_step: n_owned=%lld outside 0..n_vars=%lld", (long long)n_owned, (long long)n_vars); return LHVI_EINVAL; }
    if (uniform) {
        const int slot = 2 * m->K <= 2 ? 2 : (2 * m->K + 3) / 4 * 4;
        if ((x != nullptr && x->world > 1) || n_owned != n_vars) { set_error("lhvi_finish_step: uniform slots are stepped on one GPU only (pass the variable table)"); return LHVI_EINVAL; }
        if (m->dtype == LHVI_F32 && m->K < 2) { set_error("lhvi_finish_step: uniform float slots need K >= 2 (16-byte vectors)"); return LHVI_EINVAL; }
        if (n_vars * (int64_t)slot > m->n_param) { set_error("lhvi_finish_step: %lld uniform slots of %d elements exceed n_param=%lld", (long long)n_vars, slot, (long long)m->n_param); return LHVI_EINVAL; }
    }
    if (m->K < 1 || m->K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", m->K, LHVI_MAX_K); return LHVI_ELIMIT; }
    if (rows % LHVI_PARTIAL_ROWS != 0) { set_error("lhvi_finish_step: rows must be a multiple of LHVI_PARTIAL_ROWS"); return LHVI_EINVAL; }
    if (x != nullptr && x->world > 1) {
        if (x->world > LHVI_MAX_PEERS) { set_error("lhvi_finish_step: world=%d exceeds LHVI_MAX_PEERS=%d", x->world, LHVI_MAX_PEERS); return LHVI_ELIMIT; }
        if (x->rank < 0 || x->rank >= x->world) { set_error("lhvi_finish_step: rank %d outside world %d", x->rank, x->world); return LHVI_EINVAL; }
        if (x->blocks != 1) { set_error("lhvi_finish_step: the fused launch exchanges with one block (blocks=%d): use lhvi_finish + lhvi_param_step", x->blocks); return LHVI_EINVAL; }
        if (x->n_idx < 0 || (x->n_idx > 0 && !x->idx) || !x->seq || !x->status) { set_error("lhvi_finish_step: incomplete exchange descriptor"); return LHVI_EINVAL; }
        for (int p = 0; p < x->world; ++p)
            if (!x->recv[p] || !x->flags[p]) { set_error("lhvi_finish_step: peer %d has no mapped buffer", p); return LHVI_EINVAL; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long regions = rows / LHVI_PARTIAL_ROWS;
    void* eta = const_cast<void*>(m->eta);
    return m->dtype == LHVI_F64
        ? finish_step_t<double>(m, regions, x, m->K, n_vars, n_owned, var_kind, var_dim, var_off, eta, tau, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s)
        : finish_step_t<float>(m, regions, x, m->K, n_vars, n_owned, var_kind, var_dim, var_off, eta, tau, mom1, mom2, wstate, step, lr, b1, b2, eps, var_threshold, sgd, s);
}

extern "C" int lhvi_mixture_belief(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                                   const uint8_t* q_kind, const void* x, const void* eta, const void* w,
                                   void* out, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!q_off || !q_dim || !q_kind || !x || !eta || !w || !out) { set_error("lhvi_mixture_belief: null buffer"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LHVI_F64)
        mixture_belief_kernel<double><<<grid_for(n, 256), 256, 0, s>>>(K, n, q_off, q_dim, q_kind, (const double*)x, (const double*)eta, (const double*)w, (double*)out);
    else
        mixture_belief_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(K, n, q_off, q_dim, q_kind, (const float*)x, (const float*)eta, (const float*)w, (float*)out);
    return check_launch("mixture_belief_kernel");
}

extern "C" int lhvi_mixture_map(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                                const uint8_t* q_kind, const void* eta, const void* w, void* out, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!q_off || !q_dim || !q_kind || !eta || !w || !out) { set_error("lhvi_mixture_map: null buffer"); return LHVI_EINVAL; }
    if (K < 1 || K > LHVI_MAX_K) { set_error("K=%d out of range 1..%d", K, LHVI_MAX_K); return LHVI_ELIMIT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LHVI_F64)
        mixture_map_kernel<double><<<grid_for(n, 128), 128, 0, s>>>(K, n, q_off, q_dim, q_kind, (const double*)eta, (const double*)w, (double*)out);
    else
        mixture_map_kernel<float><<<grid_for(n, 128), 128, 0, s>>>(K, n, q_off, q_dim, q_kind, (const float*)eta, (const float*)w, (float*)out);
    return check_launch("mixture_map_kernel");
}

// ---- compact host layout <-> padded device slots ------------------------------------------------

namespace lhvi {
template <typename real, bool PACK>
__global__ void __launch_bounds__(256)
state_move_kernel(long long n, const int* __restrict__ map, const real* __restrict__ src, real* __restrict__ dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (PACK) dst[i] = src[map[i]];
        else dst[map[i]] = src[i];
    }
}
}  // namespace lhvi

template <bool PACK>
static int state_move(int dtype, int64_t n, const int32_t* map, const void* src, void* dst, void* stream) {
    if (n == 0) return LHVI_OK;
    if (!map || !src || !dst) { lhvi::set_error("lhvi_state_pack/unpack: null buffer"); return LHVI_EINVAL; }
    if (dtype != LHVI_F32 && dtype != LHVI_F64) { lhvi::set_error("dtype %d is neither LHVI_F32 nor LHVI_F64", dtype); return LHVI_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    long long blocks = (n + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == LHVI_F64)
        lhvi::state_move_kernel<double, PACK><<<(unsigned)blocks, 256, 0, s>>>(n, map, (const double*)src, (double*)dst);
    else
        lhvi::state_move_kernel<float, PACK><<<(unsigned)blocks, 256, 0, s>>>(n, map, (const float*)src, (float*)dst);
    return lhvi::check_launch("state_move_kernel");
}

extern "C" int lhvi_state_pack(int dtype, int64_t n, const int32_t* map, const void* state, void* packed, void* stream) {
    return state_move<true>(dtype, n, map, state, packed, stream);
}

extern "C" int lhvi_state_unpack(int dtype, int64_t n, const int32_t* map, const void* packed, void* state, void* stream) {
    return state_move<false>(dtype, n, map, packed, state, stream);
}
