// The persistent iteration kernel: n Jacobi iterations of the whole model in ONE cooperative launch.
//
// Reference: the body of ADAM_update / GD_update (VarInference.py:249-331) -- all gradients at the
// old parameters, then the step -- repeated `iteration` times by run() (:215-247).
//
// One grid of resident blocks (2 per SM x 148 SMs, launched cooperatively so that they ARE resident)
// loops over the iterations.  Inside an iteration every block works through the record groups
// ("phases") it is assigned to -- all of them, a slice of each, or (lhvi_group::iter_blocks) the one
// group it belongs to: each phase is the body of the per-group kernel (lhvi_spec_impl.cuh /
// lhvi_run_impl.cuh) applied to this block's slice of the group, so the arithmetic -- and the
// results -- are those of the per-group launches.  Then
//
//     grid barrier
//     block 0:  partial rows -> G_w, free energy; (several GPUs) exchange of [G_w | energy | shared
//               gradients] over NVLink peer memory; step of the mixture weights and the shared variables
//     others:   step of the owned variables (Adam / SGD, softmax Jacobians, variance clip, reset of the
//               gradient slots)
//     grid barrier
//
// What this buys over the CUDA graph of per-group launches: no launch levels (a launch costs 3-10 us
// of fixed latency, which is all an iteration has left on an eighth of the model) and small groups
// that are a share of the grid instead of a launch.  Measured schedules: DESIGN.md section 3,
// profiles/r2_iter_plan.md.
//
// The phases are noinline functions (one per instantiated body): the kernel's register count is
// their maximum and ptxas compiles them one by one.  Shared memory is one dynamic allocation viewed
// as the running phase's struct.  Parameters are rewritten between two passes of the same launch,
// so nothing here may read them with ld.global.nc (the bodies use plain loads; the grid barrier's
// acquire fence invalidates L1).
#pragma once
#include <cstring>

#include "lhvi_opt_impl.cuh"
#include "lhvi_spec_impl.cuh"

namespace lhvi {

constexpr int kIterThreads = 256;
constexpr int kIterMaxPhases = 12;
static_assert(kIterThreads == kSpecThreads && kIterThreads == kFoldThreads && kIterThreads == kRunThreads,
              "every body assumes 256 threads per block");

// ---- which body serves a group ---------------------------------------------------------------------

enum { kFamSpec = 1, kFamPun = 2, kFamFold = 3, kFamRun = 4 };

__host__ __device__ constexpr int spec_code(int nc, int ng, int ne, int fl, int w, int hub) {
    return (kFamSpec << 24) | (nc << 20) | (ng << 16) | (ne << 12) | (fl << 8) | (w << 4) | (hub + 1);
}
__host__ __device__ constexpr int pun_code(int ne, int w, int cache) { return (kFamPun << 24) | (ne << 12) | (w << 4) | cache; }
__host__ __device__ constexpr int fold_code(int w, int cache) { return (kFamFold << 24) | (w << 4) | cache; }
__host__ __device__ constexpr int run_code(int ne, int w, int hubpos) { return (kFamRun << 24) | (ne << 12) | (w << 4) | hubpos; }

// X(NC, NG, NE, FL, W, HUB): record-major walk bodies compiled into the iteration kernel
#define LHVI_ITER_SPEC(X)                                                                           \
    X(1, 0, 0, kNode, 1, -1) X(0, 1, 0, kNode, 1, -1)                                               \
    X(0, 0, 1, kPure, 0, -1) X(0, 0, 1, kPure, 1, -1) X(0, 0, 2, kPure, 0, -1) X(0, 0, 2, kPure, 1, -1) \
    X(2, 0, 0, kFull, 0, -1) X(2, 0, 0, kFull, 1, -1) X(2, 0, 0, kFull, 0, 0) X(2, 0, 0, kFull, 1, 0) \
    X(2, 0, 0, kFull, 0, 1) X(2, 0, 0, kFull, 1, 1)                                                 \
    X(2, 0, 1, kFull, 0, -1) X(2, 0, 1, kFull, 1, -1) X(2, 0, 1, kFull, 0, 0) X(2, 0, 1, kFull, 1, 0) \
    X(2, 0, 1, kFull, 0, 1) X(2, 0, 1, kFull, 1, 1)                                                 \
    X(1, 1, 0, kFull, 0, -1) X(1, 1, 0, kFull, 1, -1) X(1, 1, 1, kFull, 0, -1) X(1, 1, 1, kFull, 1, -1)
// a factor whose only integrated argument is a Gaussian-evidence class (C2F rounds).  Single precision
// only: ptxas did not finish the double-precision K = 3 unit with these bodies in it within 16 minutes
// (the same kernels compile in seconds as stand-alone launches), so double-precision C2F models run
// as per-group launches
#define LHVI_ITER_SPEC_F32(X)                                                                       \
    X(0, 1, 0, kFull, 0, -1) X(0, 1, 0, kFull, 1, -1) X(0, 1, 1, kFull, 0, -1) X(0, 1, 1, kFull, 1, -1)
// X(NE, W, CACHE): pure unary records, one variable per record or short runs
#define LHVI_ITER_PUN(X)                                                                            \
    X(0, 0, 0) X(0, 0, 1) X(0, 1, 0) X(0, 1, 1) X(1, 0, 0) X(1, 0, 1) X(1, 1, 0) X(1, 1, 1)         \
    X(2, 0, 0) X(2, 0, 1) X(2, 1, 0) X(2, 1, 1)
// X(W, CACHE): streamed (folded) unary records
#define LHVI_ITER_FOLD(X) X(0, 0) X(0, 1) X(1, 0) X(1, 1)
// X(NE, W, HUBPOS): run-major records
#define LHVI_ITER_RUN(X) X(0, 0, 0) X(0, 0, 1) X(0, 1, 0) X(0, 1, 1) X(1, 0, 0) X(1, 0, 1) X(1, 1, 0) X(1, 1, 1)

template <typename real>
struct IterPhase {
    GroupView<real> view;
    int code;            // spec_code / pun_code / fold_code / run_code
    int nblocks;         // blocks that work on this group
    int rot;             // block `rot` is the group's block 0 (small groups start on different blocks)
    int n_hubs;          // run-major groups
    long long chunk;     // SpecLaunch::chunk for `nblocks` blocks
    double* accum;       // IterArgs::accum (nullptr: rows of `partials`)
};

template <typename real>
struct IterArgs {
    int n_phases, n_iter, tick;
    int* sm_count;                       // [>= 256] per-SM arrival counters (monotonic; parity picks the phase order)
    IterPhase<real> phase[kIterMaxPhases];
    FinishArgs<real> fin;
    StepArgs<real> step;
    double* step_rw;                     // the step counter (advanced here when `tick`)
    long long n_owned;
    unsigned long long* trace;           // lhvi_optim::trace
    double* accum;                       // lhvi_optim::accum: [K + 1] running (G_w, energy) sums of the pass
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The quadrature rule of the run-major body travels in constant memory (see RunLaunch): a noinline
// function cannot take it from the kernel's parameters as constant-bank operands.
// (anonymous namespace: the constant, the setter and the setter's cache are per translation unit --
// with external linkage the linker would merge the setters of the K = 1, 2, 3 units into one that
// fills only ITS unit's constant)
namespace {
template <typename real, int T> struct IterRule;
#define LHVI_ITER_RULE(REAL, NAME)                                              \
    static __constant__ RunLaunch<REAL, 3> NAME;                                     \
    template <> struct IterRule<REAL, 3> {                                      \
        static __device__ __forceinline__ const RunLaunch<REAL, 3>& get() { return NAME; } \
        /* copied only when the values change (the rule is a property of the model, not of a call) */ \
        static cudaError_t set(const RunLaunch<REAL, 3>& v, cudaStream_t s) {   \
            static RunLaunch<REAL, 3> last;                                     \
            static bool have = false;                                           \
            if (have && memcmp(&last, &v, sizeof(v)) == 0) return cudaSuccess;  \
            cudaError_t e = cudaMemcpyToSymbolAsync(NAME, &v, sizeof(v), 0, cudaMemcpyHostToDevice, s); \
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);   /* `v` is a stack object */ \
            if (e == cudaSuccess) { last = v; have = true; }                    \
            return e;                                                           \
        }                                                                       \
    };
LHVI_ITER_RULE(float, c_iter_rule_f32)
LHVI_ITER_RULE(double, c_iter_rule_f64)
#undef LHVI_ITER_RULE
}  // namespace

// ---- phases -------------------------------------------------------------------------------------------

template <typename real, int K, int T, int NC, int NG, int NE, int FL, bool W, int HUB>
__device__ __noinline__ void phase_spec(const IterPhase<real>* ph, int vb, unsigned char* smem) {
    const GroupView<real> g = ph->view;
    SpecLaunch L;
    L.chunk = ph->chunk;
    factor_spec_body<real, K, T, NC, NG, NE, FL, W, HUB>(
        g, L, BlockSlice{vb, ph->nblocks, ph->accum}, *reinterpret_cast<SpecShared<real, K, T, NC, FL, HUB>*>(smem));
}

template <typename real, int K, int T, int NE, bool W, bool CACHE>
__device__ __noinline__ void phase_pun(const IterPhase<real>* ph, int vb, unsigned char* smem) {
    const GroupView<real> g = ph->view;
    SpecLaunch L;
    L.chunk = ph->chunk;
    pure_unary_body<real, K, T, NE, W, CACHE>(g, L, BlockSlice{vb, ph->nblocks, ph->accum},
                                              *reinterpret_cast<PureUnaryShared<real, K, T>*>(smem));
}

template <typename real, int K, bool W, bool CACHE>
__device__ __noinline__ void phase_fold(const IterPhase<real>* ph, int vb, unsigned char* smem) {
    const GroupView<real> g = ph->view;
    SpecLaunch L;
    L.chunk = ph->chunk;
    unary_fold_body<real, K, W, CACHE, true>(g, L, BlockSlice{vb, ph->nblocks, ph->accum},
                                             *reinterpret_cast<FoldBlockShared<real, K, true>*>(smem));
}

template <typename real, int K, int T>
__host__ __device__ constexpr size_t run_shared_bytes() { return (sizeof(RunShared<real, K, T>) + 15) / 16 * 16; }

template <typename real, int K, int T, int NE, bool W, int HUBPOS>
__device__ __noinline__ void phase_run(const IterPhase<real>* ph, int vb, unsigned char* smem) {
    const GroupView<real> g = ph->view;
    factor_run_body<real, K, T, NE, W, HUBPOS>(g, IterRule<real, T>::get(), ph->n_hubs, BlockSlice{vb, ph->nblocks, ph->accum},
                                               *reinterpret_cast<RunShared<real, K, T>*>(smem),
                                               reinterpret_cast<real*>(smem + run_shared_bytes<real, K, T>()));
}

template <typename real, int K, int T>
__device__ __forceinline__ void run_phase(const IterPhase<real>* ph, int vb, unsigned char* smem) {
    switch (ph->code) {
#define X(NC, NG, NE, FL, W, HUB) \
        case spec_code(NC, NG, NE, FL, W, HUB): phase_spec<real, K, T, NC, NG, NE, FL, (W) != 0, HUB>(ph, vb, smem); break;
        LHVI_ITER_SPEC(X)
#undef X
#define X(NC, NG, NE, FL, W, HUB) \
        case spec_code(NC, NG, NE, FL, W, HUB): \
            if constexpr (sizeof(real) == 4) phase_spec<real, K, T, NC, NG, NE, FL, (W) != 0, HUB>(ph, vb, smem); \
            break;
        LHVI_ITER_SPEC_F32(X)
#undef X
#define X(NE, W, CACHE) \
        case pun_code(NE, W, CACHE): phase_pun<real, K, T, NE, (W) != 0, (CACHE) != 0>(ph, vb, smem); break;
        LHVI_ITER_PUN(X)
#undef X
#define X(W, CACHE) \
        case fold_code(W, CACHE): phase_fold<real, K, (W) != 0, (CACHE) != 0>(ph, vb, smem); break;
        LHVI_ITER_FOLD(X)
#undef X
#define X(NE, W, HUBPOS) \
        case run_code(NE, W, HUBPOS): phase_run<real, K, T, NE, (W) != 0, HUBPOS>(ph, vb, smem); break;
        LHVI_ITER_RUN(X)
#undef X
        default: break;
    }
}

// dynamic shared memory a phase needs
template <typename real, int K, int T>
static size_t phase_shared_bytes(int code, int n_hubs) {
    switch (code) {
#define X(NC, NG, NE, FL, W, HUB) case spec_code(NC, NG, NE, FL, W, HUB): return sizeof(SpecShared<real, K, T, NC, FL, HUB>);
        LHVI_ITER_SPEC(X)
#undef X
#define X(NC, NG, NE, FL, W, HUB) case spec_code(NC, NG, NE, FL, W, HUB): \
            return sizeof(real) == 4 ? sizeof(SpecShared<real, K, T, NC, FL, HUB>) : (size_t)-1;
        LHVI_ITER_SPEC_F32(X)
#undef X
#define X(NE, W, CACHE) case pun_code(NE, W, CACHE): return sizeof(PureUnaryShared<real, K, T>);
        LHVI_ITER_PUN(X)
#undef X
#define X(W, CACHE) case fold_code(W, CACHE): return sizeof(FoldBlockShared<real, K, true>);
        LHVI_ITER_FOLD(X)
#undef X
#define X(NE, W, HUBPOS) case run_code(NE, W, HUBPOS): \
            return run_shared_bytes<real, K, T>() + (size_t)n_hubs * 2 * K * kRunThreads * sizeof(real);
        LHVI_ITER_RUN(X)
#undef X
        default: return (size_t)-1;      // no such body
    }
}

// ---- grid barrier -------------------------------------------------------------------------------------
// Arrivals are atomics on `count`, waiting blocks poll `gen` -- two different cache lines, so that the
// last arrival does not queue behind 295 polling loads (cooperative_groups' grid.sync() polls the
// word it adds to: 5.5 us measured for the barrier after the optimiser step, where every block but one
// is already waiting; profiles/r2_iter_plan.md).  Self-resetting; the launch must be cooperative
// (all blocks resident).  Memory ordering: release fence before arriving, acquire fence (which also
// invalidates L1) after leaving -- parameters rewritten before the barrier are re-read after it.
struct GridBarrier {
    unsigned int count;
    unsigned int pad0[31];
    unsigned int gen;
    unsigned int pad1[31];
};

__device__ __forceinline__ void grid_barrier(GridBarrier* bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&bar->gen) : "memory");
        __threadfence();
        if (atomicAdd(&bar->count, 1u) == nblocks - 1) {
            bar->count = 0;
            __threadfence();
            atomicAdd(&bar->gen, 1u);
        } else {
            unsigned int now;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(&bar->gen) : "memory");
            } while (now == gen);
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- the kernel ---------------------------------------------------------------------------------------

template <typename real, int K, int T>
__global__ void __launch_bounds__(kIterThreads, 2)
iterate_kernel(const __grid_constant__ IterArgs<real> A) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ double s_red[(kIterThreads / 32) * (LHVI_MAX_K + 1)];
    __shared__ double s_res[LHVI_MAX_K + 1];
    __shared__ int s_order;
    const int nb = (int)gridDim.x, bid = (int)blockIdx.x;
    GridBarrier* bar = reinterpret_cast<GridBarrier*>(A.sm_count + 256);

    // the two blocks of an SM walk the phases in opposite orders (arrival parity on the SM; a
    // hint: results do not depend on it)
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_order = A.sm_count != nullptr ? (atomicAdd(A.sm_count + (smid & 255u), 1) & 1) : (bid >= nb / 2 ? 1 : 0);
    }
    __syncthreads();
    const bool reversed = s_order != 0;

    for (int it = 0; it < A.n_iter; ++it) {
        unsigned long long* tr = (A.trace != nullptr && threadIdx.x == 0) ? A.trace + ((size_t)it * nb + bid) * 16 : nullptr;
        if (tr) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[0] = global_timer_ns();
            tr[15] = smid;
        }
        if (A.tick && bid == 0 && threadIdx.x == 0) {
            // t, 1 - b1^t, 1 - b2^t (VarInference.py:253,272-273); read after the first grid barrier
            double* st = A.step_rw;
            st[0] = st[0] + 1.0;
            st[1] = 1.0 - (1.0 - st[1]) * A.fin.b1;
            st[2] = 1.0 - (1.0 - st[2]) * A.fin.b2;
        }
        for (int j = 0; j < A.n_phases; ++j) {
            const int p = reversed ? A.n_phases - 1 - j : j;
            const IterPhase<real>* ph = &A.phase[p];
            int vb = bid - ph->rot;
            if (vb < 0) vb += nb;
            if (vb < ph->nblocks) {
                run_phase<real, K, T>(ph, vb, s_dyn);
                __syncthreads();           // the next phase re-uses the shared memory
                if (tr && p < 11) tr[1 + p] = global_timer_ns();
            }
        }
        grid_barrier(bar, (unsigned)nb);
        if (tr) tr[12] = global_timer_ns();

        const real c1 = (real)A.step.step[1], c2 = (real)A.step.step[2];
        if (bid == 0) {
            if (A.accum != nullptr) {
                // every block has added its sums to `accum` before the barrier: publish and clear
                if (threadIdx.x <= K) {
                    A.fin.grad[A.fin.n_param + threadIdx.x] = (real)__ldcg(A.accum + threadIdx.x);
                    A.accum[threadIdx.x] = 0.0;
                }
                __syncthreads();
            } else {
                finish_reduce<real>(A.fin, s_red, s_res);
            }
            if (A.fin.world > 1) finish_exchange<real>(A.fin, 1);
            __syncthreads();
            if (threadIdx.x < 32) step_mixture_weights_warp<real>(A.step, c1, c2);
            for (long long v = A.n_owned + threadIdx.x; v < A.step.n_vars; v += blockDim.x)
                step_variable<real>(A.step, v, c1, c2);
            if (nb == 1)
                for (long long v = threadIdx.x; v < A.n_owned; v += blockDim.x) step_variable<real>(A.step, v, c1, c2);
        } else {
            for (long long v = (bid - 1) * (long long)blockDim.x + threadIdx.x; v < A.n_owned;
                 v += (long long)(nb - 1) * blockDim.x)
                step_variable<real>(A.step, v, c1, c2);
        }
        if (tr) tr[13] = global_timer_ns();
        if (it + 1 < A.n_iter) grid_barrier(bar, (unsigned)nb);        // the end of the launch orders the last step
        if (tr) tr[14] = global_timer_ns();
    }
}

// ---- host side ---------------------------------------------------------------------------------------

// body code of a group, or -1 when the iteration kernel has none for it
static int iter_phase_code(const lhvi_model* m, const lhvi_group* g) {
    if (g->nd != 0) return -1;
    const int w = g->weighted != 0 ? 1 : 0;
    const bool hub0 = ((g->hub_mask >> g->nd) & 1) != 0;
    int code = -1;
    if (g->node) {
        if (g->nc == 1 && g->ng == 0 && g->ne == 0) code = spec_code(1, 0, 0, kNode, 1, -1);
        else if (g->nc == 0 && g->ng == 1 && g->ne == 0) code = spec_code(0, 1, 0, kNode, 1, -1);
    } else if (g->pure) {
        if (g->ng != 0) return -1;
        if (g->nc == 1 && g->fold != nullptr && hub0) code = fold_code(w, 1);
        else if (g->nc == 1 && g->ne <= 2 && !(g->ne > 1 && (g->n % kQuad) != 0)) code = pun_code(g->ne, w, hub0 ? 1 : 0);
        else if (g->nc == 0 && (g->ne == 1 || g->ne == 2)) code = spec_code(0, 0, g->ne, kPure, w, -1);
    } else {
        int hub = -1;
        for (int a = 0; a < g->nc; ++a)
            if ((g->hub_mask >> (g->nd + a)) & 1) { hub = a; break; }
        if (g->nc == 2 && g->ng == 0 && g->ne <= 1) {
            if (g->run_start != nullptr && g->n_hubs >= 1 && g->n_hubs <= kRunMaxHubs) code = run_code(g->ne, w, g->run_hub_arg);
            else code = spec_code(2, 0, g->ne, kFull, w, hub);
        } else if (g->nc <= 1 && g->ng == 1 && g->ne <= 1) {
            code = spec_code(g->nc, 1, g->ne, kFull, w, -1);
        }
    }
    (void)m;
    return code;
}

template <typename real, int K, int T>
int launch_iterate_kt(const lhvi_model* m, const lhvi_group* groups, int n_groups, const lhvi_exchange* x,
                      const lhvi_optim* o, int n_iter, int probe_only, cudaStream_t s) {
    if (m->T != T || !m->rule_symmetric) return 1;
    if (x != nullptr && x->world > 1 && x->blocks != 1) return 1;
    if (probe_only == 0 && o->sm_count == nullptr) { set_error("lhvi_iterate: lhvi_optim::sm_count (512 zero-initialised words of scratch) is required"); return LHVI_EINVAL; }
    IterArgs<real> A;
    A.n_phases = 0;
    size_t dyn = 0;
    for (int i = 0; i < n_groups; ++i) {
        const lhvi_group* g = &groups[i];
        if (g->n == 0) continue;
        if (A.n_phases >= kIterMaxPhases) return 1;
        if (g->n >= (1ll << 31) || g->n_pad >= (1ll << 31)) return 1;
        if (g->pot_kind != LHVI_POT_QUADRATIC) return 1;          // point-by-point potentials: generic kernel only
        const int code = iter_phase_code(m, g);
        if (code < 0) return 1;
        const size_t need = phase_shared_bytes<real, K, T>(code, g->n_hubs);
        if (need == (size_t)-1 || need > 200 * 1024) return 1;
        dyn = need > dyn ? need : dyn;
        IterPhase<real>& ph = A.phase[A.n_phases++];
        ph.view = make_view<real>(m, g, (int64_t)i * LHVI_PARTIAL_ROWS);
        ph.code = code;
        ph.n_hubs = g->n_hubs;
        ph.accum = o->accum;
    }
    // (an empty group's region of `partials` must read "0 valid rows": the buffer starts zeroed,
    // lhvi_factor_expect_grad writes that header for an empty group, and nothing here touches it)
    if (probe_only == 1) return 0;

    auto kernel = iterate_kernel<real, K, T>;
    cudaError_t e = cudaSuccess;
    static size_t set_dyn = 0, occ_dyn = (size_t)-1;       // per instantiation: attribute / occupancy asked once per size
    static int occ_per_sm = 0, occ_sms = 0;
    if (dyn > set_dyn) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) { set_error("iterate_kernel: cudaFuncSetAttribute(%zu bytes): %s", dyn, cudaGetErrorString(e)); cudaGetLastError(); return LHVI_ECUDA; }
        set_dyn = dyn;
    }
    if (dyn != occ_dyn) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&occ_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_per_sm, kernel, kIterThreads, dyn);
        occ_dyn = dyn;
    }
    int per_sm = occ_per_sm;
    const int sms = occ_sms;
    if (per_sm < 1) { set_error("iterate_kernel: no resident block with %zu bytes of shared memory", dyn); return LHVI_ELIMIT; }
    if (per_sm > 2) per_sm = 2;
    const int resident = per_sm * (sms > 0 ? sms : 148);
    if (probe_only == 2) return resident;

    // Schedule.  Sliced (no group carries iter_blocks): every group is cut into up to `resident`
    // slices of whole tiles and every block walks the groups one after the other; the grid is as
    // large as the largest group (and the optimiser step) can use.  Partitioned (iter_blocks > 0
    // everywhere): group p gets iter_blocks[p] blocks of its own, [rot, rot + nblocks), so that the
    // groups run side by side as they do on parallel branches of a CUDA graph.
    bool partitioned = A.n_phases > 0;
    for (int i = 0; i < n_groups; ++i)
        if (groups[i].n > 0 && groups[i].iter_blocks <= 0) partitioned = false;
    int grid = 2;
    int rot = 0;
    {
        int p = 0;
        for (int i = 0; i < n_groups; ++i) {
            if (groups[i].n == 0) continue;
            IterPhase<real>& ph = A.phase[p++];
            const long long n = ph.view.n;
            const int fam = ph.code >> 24;
            const long long cap = partitioned ? (groups[i].iter_blocks < resident ? groups[i].iter_blocks : resident) : resident;
            long long blocks, chunk;
            if (fam == kFamFold) {
                const long long tiles = ph.view.n_pad / kFoldTile;
                const long long c = (tiles + cap - 1) / cap;
                blocks = (tiles + c - 1) / c;
                chunk = tiles;                              // the body splits `chunk` tiles over `nblocks` blocks
            } else if (fam == kFamRun) {
                blocks = (ph.view.n_runs + kRunThreads - 1) / kRunThreads;
                if (blocks > cap) blocks = cap;
                chunk = 0;
            } else {
                const long long tile = fam == kFamPun ? (long long)kSpecThreads * kQuad : (long long)kSpecThreads;
                blocks = (n + tile - 1) / tile;
                if (blocks > cap) blocks = cap;
                chunk = ((n + blocks - 1) / blocks + tile - 1) / tile * tile;
                blocks = (n + chunk - 1) / chunk;
            }
            if (blocks < 1) blocks = 1;
            if (blocks > LHVI_PARTIAL_ROWS - 1) return 1;
            ph.nblocks = (int)blocks;
            ph.chunk = chunk;
            if (partitioned) {
                ph.rot = rot;
                rot += ph.nblocks;
            } else if (ph.nblocks > grid) {
                grid = ph.nblocks;
            }
        }
    }
    if (partitioned) {
        if (rot > resident) { set_error("lhvi_iterate: iter_blocks add up to %d, only %d blocks are resident", rot, resident); return LHVI_EINVAL; }
        grid = rot > 2 ? rot : 2;
    } else {
        const long long vb = (o->n_owned + kIterThreads - 1) / kIterThreads + 1;
        if (vb > grid) grid = (int)(vb < resident ? vb : resident);
        for (int p = 0; p < A.n_phases; ++p) {           // small groups start on different blocks
            IterPhase<real>& ph = A.phase[p];
            ph.rot = 0;
            if (ph.nblocks < grid) {
                ph.rot = rot % grid;
                rot += ph.nblocks;
            }
        }
    }

    // run-major rule constants
    {
        RunLaunch<real, T> L;
        memset(&L, 0, sizeof(L));
        double M0 = 0.0, M2 = 0.0, M4 = 0.0, xm = 0.0, eqm = 1.0;
        for (int t = 0; t < T; ++t) {
            const double xq = m->quad_host[t], om = m->quad_host[T + t];
            L.xi[t] = (real)xq; L.w0[t] = (real)om; L.w1[t] = (real)(om * xq); L.w2[t] = (real)(om * xq * xq);
            L.eq[t] = (real)::exp(-xq * xq);
            eqm = ::exp(-xq * xq) < eqm ? ::exp(-xq * xq) : eqm;
            M0 += om; M2 += om * xq * xq; M4 += om * xq * xq * xq * xq;
            xm = ::fabs(xq) > xm ? ::fabs(xq) : xm;
        }
        if (!(M0 > 0.5)) { set_error("lhvi_iterate: lhvi_model::quad_host is not filled in"); return LHVI_EINVAL; }
        L.eq_min = (real)eqm;
        L.cm0 = (real)(M0 * M0); L.cm2 = (real)(M0 * M2); L.cm22 = (real)(M2 * M2);
        L.cmd = (real)(M0 * M4 - M2 * M2); L.xm = (real)xm;
        e = IterRule<real, T>::set(L, s);
        if (e != cudaSuccess) { set_error("lhvi_iterate: cudaMemcpyToSymbolAsync: %s", cudaGetErrorString(e)); cudaGetLastError(); return LHVI_ECUDA; }
    }

    A.n_iter = n_iter;
    A.tick = o->sgd ? 0 : 1;
    A.sm_count = o->sm_count;
    A.step_rw = o->step;
    A.n_owned = o->n_owned;
    A.trace = (unsigned long long*)o->trace;
    A.accum = o->accum;
    FinishArgs<real>& f = A.fin;
    f.partials = m->partials; f.regions = n_groups; f.K = m->K;
    f.grad = (real*)m->grad; f.n_param = m->n_param;
    f.step = nullptr; f.b1 = o->b1; f.b2 = o->b2;
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
    StepArgs<real>& a = A.step;
    a.K = m->K; a.n_vars = o->n_vars; a.n_param = m->n_param;
    a.kind = o->var_kind; a.dim = o->var_dim; a.off = o->var_off;
    a.eta = (real*)const_cast<void*>(m->eta); a.tau = (real*)o->tau; a.grad = (real*)m->grad;
    a.m1 = (real*)o->mom1; a.m2 = (real*)o->mom2; a.wstate = (real*)o->wstate; a.step = o->step;
    a.lr = (real)o->lr; a.b1 = (real)o->b1; a.b2 = (real)o->b2; a.eps = (real)o->eps; a.var_floor = (real)o->var_threshold;
    a.sgd = o->sgd; a.zero_grad = 1;

    void* params[] = {(void*)&A};
    e = cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3(kIterThreads), params, dyn, s);
    if (e != cudaSuccess) { set_error("iterate_kernel: cooperative launch (%d blocks, %zu bytes): %s", grid, dyn, cudaGetErrorString(e)); cudaGetLastError(); return LHVI_ECUDA; }
    return LHVI_OK;
}

}  // namespace lhvi
