/*
 * lhvi.h -- C ABI of the B200 variational-inference hot path.
 *
 * This is the boundary a host binds with ctypes / cffi / cgo / JNI.  Every entry point
 *   - takes raw DEVICE pointers and plain sizes (no framework types),
 *   - is asynchronous on the CUDA stream passed as `stream` (a cudaStream_t cast to void*,
 *     NULL = legacy default stream),
 *   - never allocates, frees or synchronises: the caller owns every buffer (the only exception is
 *     the explicit lhvi_peer_* life cycle of the cross-GPU exchange buffers),
 *   - returns 0 on success or a negative LHVI_E* code; lhvi_last_error() then holds a
 *     human-readable message (thread-local).
 *
 * What each call replaces in the reference (leodd/Lifted-Hybrid-Variational-Inference):
 *
 *   lhvi_factor_expect_grad   the per-factor loops of gradient_w_tau / gradient_mu_var /
 *                             gradient_category_tau / free_energy with their calls into
 *                             expectation() and rvs_belief()
 *                             (VarInference.py:40-55,57-195,336-353; lifted weights
 *                             LiftedVarInference.py:74,90,131-132,162; Gaussian evidence
 *                             C2FVarInference.py:110-113,266-267).  The variables' own
 *                             (N-1) E[log b] terms (VarInference.py:60-72,96-106,138-139)
 *                             are the same kernel run over "node" record groups.
 *   lhvi_elbo_reduce          the running sums `energy -= ...` / `g_w[k] -= ...`
 *                             (VarInference.py:72,88,177,193).
 *   lhvi_step_tick            `self.t += 1` and the bias corrections 1-b1^t, 1-b2^t
 *                             (VarInference.py:253,272-273).
 *   lhvi_param_step           softmax Jacobians (VarInference.py:90,160), the Adam moment and
 *                             parameter update with the variance clip and re-normalisation
 *                             (VarInference.py:256-287), or the plain SGD step (:302-329).
 *   lhvi_mixture_belief       belief(x, rv) for a batch of (variable, x) queries
 *                             (VarInference.py:333-353).
 *   lhvi_mixture_map          map(rv) for a batch of variables (VarInference.py:355-376).
 *   lhvi_finish               lhvi_elbo_reduce + lhvi_step_tick in one launch, and -- when the
 *                             records are sharded over several GPUs -- the sum over ranks of
 *                             G_w, the free energy and the gradients of the variables that
 *                             more than one rank touches, exchanged through NVLink peer
 *                             memory inside the same kernel.  The reference is single-process:
 *                             these are still its `energy -= ...` / `g_w[k] -= ...` /
 *                             `g_mu -= ...` sums (VarInference.py:72,88,120-129), split by rank.
 *   lhvi_finish_step          lhvi_finish + lhvi_param_step fused into one launch.
 *   lhvi_iterate              n whole iterations -- every group's factor pass, lhvi_finish and
 *                             lhvi_param_step -- in ONE persistent cooperative launch: the loop
 *                             `for itr in range(iteration)` of ADAM_update / GD_update
 *                             (VarInference.py:249-331) with the grid barrier in the place of the
 *                             launch boundaries.
 *   lhvi_peer_*               life cycle of the peer-visible exchange buffers (CUDA IPC).
 *   lhvi_state_pack/unpack    the compact per-variable parameter arrays of the reference
 *                             (eta[rv], VarInference.py:197-213) <-> the padded device slots.
 *
 * Data layout: a flat parameter vector with one slot per hidden variable
 * (continuous: K x (mu, var) interleaved; discrete: K x D row-major probabilities), a
 * coefficient table `ptab`, and per-signature record groups stored column-major.
 */
#ifndef LHVI_H
#define LHVI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LHVI_ABI_VERSION 8

/* element type of every `void*` buffer of reals */
#define LHVI_F32 0
#define LHVI_F64 1

/* error codes */
#define LHVI_OK 0
#define LHVI_EINVAL (-1)   /* bad argument (message says which) */
#define LHVI_ELIMIT (-2)   /* model exceeds a compiled-in limit */
#define LHVI_ECUDA (-3)    /* CUDA runtime error at launch */

/* compiled-in limits of the generic kernel (the specialised kernels are stricter) */
#define LHVI_MAX_AXES 6      /* hidden + Gaussian-evidence arguments of one factor */
#define LHVI_MAX_K 8
#define LHVI_MAX_T 32
#define LHVI_MAX_DSTATES 16
#define LHVI_MAX_NODES 64    /* sum over axes of quadrature nodes / states */
#define LHVI_MAX_GACC 256    /* sum over hidden arguments of K * (2 | D) */
#define LHVI_PARTIAL_ROWS 1184   /* rows of `partials` reserved per launch: 1 header + 1183 data */
#define LHVI_RUN_MAX_HUBS 16 /* distinct hub variables of a run-major group (lhvi_group::run_*) */
#define LHVI_MAX_PEERS 16    /* GPUs of one NVLink domain taking part in lhvi_finish's exchange */
#define LHVI_IPC_HANDLE_BYTES 64

/* lhvi_group::pot_kind: how a coefficient block of ptab turns into log psi at a grid point */
#define LHVI_POT_QUADRATIC 0   /* log psi = the block's quadratic (every exp-quadratic / table potential) */
#define LHVI_POT_HARD 1        /* psi = 1 where the block's quadratic is > 0, else 0
                                  (MLNHardPotential over continuous arguments, MLNPotential.py:43-49) */
#define LHVI_POT_IMAGE_EDGE 2  /* block = [distant_cof, scaling_cof, max_threshold, exp(-thr / scaling), 0, 0];
                                  d = |x0 - x1|, psi = d distant_cof + (d > thr ? block[3] : exp(-d / scaling_cof))
                                  (ImageEdgePotential, Potential.py:411-424); two continuous arguments */

/*
 * One record group: `n` factor records sharing a canonical signature.  Arguments are
 * ordered [hidden discrete | hidden continuous | Gaussian evidence | point evidence].
 * All arrays are device pointers, column-major over records (column a starts at a*n).
 */
typedef struct lhvi_group {
    int32_t nd, nc, ng, ne;        /* argument counts per role */
    int32_t dims[LHVI_MAX_AXES];   /* states of each hidden discrete argument */
    int32_t node;                  /* 1: node-entropy records: F = log b (no ptab); energy and G_w are
                                      scaled by wf, parameter gradients by nscale */
    int32_t weighted;              /* 1: wf / gam present; 0: all weights are 1 */
    int32_t hub_mask;              /* bit a set: hidden argument a mostly comes in long runs of the
                                      same variable (a hub); its gradient is accumulated in shared
                                      memory per block before touching global memory.  A hint:
                                      results do not depend on it. */
    int32_t pure;                  /* 1: F = log psi only (unary split: the -log b part of these
                                      records is carried by the node records); no belief evaluated */
    int64_t n;                     /* records in this group */
    const int32_t* pot;            /* [n]          offset of the coefficient block in ptab */
    const int32_t* poff;           /* [(nd+nc)*n]  parameter slot offsets */
    const void* egval;             /* [ng*n]       Gaussian-evidence means */
    const void* egvar;             /* [ng*n]       Gaussian-evidence variances */
    const void* ecval;             /* [ne*n]       point-evidence values */
    const void* wf;                /* [n]          W_f, weight on energy and g_w (node: energy scale) */
    const void* gam;               /* [(nd+nc)*n]  gamma, weight on each parameter gradient */
    const void* nscale;            /* [n]          node groups: parameter-gradient scale */
    /* Optional streaming form of a pure group with exactly one hidden continuous argument
       (nd=0, nc=1, ng=0, pure=1): the record's log-potential as a quadratic in that argument
       with the point evidence already folded in, log psi(x) = c0 + l0 x + a0 x^2.
       fold = [3][n_pad] (c0 | l0 | a0), n_pad a multiple of 1024 >= n; poff / wf / gam must
       then be allocated with n_pad elements as well, records n..n_pad-1 carrying zero
       coefficients and weights and the last record's offset.  NULL: not provided. */
    const void* fold;
    int64_t n_pad;
    /* Optional run-major form of a full group with two hidden continuous arguments (nd=0, nc=2,
       ng=0, node=0, pure=0) of which argument `run_hub_arg` (0 or 1) takes at most
       LHVI_RUN_MAX_HUBS distinct values: the records (every column above) are sorted by the
       *other* hidden argument, run i = records run_start[i] .. run_start[i+1]-1 all have that
       argument at parameter offset run_key[i]; run_hid[r] indexes hub_keys (the distinct offsets
       of the hub argument).  A long run may be split into several runs with the same key.
       run_start == NULL: not provided. */
    const int32_t* run_start;      /* [n_runs + 1] */
    const int32_t* run_key;        /* [n_runs] */
    const int32_t* run_hid;        /* [n] */
    const int32_t* hub_keys;       /* [n_hubs] */
    int64_t n_runs;
    int32_t n_hubs;
    int32_t run_hub_arg;
    /* lhvi_iterate only: how many blocks of the persistent grid work on this group.  0: every block
       takes a slice of the group (the blocks walk the groups one after the other); > 0: the group
       gets that many blocks of its own, disjoint from the other groups' (the groups then run side by
       side, each block with one group's prologue and epilogue per iteration).  Either all groups
       carry a count or none does.  A hint for the schedule: results do not depend on it. */
    int32_t iter_blocks;
    /* != 0: lhvi_factor_expect_grad leaves the categorical gradients G_c of this group's hidden discrete
       arguments out (energy, G_w and the (mu, var) gradients are computed as always): they come from
       lhvi_category_grad_reference instead (the unmodified reference's behaviour, see there) */
    int32_t no_category_grad;
    /* LHVI_POT_*: != LHVI_POT_QUADRATIC groups are evaluated point by point in the generic kernel (no
       specialised kernel, no fold / run-major columns, not part of lhvi_iterate) */
    int32_t pot_kind;
    int32_t reserved0;
    /* Optional, run-major groups only: records of the run variable itself, evaluated once per run inside
       the run-major kernel, where the variable's parameters and axis tables are in registers anyway
       (they would otherwise be separate record groups and launches).  A run that was cut into several
       runs with the same key carries them on ONE of its pieces (zeros / -1 on the others).
         run_node    [2][n_runs] reals: W (energy / G_w scale) and the parameter-gradient scale of the
                     variable's node-entropy record (F = log b on its own quadrature nodes); 0, 0: none
         run_una_pot [n_runs] offset in ptab of the coefficient block (c, l, a) of ONE pure unary record
                     (nd=0, nc=1, ng=0, ne=0, LHVI_POT_QUADRATIC: F = log psi only) on the variable; -1: none
         run_una_w   [2][n_runs] reals: W_f and gamma of that record
       NULL: not provided.  Results equal those of the separate records. */
    const void* run_node;
    const int32_t* run_una_pot;
    const void* run_una_w;
    /* Optional, streamed groups only (fold != NULL): cst_n records without any integrated argument (nd=nc=ng=0,
       LHVI_POT_QUADRATIC: every argument observed) -- constants of the free energy and of G_w -- that the
       streaming kernel evaluates beside its own records (one per thread and tile, so their loads hide behind
       the tile's) instead of a launch of their own.  cst_q [cst_n]: the record's quadratic with its point
       evidence substituted, as `fold` does for the streamed records (the kernel still applies
       log(exp(q) + 1e-100) and the weight to every record in every pass); cst_wf [cst_n]: W_f (NULL: all 1). */
    const void* cst_q;
    const void* cst_wf;
    int64_t cst_n;
} lhvi_group;

/* Model-wide device buffers shared by every group launch. */
typedef struct lhvi_model {
    int32_t dtype;                 /* LHVI_F32 | LHVI_F64 */
    int32_t K, T;                  /* mixture components, quadrature points */
    int32_t rule_symmetric;        /* 1: quad is exactly mirror-symmetric (x_t = -x_{T-1-t}, equal weights,
                                      x = 0 in the middle of an odd rule), as numpy's hermgauss returns it;
                                      the specialised kernels require it (0: generic kernel) */
    int64_t n_param;               /* elements of eta / grad parameter part */
    const void* quad;              /* [2T]  Gauss-Hermite nodes, then weights / sqrt(pi) */
    const void* ptab;              /* coefficient table */
    const void* eta;               /* [n_param]  (mu,var) | probabilities */
    const void* w;                 /* [K]  mixture weights */
    void* grad;                    /* [n_param + K + 1]  parameter grads | G_w | energy */
    double* partials;              /* [rows][K+1] per-block partial sums of G_w | energy */
    double quad_host[2 * LHVI_MAX_T]; /* host copy of quad (nodes, then weights / sqrt(pi)): the run-major
                                      kernel takes the rule through its launch parameters (constant bank)
                                      instead of registers; required when a group carries run_* columns */
} lhvi_model;

const char* lhvi_last_error(void);
int lhvi_abi_version(void);

/* 1 if a template-specialised kernel exists for this (model, group), else 0 (generic). */
int lhvi_has_specialisation(const lhvi_model* m, const lhvi_group* g);

/*
 * Accumulate one group's contribution: atomically adds parameter gradients into m->grad
 * and writes per-block partial sums of (G_w[0..K-1], energy) to the region of
 * LHVI_PARTIAL_ROWS rows of m->partials starting at row0 (a multiple of LHVI_PARTIAL_ROWS):
 * row0 is a header whose first element is the number of valid data rows that follow.
 * force_generic != 0 bypasses the specialised kernels (used by the parity tests).
 */
int lhvi_factor_expect_grad(const lhvi_model* m, const lhvi_group* g, int64_t row0,
                            int force_generic, void* stream);

/* Sum the valid rows of the first rows / LHVI_PARTIAL_ROWS regions of partials into
 * grad[n_param .. n_param+K] (G_w, energy).  Every region must have been written by a
 * lhvi_factor_expect_grad call since the buffer was allocated. */
int lhvi_elbo_reduce(const lhvi_model* m, int64_t rows, void* stream);

/* step[0] = t, step[1] = 1-b1^t, step[2] = 1-b2^t (doubles, device).  Increments t. */
int lhvi_step_tick(double* step, double b1, double b2, void* stream);

/*
 * Exchange descriptor of lhvi_finish (one per rank).  The exchanged vector is
 * x = [G_w[K] | energy | grad[idx[0]] .. grad[idx[n_idx-1]]].  recv[p] / flags[p] are rank p's
 * buffers as mapped into THIS process (entry `rank` is the local buffer itself):
 *   recv  : [2][world][K+1+n_idx] reals   (double-buffered by the parity of the sequence number)
 *   flags : [world][blocks] uint64        (sequence number of the last completed send)
 * Every rank writes its x into slot `rank` of every peer's recv, publishes the sequence number
 * with a system-scope release, waits for all `world` flags of its own buffer, and adds the
 * slots in rank order -- so all ranks obtain bit-identical sums.  A wait that exceeds ~2 s sets
 * *status = 1 and gives up instead of hanging the device.
 */
typedef struct lhvi_exchange {
    int32_t world, rank;
    int32_t blocks;                       /* thread blocks of the launch (1..32); fixed per buffer */
    int32_t reserved;
    int64_t n_idx;                        /* gathered gradient elements */
    const int32_t* idx;                   /* [n_idx] offsets into grad (device) */
    void* recv[LHVI_MAX_PEERS];
    uint64_t* flags[LHVI_MAX_PEERS];
    uint64_t* seq;                        /* [blocks] local sequence counters (device, zero-initialised) */
    int32_t* status;                      /* [1] local, device */
} lhvi_exchange;

/*
 * One launch after the factor kernels of an iteration: sums the partial rows like
 * lhvi_elbo_reduce; if step != NULL also advances the step counter like lhvi_step_tick; if
 * x != NULL and x->world > 1 exchanges and sums x over the ranks (every rank must make the
 * matching call).  grad[n_param..] and grad[idx[*]] hold the all-rank sums afterwards.
 */
int lhvi_finish(const lhvi_model* m, int64_t rows, double* step, double b1, double b2,
                const lhvi_exchange* x, void* stream);

/*
 * Peer-visible device memory for lhvi_exchange (the only memory this library allocates).
 *   lhvi_peer_alloc  cudaMalloc + zero-fill + cudaIpcGetMemHandle -> *ptr, handle[64]
 *   lhvi_peer_open   map another process's handle (enables peer access lazily) -> *ptr
 *   lhvi_peer_close  unmap a pointer returned by lhvi_peer_open
 *   lhvi_peer_free   free a pointer returned by lhvi_peer_alloc
 * These synchronise the device; call them at set-up / tear-down only.
 */
int lhvi_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle);
int lhvi_peer_open(const unsigned char* handle, void** ptr);
int lhvi_peer_close(void* ptr);
int lhvi_peer_free(void* ptr);

/*
 * Parameter update for every hidden variable and for w_tau.
 *   var_kind[v] 0 continuous / 1 discrete, var_dim[v] 2 / D, var_off[v] slot offset.
 *   eta: read by the factor kernels; tau: logits of discrete slots; mom1 / mom2: Adam moments.
 *   wstate: [5K] reals: w_tau | w | mom1_w | mom2_w | scratch.
 *   sgd != 0 -> theta -= lr * g (moments untouched); else Adam with eps outside the sqrt.
 *   The softmax Jacobian is applied here to raw G_c and G_w.
 *   zero_grad != 0 -> every gradient slot of the listed variables is reset to 0 after it has been
 *   consumed, so the next iteration's factor kernels accumulate into a clean buffer without a
 *   separate memset (the variables listed must cover every slot the factor kernels touch).
 */
int lhvi_param_step(int dtype, int K, int64_t n_vars, const uint8_t* var_kind,
                    const int32_t* var_dim, const int32_t* var_off, void* eta, void* tau,
                    void* grad, int64_t n_param, void* mom1, void* mom2, void* wstate,
                    const double* step, double lr, double b1, double b2, double eps,
                    double var_threshold, int sgd, int zero_grad, void* stream);

/*
 * out[i] = sum_k w_k q_{v_i,k}(x_i) for n queries; continuous variables use the
 * reference's density (normaliser 1/(sqrt(2 pi) var)); for discrete ones x_i is the state index.
 */
int lhvi_mixture_belief(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                        const uint8_t* q_kind, const void* x, const void* eta, const void* w,
                        void* out, void* stream);

/*
 * MAP of n variables' marginal beliefs (VarInference.py:355-376): for a continuous variable the
 * arg-max of sum_k w_k q_k(x) reached from the best component mean by safeguarded Newton ascent
 * (the reference uses scipy BFGS from the same start); for a discrete one the arg-max state index.
 */
int lhvi_mixture_map(int dtype, int K, int64_t n, const int32_t* q_off, const int32_t* q_dim,
                     const uint8_t* q_kind, const void* eta, const void* w, void* out, void* stream);

/*
 * lhvi_finish and lhvi_param_step in ONE launch (the iteration's critical path loses a launch
 * level).  Block 0 reduces the partial rows, exchanges (x->blocks must be 1) and then steps the
 * mixture weights and the variables n_owned .. n_vars-1 of the table -- the ones whose gradients
 * are completed by the exchange; all other blocks step the variables 0 .. n_owned-1 right away.
 * With one GPU n_owned = n_vars.  The step counter is NOT advanced here: call lhvi_step_tick
 * earlier in the iteration (it can run beside the factor kernels).  eta and grad are taken from
 * the model; gradient slots are always reset (zero_grad = 1).
 *
 * Uniform continuous slots: with var_kind = var_dim = var_off = NULL (one GPU, n_owned = n_vars) every one of
 * the n_vars variables is continuous and variable v sits at element v * slot, slot = 2 for K = 1, else 2K
 * rounded up to a multiple of 4 (float: K >= 2).  The launch then steps one 16-byte vector per thread,
 * fully coalesced, without the variable table.
 */
int lhvi_finish_step(const lhvi_model* m, int64_t rows, const lhvi_exchange* x, int64_t n_vars,
                     int64_t n_owned, const uint8_t* var_kind, const int32_t* var_dim,
                     const int32_t* var_off, void* tau, void* mom1, void* mom2, void* wstate,
                     const double* step, double lr, double b1, double b2, double eps,
                     double var_threshold, int sgd, void* stream);

/*
 * compat="reference": the categorical gradients as the UNMODIFIED reference computes them.
 * gradient_category_tau (VarInference.py:133-160; LiftedVarInference.py:136-164) builds the quadrature
 * axes of the OTHER hidden arguments of a factor from the domain of the discrete variable being
 * differentiated (`rv.domain`, :147-150) instead of their own: another hidden argument b is
 * "enumerated" over the values of argument a's domain, with row k of b's parameter table as weights
 * (a categorical row, or (mu, var) for a continuous b).  Everything else of the reference is unaffected,
 * so a caller who wants its numbers bit for bit sets lhvi_group::no_category_grad on the full groups
 * with a hidden discrete argument and adds this call per such group.  Supported when every other hidden
 * argument's table row is as long as a's domain (booleans next to reals, equal-cardinality discrete
 * arguments): with unequal lengths the reference's zip(product, product) pairs values and weights in
 * an order that depends on the factor's argument order, which the record groups do not keep.
 *   dvals[a][j]   value j of hidden discrete argument a's domain
 *   xmap[a][b][j] state index of hidden discrete argument b whose value equals dvals[a][j]
 * Adds  -gamma * sum_grid prod(weights) * (log(psi + 1e-100) - log(b + 1e-100))  into grad.
 */
typedef struct lhvi_h2 {
    double dvals[LHVI_MAX_AXES][LHVI_MAX_DSTATES];
    int32_t xmap[LHVI_MAX_AXES][LHVI_MAX_AXES][LHVI_MAX_DSTATES];
} lhvi_h2;

int lhvi_category_grad_reference(const lhvi_model* m, const lhvi_group* g, const lhvi_h2* h, void* stream);

/*
 * Optimiser-side buffers of lhvi_iterate: what lhvi_finish_step takes as loose arguments.
 */
typedef struct lhvi_optim {
    int64_t n_vars, n_owned;       /* variable table; entries n_owned .. n_vars-1 are the shared ones */
    const uint8_t* var_kind;       /* [n_vars] 0 continuous / 1 discrete */
    const int32_t* var_dim;        /* [n_vars] 2 / D */
    const int32_t* var_off;        /* [n_vars] slot offsets */
    void* tau;                     /* [n_param] logits of the discrete slots */
    void* mom1;                    /* [n_param] Adam first moments */
    void* mom2;                    /* [n_param] Adam second moments */
    void* wstate;                  /* [5K] w_tau | w | mom1_w | mom2_w | scratch */
    double* step;                  /* [4] t, 1-b1^t, 1-b2^t, unused (advanced once per iteration unless sgd) */
    int32_t* sm_count;             /* [512] zero-initialised scratch words owned by the caller and used by no other
                                      stream at the same time: per-SM arrival counters [0, 256), the grid barrier
                                      [256, 384), the rest reserved */
    double lr, b1, b2, eps, var_threshold;
    int32_t sgd;                   /* != 0: theta -= lr * g, moments and step counter untouched */
    int32_t reserved;
    double* accum;                 /* NULL, or [K + 1] zero-initialised doubles owned by the caller: the blocks add
                                      their (G_w, energy) sums here with atomics and block 0 publishes and clears
                                      them after the barrier (NULL: rows of m->partials, summed by block 0) */
    uint64_t* trace;               /* NULL, or [n_iter][blocks][16] device words: %globaltimer (ns) of every block
                                      at the start of an iteration (0), after group i of the table (1 + i, i < 11),
                                      after the first grid barrier (12), after its share of the step (13) and
                                      after the second barrier (14); 15 = the SM it runs on.  For tools/iter_trace.py. */
} lhvi_optim;

/*
 * n_iter Jacobi iterations in ONE cooperative launch of a persistent grid (2 blocks per SM): per
 * iteration every block works through its slice of every record group (the same device code as
 * lhvi_factor_expect_grad's kernels, region i of m->partials belonging to groups[i]), a grid
 * barrier, then block 0 does lhvi_finish's work (x != NULL and x->world > 1: with the exchange,
 * x->blocks must be 1; every rank must make the matching call with the same n_iter) and steps the
 * mixture weights and the shared variables while the other blocks step the owned ones, and a
 * second grid barrier.  grad must be clean (all parameter-gradient slots zero) on entry and is
 * left clean; grad[n_param ..] holds G_w and the free energy of the LAST pass (at the parameters
 * before its step).  Equivalent to n_iter times { lhvi_factor_expect_grad for every group;
 * lhvi_step_tick; lhvi_finish_step }.
 * Returns 0, a negative error code, or 1 when some group has no body in the iteration kernel
 * (hidden discrete arguments, T != 3, K > 3, more than 12 groups, an exchange that needs more
 * than one block): nothing was launched, use the per-group calls.
 */
int lhvi_iterate(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups, const lhvi_exchange* x,
                 const lhvi_optim* opt, int32_t n_iter, void* stream);

/* 1 if lhvi_iterate would take this model (only the descriptors' shapes are looked at), else 0. */
int lhvi_iterate_supported(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups,
                           const lhvi_exchange* x);

/* Blocks of lhvi_iterate's persistent grid that are resident on the current device for this model
 * (2 per SM unless a group's shared memory allows only 1): what lhvi_group::iter_blocks may add up
 * to.  0: lhvi_iterate would not take the model; negative: error.  Queries the device, launches nothing. */
int lhvi_iterate_blocks(const lhvi_model* m, const lhvi_group* groups, int32_t n_groups,
                        const lhvi_exchange* x);

/*
 * The state a reference caller holds between ADAM_update calls is eta[rv] as K x 2 / K x D arrays
 * (VarInference.py:197-213), i.e. without the padding of the device slots.  These two calls move
 * between that compact layout and the slot layout: packed[i] = state[map[i]] (pack) and
 * state[map[i]] = packed[i] (unpack) for i < n; `map` [n] holds the element offsets of the used
 * slot elements in ascending order (device).  A host transfers `packed` instead of the padded
 * vector (25 % fewer bytes at K = 3).
 */
int lhvi_state_pack(int dtype, int64_t n, const int32_t* map, const void* state, void* packed, void* stream);
int lhvi_state_unpack(int dtype, int64_t n, const int32_t* map, const void* packed, void* state, void* stream);

/*
 * Gaussian belief propagation in information form: the device replacement of the reference's GaBP
 * (GaBP.py:7-216; run :139-165, message_rv_to_f :19-35, message_f_to_rv :37-136, get_belief_params
 * :183-195), SURVEY section 8 f-4 (the config-4 cross-check at full size).
 *
 * n_vars hidden variables; every pairwise factor f over hidden variables (i, j) with
 * log psi_f = -1/2 x' Jf x + hf' x contributes two directed message slots e = (f: i -> j) and
 * rev[e] = (f: j -> i) with coef = [Jf_ss | Jf_dd | Jf_sd | hf_s | hf_d] (s = src[e], d = dst[e]),
 * column-major [5][n_edges].  jd / hd [n_vars] are the constant messages of the unary and
 * evidence-reduced factors.  P, H [2][n_edges] (message precision, potential) and SP, SH [2][n_vars]
 * (per-variable totals of the incoming messages) are double-buffered; `parity` says which half is
 * current.  One sweep is the reference's pair of loops (all variable -> factor messages, then all
 * factor -> variable messages, both from the previous sweep's values):
 *     a = SP[s] - P[rev], b = SH[s] - H[rev];  P' = J_dd - J_sd^2 / (J_ss + a);  H' = h_d - J_sd (h_s + b) / (J_ss + a)
 * The reference starts every message at (mu, sig) = (0, 1): P = 1, H = 0, SP = number of factors on
 * the variable, SH = 0 (set by the caller).  After n sweeps the current half is (parity + n) & 1.
 */
typedef struct lhvi_gabp {
    int32_t dtype;                 /* LHVI_F32 | LHVI_F64 */
    int32_t parity;                /* 0 | 1: the half of P / H / SP / SH that holds the current state */
    int64_t n_vars, n_edges;
    const int32_t* src;            /* [n_edges] */
    const int32_t* dst;            /* [n_edges] */
    const int32_t* rev;            /* [n_edges] */
    const void* coef;              /* [5][n_edges] */
    const void* jd;                /* [n_vars] */
    const void* hd;                /* [n_vars] */
    void* P;                       /* [2][n_edges] */
    void* H;                       /* [2][n_edges] */
    void* SP;                      /* [2][n_vars] */
    void* SH;                      /* [2][n_vars] */
} lhvi_gabp;

/* n_sweeps synchronous sweeps (2 launches each: seed the new totals with jd / hd, update every edge). */
int lhvi_gabp_sweeps(const lhvi_gabp* g, int32_t n_sweeps, void* stream);
/* mean[i] = SH[i] / SP[i], var[i] = 1 / SP[i] from the current half (GaBP.py:183-195). */
int lhvi_gabp_marginals(const lhvi_gabp* g, void* mean, void* var, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LHVI_H */
