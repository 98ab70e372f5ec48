/* lhvi_lift.h -- host-side (CPU) C ABI of the lifting passes that sit in front of the device loop.
 *
 * The reference compresses a ground graph by colour passing before (LiftedVarInference.py:14-26)
 * and, in the coarse-to-fine engine, between the blocks of ten iterations
 * (C2FVarInference.py:33-61,301-352) with Python sets of objects
 * (CompressedGraphWithObs.py:47-76 split_rvs, :152-175 split_factors, :249-271 run).  Here the same
 * fixed point is computed on index arrays with hash tables: linear passes, no sort, no object per
 * ground variable.  `liblhvi_lift.so` is plain C++ (no CUDA); `lifting.colour_passing` binds it
 * through ctypes and keeps a numpy implementation of the same passes as its cross-check
 * (tests/test_lifting.py compares the partitions of the two on every model of the suite).
 *
 * Conventions: plain pointers and sizes, the caller owns every buffer, return value >= 0 on success
 * (a count) and a negative code on bad arguments or allocation failure; not thread-safe per call
 * site only in the sense that the caller's buffers are written.
 */
#ifndef LHVI_LIFT_H
#define LHVI_LIFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LHVI_LIFT_ABI_VERSION 1
#define LHVI_LIFT_MAX_ARITY 16

/* One block of ground factors that share a potential object (lifting.FactorBlock). */
typedef struct lhvi_lift_block {
    const int64_t *args;  /* [n * arity] variable indices, row-major (factor i: args[i*arity .. ]) */
    int64_t n;            /* factors in the block */
    int32_t arity;        /* arguments per factor, 1..LHVI_LIFT_MAX_ARITY */
    int32_t symmetric;    /* != 0: argument order is not part of a factor's key (Potential.symmetric,
                             CompressedGraphWithObs.py:161-164) */
    int64_t *colour;      /* [n] in: initial colour (blocks whose potentials compare equal share it);
                             out: dense factor class ids, one id space over all blocks */
} lhvi_lift_block;

int32_t lhvi_lift_abi_version(void);

/* Coarsest equitable refinement of `var_colour` (CompressedGraph.run, :264-271): until the number of
 * variable classes stops growing, split the factors by (own class, classes of the arguments) and the
 * variables by (own class, multiset of the classes of the incident factors, positions not part of
 * the key).  The multiset is compared through two independent 64-bit sums of mixed class ids and
 * the own class exactly; factor keys are compared exactly.
 *   var_colour  [n_vars] in: start colouring (any int64 labels >= 0); out: dense class ids in order
 *               of first appearance.
 *   sweeps_out  optional: number of sweeps done.
 * Returns the number of variable classes, or < 0: -1 null pointer / bad size, -2 bad arity,
 * -3 a variable index out of range, -4 out of memory, -6 more than 2^31 - 1 variables or factors.
 * The passes run on all OpenMP threads (OMP_NUM_THREADS): the hash space of the keys is split over
 * the threads, every thread keeps the first item of each key of its share, and class ids are handed
 * out in one streaming pass in order of first appearance -- the result does not depend on the thread
 * count. */
int64_t lhvi_lift_colour_passing(int64_t n_vars, int64_t *var_colour, lhvi_lift_block *blocks,
                                 int32_t n_blocks, int32_t max_sweeps, int32_t *sweeps_out);

/* The same on a prepared graph, for callers that refine the same ground graph again and again (the
 * coarse-to-fine engine: once per block of ten iterations): `lhvi_lift_graph_create` validates the
 * blocks, builds the incidence lists by variable and allocates the scratch buffers of the sweeps once;
 * the argument arrays stay the caller's and must outlive the graph (`colour` of the blocks is ignored
 * here).  `status` (optional) receives 0 or the negative code when NULL is returned.
 * `lhvi_lift_graph_colour_passing` takes the factor colours per block in `factor_colour[b]`
 * (in: initial colour, out: class ids) and otherwise behaves as `lhvi_lift_colour_passing`. */
typedef struct lhvi_lift_graph lhvi_lift_graph;
lhvi_lift_graph *lhvi_lift_graph_create(int64_t n_vars, const lhvi_lift_block *blocks, int32_t n_blocks,
                                        int32_t *status);
void lhvi_lift_graph_destroy(lhvi_lift_graph *graph);
int64_t lhvi_lift_graph_colour_passing(lhvi_lift_graph *graph, int64_t *var_colour,
                                       int64_t *const *factor_colour, int32_t max_sweeps,
                                       int32_t *sweeps_out);

/* Evidence split of the coarse-to-fine engine (CompressedGraph.split_evidence, :236-247, with
 * SuperRV.split_by_evidence, :78-130), repeated until nothing changes: every class flagged in
 * `may_split` whose members' values spread more than `epsilon` (standard deviation) is cut by a
 * one-dimensional k-means over the histogram of its values -- centroids start at the first k
 * distinct values in member order, `iterations` Lloyd sweeps, members go to the nearest centroid --
 * piece 0 keeps the class id, the other non-empty pieces get new ids from n_classes upwards; every
 * piece records its centroid as the class value; pieces whose variance still exceeds `epsilon` stay
 * (or become) flagged.  Arithmetic follows numpy's (pairwise sums in the variances, first-minimum
 * arg-min in which a NaN distance wins) so that the result equals lifting.py's statement bit for bit.
 *   var_colour   [n_vars] in/out class of every variable
 *   value        [n_vars] evidence values (read for members of flagged classes only)
 *   capacity     length of may_split / has_centroid / centroid; must be >= n_classes + number of
 *                members of flagged classes (every new class has at least one member)
 * Returns the new number of classes, or < 0 (-1 bad argument, -4 out of memory, -5 capacity). */
int64_t lhvi_lift_split_evidence(int64_t n_vars, int64_t *var_colour, const double *value,
                                 int64_t n_classes, int64_t capacity, uint8_t *may_split,
                                 uint8_t *has_centroid, double *centroid, double epsilon, int32_t k,
                                 int32_t iterations);

/* Dense ids (order of first appearance) of 64-bit keys; returns the number of distinct keys or < 0. */
int64_t lhvi_lift_rank64(const uint64_t *key, int64_t n, int64_t *ids);

#ifdef __cplusplus
}
#endif
#endif
