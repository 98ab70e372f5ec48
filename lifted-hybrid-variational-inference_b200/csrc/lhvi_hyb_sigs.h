// Which record groups the kernels of lhvi_hyb_impl.cuh serve (host-side tables shared by the
// dispatcher lhvi_spec.cu and the instantiation units).
#pragma once
#include "lhvi_common.cuh"

namespace lhvi {

#ifndef LHVI_FLAVOURS_DEFINED
#define LHVI_FLAVOURS_DEFINED
enum Flavour { kFull = 0, kPure = 1, kNode = 2 };
#endif

// ---- dispatch ----------------------------------------------------------------------------------------
//
// X(ND, NC, NE, FL): the signatures with a kernel, for D in {2, 3} (a signature with 81 or more grid
// points per component is left to the generic kernel).  Groups without a continuous argument do not
// look at the quadrature rule: they are instantiated once (T = 1) and serve every T.
#define LHVI_HYB_DISCRETE(X)                                                                  \
    X(1, 0, 0, kNode) X(1, 0, 0, kPure) X(1, 0, 1, kPure) X(1, 0, 2, kPure)                    \
    X(2, 0, 0, kFull) X(2, 0, 1, kFull) X(3, 0, 0, kFull) X(3, 0, 1, kFull) X(4, 0, 0, kFull)
#define LHVI_HYB_MIXED(X)                                                                     \
    X(1, 1, 0, kFull) X(1, 1, 1, kFull) X(1, 1, 2, kFull) X(1, 2, 0, kFull) X(1, 2, 1, kFull)  \
    X(2, 1, 0, kFull) X(2, 1, 1, kFull)
// continuous-only signatures at quadrature degrees the walk kernels of lhvi_spec_impl.cuh do not cover
#define LHVI_HYB_CONT(X)                                                                      \
    X(0, 1, 0, kNode) X(0, 1, 0, kPure) X(0, 1, 1, kPure) X(0, 1, 2, kPure)                    \
    X(0, 2, 0, kFull) X(0, 2, 1, kFull)
// records whose arguments are all observed (constants of the free energy): no rule involved, any T
#define LHVI_HYB_CONST(X) X(0, 0, 1, kPure) X(0, 0, 2, kPure)

constexpr int hyb_code(int nd, int nc, int ne, int fl) { return ((nd * 8 + nc) * 8 + ne) * 4 + fl; }

inline int hyb_flavour(const lhvi_group* g) { return g->node ? kNode : (g->pure ? kPure : kFull); }

// D of a group whose hidden discrete arguments all have the same number of states (0: none or mixed)
inline int hyb_states(const lhvi_group* g) {
    if (g->nd < 1) return 0;
    for (int a = 1; a < g->nd; ++a)
        if (g->dims[a] != g->dims[0]) return 0;
    return g->dims[0];
}

inline bool hyb_available(const lhvi_model* m, const lhvi_group* g) {
    if (m->K < 1 || m->K > 3 || g->ng != 0) return false;
    const int code = hyb_code(g->nd, g->nc, g->ne, hyb_flavour(g));
    const int D = hyb_states(g);
    if (g->nd > 0) {
        if (D != 2 && D != 3) return false;
        if (D == 3 && g->nd > 3) return false;
        switch (code) {
#define X(ND_, NC_, NE_, FL_) case hyb_code(ND_, NC_, NE_, FL_): return true;
            LHVI_HYB_DISCRETE(X)
#undef X
            default: break;
        }
        if (m->T != 3 && !(m->T == 10 && m->K <= 2)) return false;
        if (D == 3 && g->nd + g->nc > (m->T == 3 ? 3 : 2)) return false;
        switch (code) {
#define X(ND_, NC_, NE_, FL_) case hyb_code(ND_, NC_, NE_, FL_): return true;
            LHVI_HYB_MIXED(X)
#undef X
            default: return false;
        }
    }
    if (g->nc == 0) {
        switch (code) {
#define X(ND_, NC_, NE_, FL_) case hyb_code(ND_, NC_, NE_, FL_): return true;
            LHVI_HYB_CONST(X)
#undef X
            default: return false;
        }
    }
    if (!(m->T == 10 && m->K <= 2)) return false;      // T = 3 continuous-only groups: lhvi_spec_impl.cuh
    switch (code) {
#define X(ND_, NC_, NE_, FL_) case hyb_code(ND_, NC_, NE_, FL_): return true;
        LHVI_HYB_CONT(X)
#undef X
        default: return false;
    }
}

}  // namespace lhvi
