// Host-side lifting passes (include/lhvi_lift.h): colour passing on index arrays with open-addressing
// hash tables, and the evidence k-means of the coarse-to-fine engine.  Follows
// CompressedGraphWithObs.py:47-76 (split_rvs), :152-175 (split_factors), :249-271 (run), :78-130
// (split_by_evidence) and :236-247 (split_evidence) of the reference; lifting.py holds the numpy
// statement of the same passes.  LHVI_LIFT_TIMING=1 prints the phase times of a colour passing call.
#include "lhvi_lift.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <vector>

namespace {

inline uint64_t mix(uint64_t x, uint64_t seed) {        // splitmix64 finaliser
    uint64_t z = (x + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Open-addressing table from a hash to a dense class id.  The caller resolves hash matches with an
// exact comparison against the class's stored key and creates the class on a miss.
struct Table {
    struct Slot {
        uint64_t tag;               // full hash of the class's key
        int64_t id;                 // class id or -1
    };
    std::vector<Slot> slot;         // one cache line per probe
    uint64_t mask = 0;

    int64_t count = 0;              // classes stored

    // Sized for the classes, not the items: a table of a few thousand classes stays in cache however
    // many items are looked up.  Grows (rehash by the stored tags) when half full.
    bool reserve(int64_t classes) {
        uint64_t cap = 1024;
        while (cap < static_cast<uint64_t>(classes) * 2 + 2) cap <<= 1;
        try {
            slot.assign(cap, Slot{0, -1});
        } catch (const std::bad_alloc &) {
            return false;
        }
        mask = cap - 1;
        count = 0;
        return true;
    }
    void clear() {
        std::fill(slot.begin(), slot.end(), Slot{0, -1});
        count = 0;
    }
    void grow() {                   // may throw std::bad_alloc; callers run inside try blocks
        std::vector<Slot> old;
        old.swap(slot);
        slot.assign(old.size() * 2, Slot{0, -1});
        mask = slot.size() - 1;
        for (const Slot &s : old) {
            if (s.id < 0) continue;
            uint64_t p = s.tag & mask;
            while (slot[p].id >= 0) p = (p + 1) & mask;
            slot[p] = s;
        }
    }
    void prefetch(uint64_t h) const { __builtin_prefetch(&slot[h & mask], 1, 1); }

    template <class Same, class Create>
    int64_t find_or_insert(uint64_t h, Same same, Create create) {
        uint64_t p = h & mask;
        for (;;) {
            Slot &s = slot[p];
            if (s.id < 0) {
                if (static_cast<uint64_t>(count + 1) * 2 > slot.size()) {
                    grow();
                    return find_or_insert(h, same, create);
                }
                s.id = create();
                s.tag = h;
                ++count;
                return s.id;
            }
            if (s.tag == h && same(s.id)) return s.id;
            p = (p + 1) & mask;
        }
    }
};

constexpr int64_t AHEAD = 16;       // items whose table slot / gathered operands are prefetched ahead

struct Mixed {
    uint64_t m1, m2;
};

// Dense ids, in order of first appearance, of n items given their hashes and an exact equality
// `same(i, j)` on item indices.  The hash space is split over the OpenMP threads: every thread scans
// all hashes and keeps, in a private table, the first item of every key of its share; ids are then
// handed out in one streaming pass, so they do not depend on the number of threads.
// `first` is scratch (n entries); returns the number of distinct keys or -4.
template <class Same>
int64_t rank_first(int64_t n, const uint64_t *hs, Same same, std::vector<int32_t> &first, int64_t *ids,
                   std::vector<Table> &tables, int64_t expect_classes) {
    int threads = 1;
#ifdef _OPENMP
    if (n > 65536) threads = std::max(1, omp_get_max_threads());
#endif
    if (static_cast<int>(tables.size()) < threads) tables.resize(static_cast<size_t>(threads));
    first.resize(static_cast<size_t>(n));
    int oom = 0;
    int32_t *fst = first.data();
#pragma omp parallel num_threads(threads)
    {
        int t = 0, nt = 1;
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        Table &tb = tables[static_cast<size_t>(t)];
        try {
            if (!tb.reserve(2 * expect_classes / nt + 1024)) throw std::bad_alloc();
            const uint64_t unt = static_cast<uint64_t>(nt), ut = static_cast<uint64_t>(t);
            for (int64_t i = 0; i < n; ++i) {
                if (i + AHEAD < n) {
                    const uint64_t hn = hs[i + AHEAD];
                    if (nt == 1 || (((hn >> 40) * unt) >> 24) == ut) tb.prefetch(hn);
                }
                const uint64_t h = hs[i];
                if (nt > 1 && (((h >> 40) * unt) >> 24) != ut) continue;
                fst[i] = static_cast<int32_t>(tb.find_or_insert(
                    h, [&](int64_t c) { return same(c, i); }, [&]() { return i; }));
            }
        } catch (const std::bad_alloc &) {
#pragma omp atomic write
            oom = 1;
        }
    }
    if (oom) return -4;
    int64_t next = 0;
    for (int64_t i = 0; i < n; ++i) ids[i] = fst[i] == i ? next++ : ids[fst[i]];
    return next;
}


}  // namespace

extern "C" int32_t lhvi_lift_abi_version(void) { return LHVI_LIFT_ABI_VERSION; }

extern "C" int64_t lhvi_lift_rank64(const uint64_t *key, int64_t n, int64_t *ids) {
    if (n < 0 || (n > 0 && (!key || !ids))) return -1;
    if (n > INT32_MAX) return -6;
    try {
        std::vector<uint64_t> hs(static_cast<size_t>(n));
#pragma omp parallel for schedule(static) if (n > 65536)
        for (int64_t i = 0; i < n; ++i) hs[i] = mix(key[i], 0x243F6A8885A308D3ull);
        std::vector<int32_t> first;
        std::vector<Table> tables;
        return rank_first(n, hs.data(), [&](int64_t a, int64_t b) { return key[a] == key[b]; }, first, ids, tables, 1024);
    } catch (const std::bad_alloc &) {
        return -4;
    }
}

// A ground graph prepared for repeated colour passing: validated blocks, incidences by variable and
// the scratch buffers of the sweeps.  The argument arrays stay the caller's and must outlive it.
struct lhvi_lift_graph {
    int64_t n_vars = 0, n_fac = 0, width = 2;
    std::vector<lhvi_lift_block> blocks;            // colour pointers are set per call
    std::vector<int64_t> foff, inc_ptr;
    std::vector<int32_t> inc_fac, key32, first, header;
    std::vector<int64_t> vcol, vnew, fid;
    std::vector<uint64_t> hash64, vhash, H1, H2;
    std::vector<Mixed> mixed;
    std::vector<Table> tables;                      // one per thread
};

extern "C" lhvi_lift_graph *lhvi_lift_graph_create(int64_t n_vars, const lhvi_lift_block *blocks_in, int32_t n_blocks,
                                                   int32_t *status) {
    auto fail = [&](int32_t code) {
        if (status) *status = code;
        return static_cast<lhvi_lift_graph *>(nullptr);
    };
    if (n_vars < 0 || n_blocks < 0 || (n_blocks > 0 && !blocks_in)) return fail(-1);
    int64_t n_fac = 0;
    for (int32_t b = 0; b < n_blocks; ++b) {
        const lhvi_lift_block &B = blocks_in[b];
        if (B.n < 0 || (B.n > 0 && !B.args)) return fail(-1);
        if (B.arity < 1 || B.arity > LHVI_LIFT_MAX_ARITY) return fail(-2);
        int bad = 0;
        const int64_t total = B.n * B.arity;
#pragma omp parallel for schedule(static) reduction(| : bad) if (total > 65536)
        for (int64_t i = 0; i < total; ++i) bad |= (B.args[i] < 0 || B.args[i] >= n_vars);
        if (bad) return fail(-3);
        n_fac += B.n;
    }
    if (n_vars > INT32_MAX || n_fac > INT32_MAX) return fail(-6);       // class ids are keyed in 32 bits
    lhvi_lift_graph *g = nullptr;
    try {
        g = new lhvi_lift_graph();
        g->n_vars = n_vars;
        g->n_fac = n_fac;
        g->blocks.assign(blocks_in, blocks_in + n_blocks);
        g->foff.assign(static_cast<size_t>(n_blocks) + 1, 0);
        for (int32_t b = 0; b < n_blocks; ++b) g->foff[b + 1] = g->foff[b] + blocks_in[b].n;
        // incidences by variable (one entry per argument position)
        g->inc_ptr.assign(static_cast<size_t>(n_vars) + 1, 0);
        for (int32_t b = 0; b < n_blocks; ++b) {
            const int64_t total = blocks_in[b].n * blocks_in[b].arity;
            for (int64_t i = 0; i < total; ++i) ++g->inc_ptr[blocks_in[b].args[i] + 1];
        }
        for (int64_t v = 0; v < n_vars; ++v) g->inc_ptr[v + 1] += g->inc_ptr[v];
        g->inc_fac.resize(static_cast<size_t>(g->inc_ptr[n_vars]));
        {
            std::vector<int64_t> fill(g->inc_ptr.begin(), g->inc_ptr.end() - 1);
            for (int32_t b = 0; b < n_blocks; ++b) {
                const lhvi_lift_block &B = blocks_in[b];
                for (int64_t i = 0; i < B.n; ++i)
                    for (int32_t j = 0; j < B.arity; ++j)
                        g->inc_fac[fill[B.args[i * B.arity + j]]++] = static_cast<int32_t>(g->foff[b] + i);
            }
        }
        g->mixed.resize(static_cast<size_t>(n_fac));
        g->width = 2;
        for (int32_t b = 0; b < n_blocks; ++b) g->width = std::max<int64_t>(g->width, blocks_in[b].arity + 2);
        g->key32.assign(static_cast<size_t>(n_fac * g->width), 0);          // padding stays zero
        g->hash64.resize(static_cast<size_t>(n_fac));
        g->vcol.resize(static_cast<size_t>(n_vars));
        g->vnew.resize(static_cast<size_t>(n_vars));
        g->H1.resize(static_cast<size_t>(n_vars));
        g->H2.resize(static_cast<size_t>(n_vars));
        g->vhash.resize(static_cast<size_t>(n_vars));
        g->fid.resize(static_cast<size_t>(n_fac));
        g->header.resize(static_cast<size_t>(n_blocks));
        for (int32_t b = 0; b < n_blocks; ++b) g->header[b] = blocks_in[b].arity | (blocks_in[b].symmetric ? 256 : 0);
    } catch (const std::bad_alloc &) {
        delete g;
        return fail(-4);
    }
    if (status) *status = 0;
    return g;
}

extern "C" void lhvi_lift_graph_destroy(lhvi_lift_graph *g) { delete g; }

extern "C" int64_t lhvi_lift_graph_colour_passing(lhvi_lift_graph *g, int64_t *var_colour, int64_t *const *factor_colour,
                                                  int32_t max_sweeps, int32_t *sweeps_out) {
    if (!g || (g->n_vars > 0 && !var_colour) || (!g->blocks.empty() && !factor_colour)) return -1;
    const int64_t n_vars = g->n_vars;
    const int32_t n_blocks = static_cast<int32_t>(g->blocks.size());
    for (int32_t b = 0; b < n_blocks; ++b) {
        lhvi_lift_block &B = g->blocks[b];
        B.colour = factor_colour[b];
        if (B.n > 0 && !B.colour) return -1;
        int bad = 0;
        for (int64_t i = 0; i < B.n; ++i) bad |= (B.colour[i] < 0 || B.colour[i] > INT32_MAX);
        if (bad) return -1;
    }
    for (int64_t v = 0; v < n_vars; ++v)
        if (var_colour[v] < 0) return -1;
    lhvi_lift_block *blocks = g->blocks.data();
    std::vector<int64_t> &foff = g->foff, &inc_ptr = g->inc_ptr;
    std::vector<int64_t> &vcol = g->vcol, &vnew = g->vnew, &fid = g->fid;
    std::vector<int32_t> &inc_fac = g->inc_fac, &key32 = g->key32, &header = g->header;
    std::vector<uint64_t> &hash64 = g->hash64, &vhash = g->vhash, &H1 = g->H1, &H2 = g->H2;
    std::vector<Mixed> &mixed = g->mixed;
    const int64_t n_fac = g->n_fac;
    try {
        // dense start colouring, order of first appearance
        int64_t n_classes = 0;
        {
            uint64_t *vh = vhash.data();
#pragma omp parallel for schedule(static) if (n_vars > 65536)
            for (int64_t v = 0; v < n_vars; ++v) vh[v] = mix(static_cast<uint64_t>(var_colour[v]), 1);
            n_classes = rank_first(n_vars, vh, [&](int64_t a, int64_t b) { return var_colour[a] == var_colour[b]; },
                                   g->first, vcol.data(), g->tables, 1024);
            if (n_classes < 0) return n_classes;
        }
        const int64_t W = g->width;          // key row: arity | symmetric << 8, own class, argument classes, zero padding
        auto same_factor = [&](int64_t fa, int64_t fb) {
            const int32_t *ka = key32.data() + fa * W, *kb = key32.data() + fb * W;
            return std::equal(ka, ka + W, kb);
        };
        auto same_variable = [&](int64_t a, int64_t b) { return vcol[a] == vcol[b] && H1[a] == H1[b] && H2[a] == H2[b]; };

        int64_t before = -1, n_fclasses = 1024;
        int32_t sweeps = 0;
        const bool timing = std::getenv("LHVI_LIFT_TIMING") != nullptr;
        double t_gather = 0, t_lookup = 0, t_scatter = 0, t_vars = 0;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        while (before != n_classes && sweeps < max_sweeps) {
            before = n_classes;
            ++sweeps;
            // ---- factors: (own class, classes of the arguments; sorted for a symmetric potential):
            // gather every factor's key (32-bit ids) and hash it, then rank the keys
            double t0 = now();
            for (int32_t b = 0; b < n_blocks; ++b) {
                lhvi_lift_block &B = blocks[b];
                const int32_t arity = B.arity, hdr = header[b];
                int32_t *keys = key32.data() + foff[b] * W;
                uint64_t *hs = hash64.data() + foff[b];
#pragma omp parallel for schedule(static) if (B.n > 32768)
                for (int64_t i = 0; i < B.n; ++i) {
                    const int64_t *a = B.args + i * arity;
                    int32_t *k = keys + i * W;
                    k[0] = hdr;
                    k[1] = static_cast<int32_t>(B.colour[i]);
                    for (int32_t j = 0; j < arity; ++j) k[2 + j] = static_cast<int32_t>(vcol[a[j]]);
                    if (B.symmetric) std::sort(k + 2, k + 2 + arity);
                    uint64_t h = mix(static_cast<uint64_t>(k[1]), 0x452821E638D01377ull + static_cast<uint64_t>(hdr));
                    for (int32_t j = 0; j < arity; ++j) h = mix(h ^ static_cast<uint64_t>(k[2 + j]), 0xBE5466CF34E90C6Cull);
                    hs[i] = h;
                }
            }
            double t1 = now();
            t_gather += t1 - t0;
            n_fclasses = rank_first(n_fac, hash64.data(), same_factor, g->first, fid.data(), g->tables, n_fclasses);
            if (n_fclasses < 0) return n_fclasses;
            for (int32_t b = 0; b < n_blocks; ++b)
                std::copy(fid.begin() + foff[b], fid.begin() + foff[b + 1], blocks[b].colour);
            double t2 = now();
            t_lookup += t2 - t1;
            // ---- variables: (own class, multiset of incident factor classes) through two 64-bit sums:
            // every factor's two mixed class ids once, then every variable sums over its incidences
            // (no atomics: a group-level variable with 10^5 incidences would serialise them)
            {
                Mixed *out = mixed.data();
                const int64_t *id = fid.data();
#pragma omp parallel for schedule(static) if (n_fac > 32768)
                for (int64_t f = 0; f < n_fac; ++f) {
                    uint64_t c = static_cast<uint64_t>(id[f]);
                    out[f] = Mixed{mix(c, 0x243F6A8885A308D3ull), mix(c, 0x13198A2E03707344ull)};
                }
                const int64_t *ptr = inc_ptr.data();
                const int32_t *inc = inc_fac.data();
                const int64_t *vc = vcol.data();
                uint64_t *h1 = H1.data(), *h2 = H2.data(), *vh = vhash.data();
#pragma omp parallel for schedule(dynamic, 2048) if (n_vars > 32768)
                for (int64_t v = 0; v < n_vars; ++v) {
                    uint64_t a1 = 0, a2 = 0;
                    for (int64_t e = ptr[v]; e < ptr[v + 1]; ++e) {
                        const Mixed &m = out[inc[e]];
                        a1 += m.m1;
                        a2 += m.m2;
                    }
                    h1[v] = a1;
                    h2[v] = a2;
                    vh[v] = mix(static_cast<uint64_t>(vc[v]), 0xA4093822299F31D0ull) + a1 + mix(a2, 0x082EFA98EC4E6C89ull);
                }
            }
            double t3 = now();
            t_scatter += t3 - t2;
            n_classes = rank_first(n_vars, vhash.data(), same_variable, g->first, vnew.data(), g->tables, n_classes);
            if (n_classes < 0) return n_classes;
            vcol.swap(vnew);
            t_vars += now() - t3;
        }
        if (timing)
            std::fprintf(stderr, "[lhvi_lift] %d sweeps: gather %.3f s, factor lookup %.3f s, scatter %.3f s, variable lookup %.3f s\n",
                         sweeps, t_gather, t_lookup, t_scatter, t_vars);
        std::copy(vcol.begin(), vcol.end(), var_colour);
        if (sweeps_out) *sweeps_out = sweeps;
        return n_classes;
    } catch (const std::bad_alloc &) {
        return -4;
    }
}

extern "C" int64_t lhvi_lift_colour_passing(int64_t n_vars, int64_t *var_colour, lhvi_lift_block *blocks,
                                            int32_t n_blocks, int32_t max_sweeps, int32_t *sweeps_out) {
    if (n_vars > 0 && !var_colour) return -1;
    int32_t status = 0;
    lhvi_lift_graph *g = lhvi_lift_graph_create(n_vars, blocks, n_blocks, &status);
    if (!g) return status;
    std::vector<int64_t *> colours(static_cast<size_t>(n_blocks));
    for (int32_t b = 0; b < n_blocks; ++b) colours[b] = blocks[b].colour;
    int64_t rc = lhvi_lift_graph_colour_passing(g, var_colour, colours.data(), max_sweeps, sweeps_out);
    lhvi_lift_graph_destroy(g);
    return rc;
}

// ---- evidence split ------------------------------------------------------------------------------
namespace {

// numpy's pairwise summation of a contiguous float64 vector (the rounding np.var / np.mean see)
double pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

double np_var(const double *a, int64_t n, std::vector<double> &scratch) {     // ndarray.var(), ddof = 0
    double mean = pairwise_sum(a, n) / static_cast<double>(n);
    scratch.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        double d = a[i] - mean;
        scratch[i] = d * d;
    }
    return pairwise_sum(scratch.data(), n) / static_cast<double>(n);
}

// np.abs(x - centroids).argmin(): the first NaN wins, else the first minimum
inline int nearest(double x, const double *cen, int k) {
    int best = 0;
    double bd = std::fabs(x - cen[0]);
    if (std::isnan(bd)) return 0;
    for (int c = 1; c < k; ++c) {
        double d = std::fabs(x - cen[c]);
        if (std::isnan(d)) return c;
        if (d < bd) {
            bd = d;
            best = c;
        }
    }
    return best;
}

}  // namespace

extern "C" int64_t lhvi_lift_split_evidence(int64_t n_vars, int64_t *var_colour, const double *value,
                                            int64_t n_classes, int64_t capacity, uint8_t *may_split,
                                            uint8_t *has_centroid, double *centroid, double epsilon, int32_t k,
                                            int32_t iterations) {
    if (n_vars < 0 || n_classes < 0 || capacity < n_classes || k < 1 || iterations < 0) return -1;
    if (n_vars > 0 && (!var_colour || !value)) return -1;
    if (capacity > 0 && (!may_split || !has_centroid || !centroid)) return -1;
    for (int64_t v = 0; v < n_vars; ++v)
        if (var_colour[v] < 0 || var_colour[v] >= n_classes) return -1;
    try {
        std::vector<int64_t> cand, start, fill, order;
        std::vector<double> vals, scratch, piece, hist_val, hist_cnt, hist_w, cen, mass, tot;
        std::vector<std::pair<double, int64_t>> sorted;
        std::vector<int> owner;
        std::vector<int64_t> piece_id;
        bool changed = true;
        while (changed) {
            changed = false;
            int64_t next_id = n_classes;
            // members of the flagged classes grouped by class (ascending id), ascending inside a class
            start.assign(static_cast<size_t>(n_classes) + 1, 0);
            int64_t n_cand = 0;
            for (int64_t v = 0; v < n_vars; ++v)
                if (may_split[var_colour[v]]) {
                    ++start[var_colour[v] + 1];
                    ++n_cand;
                }
            if (n_cand == 0) break;
            for (int64_t c = 0; c < n_classes; ++c) start[c + 1] += start[c];
            cand.resize(static_cast<size_t>(n_cand));
            fill.assign(start.begin(), start.end() - 1);
            for (int64_t v = 0; v < n_vars; ++v)
                if (may_split[var_colour[v]]) cand[fill[var_colour[v]]++] = v;
            for (int64_t cid = 0; cid < n_classes; ++cid) {
                int64_t m = start[cid + 1] - start[cid];
                if (m == 0) continue;
                const int64_t *members = cand.data() + start[cid];
                vals.resize(static_cast<size_t>(m));
                for (int64_t i = 0; i < m; ++i) vals[i] = value[members[i]];
                if (!(std::sqrt(np_var(vals.data(), m, scratch)) > epsilon)) continue;
                // histogram of the values in first-seen order
                sorted.resize(static_cast<size_t>(m));
                for (int64_t i = 0; i < m; ++i) sorted[i] = {vals[i], i};
                std::sort(sorted.begin(), sorted.end());
                order.clear();            // first index of every distinct value
                hist_cnt.clear();
                for (int64_t i = 0; i < m;) {
                    int64_t j = i;
                    while (j < m && sorted[j].first == sorted[i].first) ++j;
                    order.push_back(sorted[i].second);
                    hist_cnt.push_back(static_cast<double>(j - i));
                    i = j;
                }
                int64_t n_uniq = static_cast<int64_t>(order.size());
                int kk = static_cast<int>(std::min<int64_t>(k, n_uniq));
                if (m <= 1 || kk <= 1) {
                    if (m == 1) may_split[cid] = 0;
                    continue;
                }
                // sort the histogram by first appearance
                std::vector<int64_t> perm(static_cast<size_t>(n_uniq));
                for (int64_t i = 0; i < n_uniq; ++i) perm[i] = i;
                std::sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return order[a] < order[b]; });
                hist_val.resize(static_cast<size_t>(n_uniq));
                hist_w.resize(static_cast<size_t>(n_uniq));
                for (int64_t i = 0; i < n_uniq; ++i) {
                    hist_val[i] = vals[order[perm[i]]];
                    hist_w[i] = hist_cnt[perm[i]];
                }
                cen.assign(hist_val.begin(), hist_val.begin() + kk);
                mass.resize(kk);
                tot.resize(kk);
                for (int32_t it = 0; it < iterations; ++it) {
                    std::fill(mass.begin(), mass.end(), 0.0);
                    std::fill(tot.begin(), tot.end(), 0.0);
                    for (int64_t i = 0; i < n_uniq; ++i) {
                        int o = nearest(hist_val[i], cen.data(), kk);
                        mass[o] += hist_w[i];
                        tot[o] += hist_val[i] * hist_w[i];
                    }
                    for (int c = 0; c < kk; ++c) cen[c] = tot[c] / mass[c];
                }
                owner.resize(static_cast<size_t>(m));
                for (int64_t i = 0; i < m; ++i) owner[i] = nearest(vals[i], cen.data(), kk);
                // pieces: 0 keeps the id, the other non-empty ones get new ids
                piece_id.assign(static_cast<size_t>(kk), -1);
                piece_id[0] = cid;
                has_centroid[cid] = 1;
                centroid[cid] = cen[0];
                int n_pieces = 1;
                for (int c = 1; c < kk; ++c) {
                    bool any = false;
                    for (int64_t i = 0; i < m && !any; ++i) any = owner[i] == c;
                    if (!any) continue;
                    if (next_id >= capacity) return -5;
                    piece_id[c] = next_id;
                    for (int64_t i = 0; i < m; ++i)
                        if (owner[i] == c) var_colour[members[i]] = next_id;
                    has_centroid[next_id] = 1;
                    centroid[next_id] = cen[c];
                    may_split[next_id] = 0;
                    ++next_id;
                    ++n_pieces;
                }
                if (n_pieces > 1) {
                    changed = true;
                    for (int c = 0; c < kk; ++c) {
                        if (piece_id[c] < 0) continue;
                        piece.clear();
                        for (int64_t i = 0; i < m; ++i)
                            if (owner[i] == c) piece.push_back(vals[i]);
                        bool wide = !piece.empty() &&
                                    np_var(piece.data(), static_cast<int64_t>(piece.size()), scratch) > epsilon;
                        if (wide)
                            may_split[piece_id[c]] = 1;
                        else if (piece_id[c] != cid)
                            may_split[piece_id[c]] = 0;
                    }
                }
            }
            n_classes = next_id;
        }
    } catch (const std::bad_alloc &) {
        return -4;
    }
    return n_classes;
}
