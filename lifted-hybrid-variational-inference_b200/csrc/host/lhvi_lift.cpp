// Host-side lifting passes (include/lhvi_lift.h): colour passing on index arrays with open-addressing
// hash tables.  Follows CompressedGraphWithObs.py:47-76 (split_rvs), :152-175 (split_factors),
// :249-271 (run) of the reference; see lifting.colour_passing for the numpy statement of the same
// passes.
#include "lhvi_lift.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace {

inline uint64_t mix(uint64_t x, uint64_t seed) {        // splitmix64 finaliser
    uint64_t z = (x + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Open-addressing table from a hash to the index of the first item seen with that key.  The caller
// resolves hash matches with an exact comparison of the two items.
struct Table {
    std::vector<int64_t> slot;      // item index or -1
    std::vector<uint64_t> tag;      // full hash of the stored item
    uint64_t mask = 0;

    bool reserve(int64_t items) {
        uint64_t cap = 16;
        while (cap < static_cast<uint64_t>(items) * 2 + 2) cap <<= 1;
        try {
            slot.assign(cap, -1);
            tag.assign(cap, 0);
        } catch (const std::bad_alloc &) {
            return false;
        }
        mask = cap - 1;
        return true;
    }
    void clear() { std::fill(slot.begin(), slot.end(), int64_t(-1)); }

    template <class Same>
    int64_t find_or_insert(uint64_t h, int64_t item, Same same) {     // returns the first item with this key
        uint64_t p = h & mask;
        for (;;) {
            int64_t s = slot[p];
            if (s < 0) {
                slot[p] = item;
                tag[p] = h;
                return item;
            }
            if (tag[p] == h && same(s, item)) return s;
            p = (p + 1) & mask;
        }
    }
};

struct FactorRef {
    int32_t block;
    int64_t row;
};

}  // namespace

extern "C" int32_t lhvi_lift_abi_version(void) { return LHVI_LIFT_ABI_VERSION; }

extern "C" int64_t lhvi_lift_rank64(const uint64_t *key, int64_t n, int64_t *ids) {
    if (n < 0 || (n > 0 && (!key || !ids))) return -1;
    Table t;
    if (!t.reserve(n)) return -4;
    std::vector<int64_t> id_of;      // dense id of the item that owns a key, indexed by that item
    try {
        id_of.assign(static_cast<size_t>(n), -1);
    } catch (const std::bad_alloc &) {
        return -4;
    }
    int64_t next = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t first = t.find_or_insert(mix(key[i], 0x243F6A8885A308D3ull), i,
                                         [&](int64_t a, int64_t b) { return key[a] == key[b]; });
        if (first == i) id_of[i] = next++;
        ids[i] = id_of[first];
    }
    return next;
}

extern "C" int64_t lhvi_lift_colour_passing(int64_t n_vars, int64_t *var_colour, lhvi_lift_block *blocks,
                                            int32_t n_blocks, int32_t max_sweeps, int32_t *sweeps_out) {
    if (n_vars < 0 || n_blocks < 0 || (n_vars > 0 && !var_colour) || (n_blocks > 0 && !blocks)) return -1;
    int64_t n_fac = 0, n_inc = 0;
    for (int32_t b = 0; b < n_blocks; ++b) {
        const lhvi_lift_block &B = blocks[b];
        if (B.n < 0 || (B.n > 0 && (!B.args || !B.colour))) return -1;
        if (B.arity < 1 || B.arity > LHVI_LIFT_MAX_ARITY) return -2;
        for (int64_t i = 0; i < B.n * B.arity; ++i)
            if (B.args[i] < 0 || B.args[i] >= n_vars) return -3;
        n_fac += B.n;
        n_inc += B.n * B.arity;
    }
    (void)n_inc;
    std::vector<int64_t> vcol, vnew, fold, h_first;
    std::vector<uint64_t> H1, H2;
    std::vector<int64_t> foff(static_cast<size_t>(n_blocks) + 1, 0);
    Table vt, ft;
    try {
        vcol.assign(var_colour, var_colour + n_vars);
        vnew.assign(static_cast<size_t>(n_vars), 0);
        H1.assign(static_cast<size_t>(n_vars), 0);
        H2.assign(static_cast<size_t>(n_vars), 0);
        fold.assign(static_cast<size_t>(n_fac), 0);
        h_first.assign(static_cast<size_t>(std::max(n_fac, n_vars)), -1);
    } catch (const std::bad_alloc &) {
        return -4;
    }
    if (!vt.reserve(n_vars) || !ft.reserve(n_fac)) return -4;
    for (int32_t b = 0; b < n_blocks; ++b) foff[b + 1] = foff[b] + blocks[b].n;

    // dense start colouring, order of first appearance
    int64_t n_classes = 0;
    {
        for (int64_t v = 0; v < n_vars; ++v) {
            int64_t first = vt.find_or_insert(mix(static_cast<uint64_t>(vcol[v]), 1), v,
                                              [&](int64_t a, int64_t c) { return vcol[a] == vcol[c]; });
            if (first == v) h_first[v] = n_classes++;
            vnew[v] = h_first[first];
        }
        vcol.swap(vnew);
    }

    auto block_of = [&](int64_t f) {          // global factor index -> (block, row)
        int32_t b = static_cast<int32_t>(std::upper_bound(foff.begin(), foff.end(), f) - foff.begin()) - 1;
        return FactorRef{b, f - foff[b]};
    };
    auto key_of = [&](const lhvi_lift_block &B, int64_t row, int64_t *out) {   // classes of the arguments
        const int64_t *a = B.args + row * B.arity;
        for (int32_t j = 0; j < B.arity; ++j) out[j] = vcol[a[j]];
        if (B.symmetric) std::sort(out, out + B.arity);
    };

    int64_t before = -1;
    int32_t sweeps = 0;
    while (before != n_classes && sweeps < max_sweeps) {
        before = n_classes;
        ++sweeps;
        // ---- factors: (own class, classes of the arguments)
        for (int32_t b = 0; b < n_blocks; ++b)
            std::memcpy(fold.data() + foff[b], blocks[b].colour, sizeof(int64_t) * static_cast<size_t>(blocks[b].n));
        ft.clear();
        int64_t n_fclasses = 0;
        for (int32_t b = 0; b < n_blocks; ++b) {
            lhvi_lift_block &B = blocks[b];
            int64_t mine[LHVI_LIFT_MAX_ARITY], theirs[LHVI_LIFT_MAX_ARITY];
            for (int64_t i = 0; i < B.n; ++i) {
                key_of(B, i, mine);
                uint64_t h = mix(static_cast<uint64_t>(fold[foff[b] + i]), 0x452821E638D01377ull + B.arity);
                for (int32_t j = 0; j < B.arity; ++j) h = mix(h ^ static_cast<uint64_t>(mine[j]), 0xBE5466CF34E90C6Cull);
                int64_t self = foff[b] + i;
                int64_t first = ft.find_or_insert(h, self, [&](int64_t s, int64_t) {
                    if (fold[s] != fold[self]) return false;
                    FactorRef r = block_of(s);
                    const lhvi_lift_block &R = blocks[r.block];
                    if (R.arity != B.arity || (R.symmetric != 0) != (B.symmetric != 0)) return false;
                    key_of(R, r.row, theirs);
                    return std::equal(mine, mine + B.arity, theirs);
                });
                if (first == self) h_first[self] = n_fclasses++;
                B.colour[i] = h_first[first];
            }
        }
        // ---- variables: (own class, multiset of incident factor classes) through two 64-bit sums
        std::fill(H1.begin(), H1.end(), 0);
        std::fill(H2.begin(), H2.end(), 0);
        for (int32_t b = 0; b < n_blocks; ++b) {
            const lhvi_lift_block &B = blocks[b];
            for (int64_t i = 0; i < B.n; ++i) {
                uint64_t c = static_cast<uint64_t>(B.colour[i]);
                uint64_t m1 = mix(c, 0x243F6A8885A308D3ull), m2 = mix(c, 0x13198A2E03707344ull);
                const int64_t *a = B.args + i * B.arity;
                for (int32_t j = 0; j < B.arity; ++j) {
                    H1[a[j]] += m1;
                    H2[a[j]] += m2;
                }
            }
        }
        vt.clear();
        n_classes = 0;
        for (int64_t v = 0; v < n_vars; ++v) {
            uint64_t h = mix(static_cast<uint64_t>(vcol[v]), 0xA4093822299F31D0ull) + H1[v] + mix(H2[v], 0x082EFA98EC4E6C89ull);
            int64_t first = vt.find_or_insert(h, v, [&](int64_t a, int64_t c) {
                return vcol[a] == vcol[c] && H1[a] == H1[c] && H2[a] == H2[c];
            });
            if (first == v) h_first[v] = n_classes++;
            vnew[v] = h_first[first];
        }
        vcol.swap(vnew);
    }
    std::copy(vcol.begin(), vcol.end(), var_colour);
    if (sweeps_out) *sweeps_out = sweeps;
    return n_classes;
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
