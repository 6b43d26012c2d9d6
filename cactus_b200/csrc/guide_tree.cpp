// guide_tree.cpp -- host side of the progressive-POA order (which read is aligned when).
//
// With seeding disabled Cactus still lets abPOA pick the alignment order from a guide tree
// (abpoa_build_guide_tree_partition, abPOA src/abpoa_seed.c:705-722): (w,k)-minimizers of every sequence
// (mm_sketch, :85-156; forward strand only, no homopolymer compression), pairwise min-count Jaccard similarity of the
// minimizer multisets and a greedy order (abpoa_build_guide_tree, :232-325). O(K * L) + O(K^2 * #distinct minimizers)
// per job, negligible next to the DP, so it stays on the host; jobs are processed by OpenMP threads.
// Doubles are summed in the reference's order so that ties resolve identically.
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "host_api.h"

namespace barb200 {

static inline uint64_t invertible_hash(uint64_t key, uint64_t mask) {   // minimap2's hash64, abpoa_seed.c:36-46
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

namespace {
struct Mini { uint64_t x, y; };   // x = hash << 8 | span, y = read << 32 | last position << 1
}

// Windowed minimizer sampling with the reference's treatment of ties: when several k-mers of a window share the
// minimal hash all of them are reported, in sequence order.
static void sketch(const uint8_t *s, int len, int w, int k, uint32_t rid, std::vector<Mini> &out) {
    const uint64_t mask = (1ULL << 2 * k) - 1;
    const Mini none = {UINT64_MAX, UINT64_MAX};
    std::vector<Mini> ring(w, none);
    Mini best = none;
    uint64_t kmer = 0;
    int run = 0, pos = 0, best_pos = 0;           // run: bases since the last ambiguous one
    auto emit_ties = [&](int from, int to, const Mini &m) {
        for (int j = from; j < to; ++j) if (m.x == ring[j].x && m.y != ring[j].y) out.push_back(ring[j]);
    };
    for (int i = 0; i < len; ++i) {
        const int c = s[i];
        Mini cur = none;
        if (c < 4) {
            const int span = run + 1 < k ? run + 1 : k;
            kmer = (kmer << 2 | (uint64_t)c) & mask;
            ++run;
            if (run >= k) { cur.x = invertible_hash(kmer, mask) << 8 | (uint64_t)span; cur.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1; }
        } else run = 0;
        ring[pos] = cur;
        if (run == w + k - 1 && best.x != UINT64_MAX) { emit_ties(pos + 1, w, best); emit_ties(0, pos, best); }
        if (cur.x <= best.x) {
            if (run >= w + k && best.x != UINT64_MAX) out.push_back(best);
            best = cur; best_pos = pos;
        } else if (pos == best_pos) {
            if (run >= w + k - 1 && best.x != UINT64_MAX) out.push_back(best);
            best.x = UINT64_MAX;
            for (int j = pos + 1; j < w; ++j) if (best.x >= ring[j].x) { best = ring[j]; best_pos = j; }
            for (int j = 0; j <= pos; ++j) if (best.x >= ring[j].x) { best = ring[j]; best_pos = j; }
            if (run >= w + k - 1 && best.x != UINT64_MAX) { emit_ties(pos + 1, w, best); emit_ties(0, pos + 1, best); }
        }
        if (++pos == w) pos = 0;
    }
    if (best.x != UINT64_MAX) out.push_back(best);
}

void guide_tree_order(const HostParams &hp, int progressive, int n, const uint8_t *const *seqs, const int *lens, int *order) {
    for (int i = 0; i < n; ++i) order[i] = i;
    if (!(progressive && n > 2)) return;
    std::vector<Mini> mm;
    for (int i = 0; i < n; ++i) sketch(seqs[i], lens[i], hp.w, hp.k, (uint32_t)i, mm);
    if (mm.empty()) return;
    std::sort(mm.begin(), mm.end(), [](const Mini &a, const Mini &b) { return a.x < b.x; });   // only grouping by x matters
    auto tri = [](int64_t i, int64_t j) { return i * (i + 1) / 2 + j; };                          // i >= j
    std::vector<int> hit((size_t)n * (n + 1) / 2, 0), cnt(n);
    for (size_t s = 0; s < mm.size();) {
        size_t e = s;
        std::fill(cnt.begin(), cnt.end(), 0);
        while (e < mm.size() && mm[e].x == mm[s].x) { const int r = (int)(mm[e].y >> 32); ++cnt[r]; ++hit[tri(r, r)]; ++e; }
        for (int a = 0; a < n - 1; ++a) for (int b = a + 1; b < n; ++b) hit[tri(b, a)] += std::min(cnt[a], cnt[b]);
        s = e;
    }
    std::vector<double> jac((size_t)n * (n - 1) / 2, 0.0);
    auto jidx = [](int64_t i, int64_t j) { return i * (i - 1) / 2 + j; };                         // i > j
    double best = -1.0; int bi = -1, bj = -1;
    for (int i = 1; i < n; ++i) for (int j = 0; j < i; ++j) {
        const int shared = hit[tri(i, j)], tot = hit[tri(i, i)] + hit[tri(j, j)] - shared;
        const double v = tot == 0 ? 0.0 : (0.0 + shared) / tot;
        jac[jidx(i, j)] = v;
        if (v > best) { best = v; bi = i; bj = j; }
    }
    int placed = 2; order[0] = bj; order[1] = bi;
    while (placed < n) {
        best = -1.0; int pick = n;
        for (int r1 = 0; r1 < n; ++r1) {
            double v = 0.0;
            for (int i = 0; i < placed; ++i) {
                const int r2 = order[i];
                if (r1 == r2) { v = -1.0; break; }
                v += r1 > r2 ? jac[jidx(r1, r2)] : jac[jidx(r2, r1)];
            }
            if (v > best) { best = v; pick = r1; }
        }
        order[placed++] = pick;
    }
}

}  // namespace barb200
