// guide_tree.cuh -- the progressive-POA read order (which sequence is aligned when) computed by the CTA that owns the job.
//
// With seeding disabled Cactus still lets abPOA pick the alignment order from a guide tree
// (abpoa_build_guide_tree_partition, abPOA src/abpoa_seed.c:705-722): (w,k)-minimizers of every sequence (mm_sketch,
// :85-156; forward strand only, no homopolymer compression), pairwise min-count Jaccard similarity of the minimizer
// multisets and a greedy order (abpoa_build_guide_tree, :232-325). Round 1 computed this on host threads and streamed the
// orders in behind the running kernel; here it is the first phase of the job on the device, so the host does
// no per-job work at all and K in the hundreds costs no host time:
//   1. sketch: one THREAD per sequence runs the reference's serial window scan (ties and duplicates exactly as mm_sketch
//      reports them -- the multiset matters) and appends keys (hash << 8 | span) << 16 | read to one array;
//   2. the keys are sorted by the CTA (bitonic network in global memory; only the grouping by hash matters);
//   3. every group of equal hashes adds min(count_a, count_b) to the pair counters (integer atomics: order free);
//   4. Jaccard = shared / total in double (one IEEE division of two integers, the same value as the reference's), first
//      maximum in the reference's loop order; then the greedy order: a read's score is the sum of its similarities to the
//      reads placed so far, accumulated in placement order exactly like the reference's inner loop, first maximum wins.
// The same source compiles for the host (tests/hosttest), where the T threads of every phase run one after the other.
#pragma once
#include <stdint.h>
#include "poa_types.h"

#if defined(__CUDA_ARCH__)
#define GT_THREADS(tid, T) for (int tid = (int)threadIdx.x, _gt_once = 1; _gt_once; _gt_once = 0)
#define GT_SYNC() __syncthreads()
#else
#define GT_THREADS(tid, T) for (int tid = 0; tid < (T); ++tid)
#define GT_SYNC() ((void)0)
#endif

namespace barb200 {

struct GuideTreeParams { int k, w; };        // partialOrderAlignmentMinimizerK / W

// per-slot scratch of the guide tree (carved from the slot by the kernel)
struct GtScratch {
    uint64_t *keys; int key_cap;             // minimizer keys, capacity a power of two
    int *hit;                                // [K (K + 1) / 2] pair counters, tri(i, j) = i (i + 1) / 2 + j, i >= j
    double *jac;                             // [K (K - 1) / 2] similarities, jidx(i, j) = i (i - 1) / 2 + j, i > j
    double *score;                           // [K] running greedy scores (-1 once placed)
    int *n_keys;                             // 1 int: keys appended (shared or global)
    int *red_i; double *red_v;               // [T / 32 + 1] block reductions
};

HD uint64_t gt_hash64(uint64_t key, uint64_t mask) {   // minimap2's hash64, abpoa_seed.c:36-46
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

HD void gt_push(const GtScratch &G, uint64_t x, uint32_t rid) {
#if defined(__CUDA_ARCH__)
    const int pos = atomicAdd(G.n_keys, 1);
#else
    const int pos = (*G.n_keys)++;
#endif
    if (pos < G.key_cap) G.keys[pos] = x << 16 | (uint64_t)rid;
}

// Windowed minimizer sampling of ONE sequence with the reference's treatment of ties (abpoa_seed.c:85-156): when several k-mers of
// a window share the minimal hash all of them are reported. Serial; x = hash << 8 | span, y = position << 1 (only compared).
HD void gt_sketch(const GtScratch &G, const uint8_t *s, int len, int w, int k, uint32_t rid) {
    const uint64_t mask = (1ULL << 2 * k) - 1, NONE = ~0ULL;
    uint64_t rx[256], ry[256];               // the window (w < 256, checked at create)
    for (int j = 0; j < w; ++j) { rx[j] = NONE; ry[j] = NONE; }
    uint64_t bx = NONE, by = NONE, kmer = 0;
    int run = 0, pos = 0, best_pos = 0;      // run: bases since the last ambiguous one
    for (int i = 0; i < len; ++i) {
        const int c = s[i];
        uint64_t cx = NONE, cy = NONE;
        if (c < 4) {
            const int span = run + 1 < k ? run + 1 : k;
            kmer = (kmer << 2 | (uint64_t)c) & mask;
            ++run;
            if (run >= k) { cx = gt_hash64(kmer, mask) << 8 | (uint64_t)span; cy = (uint64_t)(uint32_t)i << 1; }
        } else run = 0;
        rx[pos] = cx; ry[pos] = cy;
        if (run == w + k - 1 && bx != NONE) {
            for (int j = pos + 1; j < w; ++j) if (bx == rx[j] && by != ry[j]) gt_push(G, rx[j], rid);
            for (int j = 0; j < pos; ++j) if (bx == rx[j] && by != ry[j]) gt_push(G, rx[j], rid);
        }
        if (cx <= bx) {
            if (run >= w + k && bx != NONE) gt_push(G, bx, rid);
            bx = cx; by = cy; best_pos = pos;
        } else if (pos == best_pos) {
            if (run >= w + k - 1 && bx != NONE) gt_push(G, bx, rid);
            bx = NONE;
            for (int j = pos + 1; j < w; ++j) if (bx >= rx[j]) { bx = rx[j]; by = ry[j]; best_pos = j; }
            for (int j = 0; j <= pos; ++j) if (bx >= rx[j]) { bx = rx[j]; by = ry[j]; best_pos = j; }
            if (run >= w + k - 1 && bx != NONE) {
                for (int j = pos + 1; j < w; ++j) if (bx == rx[j] && by != ry[j]) gt_push(G, rx[j], rid);
                for (int j = 0; j < pos + 1; ++j) if (bx == rx[j] && by != ry[j]) gt_push(G, rx[j], rid);
            }
        }
        if (++pos == w) pos = 0;
    }
    if (bx != NONE) gt_push(G, bx, rid);
}

HD int64_t gt_tri(int64_t i, int64_t j) { return i * (i + 1) / 2 + j; }     // i >= j
HD int64_t gt_jidx(int64_t i, int64_t j) { return i * (i - 1) / 2 + j; }    // i > j

// first maximum of (value, index): larger value wins, equal values -> smaller index (the reference's `if (v > best)` scans)
HD bool gt_better(double v, int64_t i, double bv, int64_t bi) { return v > bv || (v == bv && i < bi); }

// Block-wide first-maximum of per-thread candidates. All threads; result in red_v[0] / red_i64 (returned to every thread).
// T must be a multiple of 32 on the device. ws_v / ws_i: >= 33 entries.
HD void gt_block_argmax(double &v, int64_t &i, double *ws_v, long long *ws_i, int T) {
#if defined(__CUDA_ARCH__)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = T >> 5;
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, off); const long long oi = __shfl_down_sync(0xffffffffu, (long long)i, off);
        if (gt_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    __syncthreads();
    if (lane == 0) { ws_v[warp] = v; ws_i[warp] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = ws_v[0]; long long bi = ws_i[0];
        for (int q = 1; q < nw; ++q) if (gt_better(ws_v[q], ws_i[q], bv, bi)) { bv = ws_v[q]; bi = ws_i[q]; }
        ws_v[32] = bv; ws_i[32] = bi;
    }
    __syncthreads();
    v = ws_v[32]; i = ws_i[32];
#else
    (void)ws_v; (void)ws_i; (void)T; (void)v; (void)i;      // the host emulation reduces in the caller (threads run one after the other)
#endif
}

// The read order of one job into order[0 .. n). seq(i) / len(i) give sequence i. Returns 0, or -1 if the key array was too small.
// All threads of the CTA (T of them); ws_v / ws_i: shared scratch of >= 33 doubles / long longs.
template <class SeqFn>
HD int cta_guide_tree(const GuideTreeParams &P, int progressive, int n, SeqFn seq, const int *lens, int *order, const GtScratch &G,
                      double *ws_v, long long *ws_i, int T) {
    GT_THREADS(tid, T) { for (int i = tid; i < n; i += T) order[i] = i; if (tid == 0) *G.n_keys = 0; }
    GT_SYNC();
    if (!(progressive && n > 2)) return 0;
    // ---- 1. sketches: one thread per sequence ----
    GT_THREADS(tid, T) { for (int i = tid; i < n; i += T) gt_sketch(G, seq(i), lens[i], P.w, P.k, (uint32_t)i); }
    GT_SYNC();
    const int nk = *G.n_keys;
    if (nk > G.key_cap) return -1;
    if (nk == 0) return 0;
    // ---- 2. sort (bitonic, padded to a power of two with the maximal key) ----
    int n2 = 1;
    while (n2 < nk) n2 <<= 1;
    GT_THREADS(tid, T) { for (int i = nk + tid; i < n2; i += T) G.keys[i] = ~0ULL; }
    GT_SYNC();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            GT_THREADS(tid, T) {
                for (int t = tid; t < (n2 >> 1); t += T) {
                    const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const uint64_t a = G.keys[lo], b = G.keys[hi];
                    if ((a > b) == up) { G.keys[lo] = b; G.keys[hi] = a; }
                }
            }
            GT_SYNC();
        }
    }
    // ---- 3. pair counters: every group of equal hashes (abpoa_seed.c:232-284) ----
    const int64_t n_tri = (int64_t)n * (n + 1) / 2;
    GT_THREADS(tid, T) { for (int64_t i = tid; i < n_tri; i += T) G.hit[i] = 0; }
    GT_SYNC();
    GT_THREADS(tid, T) {
        for (int s = tid; s < nk; s += T) {
            const uint64_t x = G.keys[s] >> 16;
            if (s > 0 && (G.keys[s - 1] >> 16) == x) continue;             // not the first entry of its group
            // the group is sorted by read: runs of equal reads are the per-read counts
            int e = s;
            while (e < nk && (G.keys[e] >> 16) == x) ++e;
            for (int a = s; a < e;) {
                const int ra = (int)(G.keys[a] & 0xffff);
                int ae = a; while (ae < e && (int)(G.keys[ae] & 0xffff) == ra) ++ae;
                const int ca = ae - a;
#if defined(__CUDA_ARCH__)
                atomicAdd(&G.hit[gt_tri(ra, ra)], ca);
#else
                G.hit[gt_tri(ra, ra)] += ca;
#endif
                for (int b = ae; b < e;) {
                    const int rb = (int)(G.keys[b] & 0xffff);
                    int be = b; while (be < e && (int)(G.keys[be] & 0xffff) == rb) ++be;
                    const int cb = be - b, m = ca < cb ? ca : cb;           // rb > ra (sorted)
#if defined(__CUDA_ARCH__)
                    atomicAdd(&G.hit[gt_tri(rb, ra)], m);
#else
                    G.hit[gt_tri(rb, ra)] += m;
#endif
                    b = be;
                }
                a = ae;
            }
        }
    }
    GT_SYNC();
    // ---- 4. similarities + the most similar pair (first maximum in the order i ascending, j ascending) ----
    const int64_t n_pairs = (int64_t)n * (n - 1) / 2;
    double bv = -1.0; int64_t bi = INT64_MAX;
    GT_THREADS(tid, T) {
        double tv = -1.0; int64_t ti = INT64_MAX;
        for (int64_t q = tid; q < n_pairs; q += T) {
            // q = jidx(i, j): i = the row with i (i - 1) / 2 <= q
            int64_t i = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)q)) / 2.0);
            while (i * (i - 1) / 2 > q) --i;
            while ((i + 1) * i / 2 <= q) ++i;
            const int64_t j = q - i * (i - 1) / 2;
            const int shared = G.hit[gt_tri(i, j)], tot = G.hit[gt_tri(i, i)] + G.hit[gt_tri(j, j)] - shared;
            const double v = tot == 0 ? 0.0 : (0.0 + shared) / tot;
            G.jac[q] = v;
            if (gt_better(v, q, tv, ti)) { tv = v; ti = q; }
        }
#if defined(__CUDA_ARCH__)
        gt_block_argmax(tv, ti, ws_v, ws_i, T);
        bv = tv; bi = ti;
#else
        if (gt_better(tv, ti, bv, bi)) { bv = tv; bi = ti; }
#endif
    }
    GT_SYNC();
    int64_t pi = (int64_t)((1.0 + sqrt(1.0 + 8.0 * (double)bi)) / 2.0);
    while (pi * (pi - 1) / 2 > bi) --pi;
    while ((pi + 1) * pi / 2 <= bi) ++pi;
    const int64_t pj = bi - pi * (pi - 1) / 2;
    // ---- 5. greedy order (abpoa_seed.c:286-325): score(r) = sum over the placed reads, in placement order ----
    GT_THREADS(tid, T) {
        for (int r = tid; r < n; r += T) G.score[r] = (r == pi || r == pj) ? -1.0 : 0.0;
        if (tid == 0) { order[0] = (int)pj; order[1] = (int)pi; }
    }
    GT_SYNC();
    int last[2] = {(int)pj, (int)pi}, n_last = 2;
    for (int placed = 2; placed < n; ++placed) {
        double sv = -1.0; int64_t si = INT64_MAX;
        GT_THREADS(tid, T) {
            double tv = -2.0; int64_t ti = INT64_MAX;
            for (int r = tid; r < n; r += T) {
                double v = G.score[r];
                if (v >= 0.0) {
                    for (int q = 0; q < n_last; ++q) { const int o = last[q]; v += r > o ? G.jac[gt_jidx(r, o)] : G.jac[gt_jidx(o, r)]; }
                    G.score[r] = v;
                    if (gt_better(v, r, tv, ti)) { tv = v; ti = r; }
                }
            }
#if defined(__CUDA_ARCH__)
            gt_block_argmax(tv, ti, ws_v, ws_i, T);
            sv = tv; si = ti;
#else
            if (gt_better(tv, ti, sv, si)) { sv = tv; si = ti; }
#endif
        }
        GT_SYNC();
        GT_THREADS(tid, T) { if (tid == 0) { order[placed] = (int)si; G.score[si] = -1.0; } }
        GT_SYNC();
        last[0] = (int)si; n_last = 1;
    }
    return 0;
}

}  // namespace barb200
