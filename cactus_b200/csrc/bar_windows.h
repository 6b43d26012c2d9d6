// bar_windows.h -- the MSA-level host logic of Cactus' POA shim, restated on a flat window type and cut into the steps the
// end queue (end_queue.h) needs: no CUDA, no context -- tests/hosttest drives the same code on the CPU with a stand-in device.
//
// What stays on the host (cheap, serial, O(MSA size)) and why:
//   * slicing every end into sliding windows of `window_size` bases with 50 % overlap, the N stand-in for rows that
//     ran out of bases, the per-window progressive-mode switch           (bar/impl/poaBarAligner.c:485-571)
//   * trimming consecutive windows against each other and stitching them (:668-736; trim :376-434)
//   * trimming the MSAs of two ends that share a string                   (:751-801)
// What changes: the reference aligns window after window, end after end (one abpoa_msa each, :609). Here the next windows of
// ALL unfinished ends in the queue form one device batch (window n+1 of an end needs window n's trimmed MSA, so an end's
// rounds are sequential, the ends within a round are not).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <string>
#include <vector>
#include "host_api.h"

namespace barb200 {
namespace barwin {

constexpr uint8_t GAPB = 5;
static const uint8_t kRc[6] = {3, 2, 1, 0, 4, 5};                  // complement in the POA alphabet (poaBarAligner.c:163)

inline uint8_t ascii_to_code(char c) {                      // nst_nt4_table (poaBarAligner.c:116-133)
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

// A window MSA: rows keep their original stride when trailing empty columns are clipped.
struct Window {
    int64_t seq_no = 0, column_no = 0, stride = 0;
    std::vector<int> seq_lens;
    std::vector<uint8_t> m;
    uint8_t &at(int64_t i, int64_t j) { return m[(size_t)(i * stride + j)]; }
    uint8_t at(int64_t i, int64_t j) const { return m[(size_t)(i * stride + j)]; }
};

inline void flip(Window &w) {                                       // flip_msa_seq (:302-317)
    const int64_t n = w.column_no, mid = n / 2;
    for (int64_t i = 0; i < w.seq_no; ++i) {
        for (int64_t j = 0; j < mid; ++j) {
            const uint8_t a = w.at(i, j);
            w.at(i, j) = kRc[w.at(i, n - 1 - j)];
            w.at(i, n - 1 - j) = kRc[a];
        }
        if (n & 1) w.at(i, mid) = kRc[w.at(i, mid)];
    }
}

inline std::vector<float> column_scores(const Window &w) {          // make_column_scores (:323-338): max(#bases - 1, 0)
    std::vector<float> s((size_t)w.column_no, 0.f);
    for (int64_t c = 0; c < w.column_no; ++c) {
        for (int64_t r = 0; r < w.seq_no; ++r) if (w.at(r, c) != GAPB) s[c]++;
        if (s[c] >= 1.0f) s[c]--;
    }
    return s;
}

inline std::vector<float> cumulative(int64_t row, const Window &w, const std::vector<float> &cs) {   // sum_column_scores (:344-354)
    std::vector<float> cu; cu.reserve((size_t)w.seq_lens[row] + 1);
    float acc = 0.f;
    for (int64_t c = 0; c < w.column_no; ++c) if (w.at(row, c) != GAPB) { acc += cs[c]; cu.push_back(acc); }
    return cu;
}

inline void trim_suffix(Window &w, std::vector<float> &cs, int64_t row, int64_t start) {             // trim_msa_suffix (:360-371)
    int64_t k = 0;
    for (int64_t c = 0; c < w.column_no; ++c)
        if (w.at(row, c) != GAPB && k++ >= start) { w.at(row, c) = GAPB; cs[c] = cs[c] > 1 ? cs[c] - 1 : 0; }
}

// trim (:376-434): the cut inside the shared overlap that maximises the summed column scores kept on both sides
inline bool trim(int64_t r1, Window &m1, std::vector<float> &cs1, int64_t r2, Window &m2, std::vector<float> &cs2, int64_t overlap) {
    if (overlap == 0) return true;
    const int64_t l1 = m1.seq_lens[r1], l2 = m2.seq_lens[r2];
    if (overlap < 0 || overlap > l1 || overlap > l2) return false;
    const std::vector<float> cu1 = cumulative(r1, m1, cs1), cu2 = cumulative(r2, m2, cs2);
    if ((int64_t)cu1.size() != l1 || (int64_t)cu2.size() != l2) return false;
    float best = cu2[l2 - 1];
    if (overlap < l1) best += cu1[l1 - overlap - 1];
    int64_t cut = 0;
    for (int64_t i = 0; i < overlap - 1; ++i) {
        const float c = cu1[l1 - overlap + i] + cu2[l2 - i - 2];
        if (c > best) { cut = i + 1; best = c; }
    }
    float f = cu1[l1 - 1];
    if (overlap < l2) f += cu2[l2 - overlap - 1];
    if (f > best) { best = f; cut = overlap; }
    trim_suffix(m1, cs1, r1, l1 - overlap + cut);
    trim_suffix(m2, cs2, r2, l2 - cut);
    return true;
}

inline void fix_trimmed(Window &w) {                                // msa_fix_trimmed (:440-461)
    for (int64_t r = 0; r < w.seq_no; ++r) {
        int n = 0;
        for (int64_t c = 0; c < w.column_no; ++c) if (w.at(r, c) != GAPB) ++n;
        w.seq_lens[r] = n;
    }
    int64_t empty = 0; bool still = true;
    for (; empty < w.column_no; ++empty) {
        for (int64_t r = 0; r < w.seq_no && still; ++r) still = w.at(r, w.column_no - 1 - empty) == GAPB;
        if (!still) break;
    }
    w.column_no -= empty;
}

// the state of one end while its windows are being aligned (locals of msa_make_partial_order_alignment, :485-516)
struct EndState {
    int64_t seq_no = 0;
    std::vector<std::vector<uint8_t>> codes;                // inputs converted once
    std::vector<int> seq_lens;
    std::vector<int64_t> offsets, row_overlaps;
    std::vector<char> empty;
    int64_t bases_remaining = 0;
    std::vector<std::unique_ptr<Window>> windows;
    // the window in flight
    std::vector<int> cur_lens; std::vector<uint8_t> cur_flat; int cur_progressive = 1;
    bool done = false;
};

// :471-483, 494-516. Returns an error text or "".
inline std::string end_init(EndState &E, int64_t seq_no, char **seqs, const int *seq_lens) {
    E.seq_no = seq_no;
    if (E.seq_no <= 0) return "end without sequences";         // the reference asserts seq_no > 0 (poaBarAligner.c:466)
    E.seq_lens.assign(seq_lens, seq_lens + E.seq_no);
    E.codes.resize(E.seq_no);
    for (int64_t i = 0; i < E.seq_no; ++i) {
        if (E.seq_lens[i] < 0) return "negative sequence length";
        E.codes[i].resize(E.seq_lens[i]);
        for (int t = 0; t < E.seq_lens[i]; ++t) E.codes[i][t] = ascii_to_code(seqs[i][t]);
        E.bases_remaining += E.seq_lens[i];
    }
    E.offsets.assign(E.seq_no, 0); E.row_overlaps.assign(E.seq_no, 0); E.empty.assign(E.seq_no, 0);
    // a single string is its own alignment (:471-483); nothing to align when no bases are left
    E.done = E.seq_no == 1 || E.bases_remaining == 0;
    return "";
}

inline int64_t overlap_of_window(int64_t window_size) {        // :487-491
    int64_t overlap_size = (int64_t)(0.5f * window_size);
    if (overlap_size > 0) --overlap_size;
    return overlap_size;
}

// the next window of an unfinished end (:520-571) as a device job (views into E). Returns an error text or "".
inline std::string end_prepare_window(EndState &E, int64_t window_size, int default_progressive, int64_t max_prog_rows,
                                      double max_prog_length_diff, HostJob &job) {
    const int64_t overlap_size = overlap_of_window(window_size);
    if (!E.windows.empty()) {
        const Window &prev = *E.windows.back();
        if (prev.column_no <= overlap_size) return "window shorter than the window overlap (reference asserts, poaBarAligner.c:522)";
        for (int64_t i = 0; i < E.seq_no; ++i) {
            int64_t ov = 0;
            for (int64_t c = prev.column_no - overlap_size; c < prev.column_no; ++c) if (prev.at(i, c) != GAPB) ++ov;
            E.row_overlaps[i] = ov; E.offsets[i] -= ov; E.bases_remaining += ov;
        }
    }
    E.cur_lens.assign(E.seq_no, 0); E.cur_flat.clear();
    for (int64_t i = 0; i < E.seq_no; ++i) {
        int64_t n = std::min<int64_t>(window_size, E.seq_lens[i] - E.offsets[i]);
        if (n <= 0) { E.empty[i] = 1; E.cur_lens[i] = 1; E.cur_flat.push_back(4); }              // the N stand-in, :551-562
        else { E.empty[i] = 0; E.cur_lens[i] = (int)n; E.cur_flat.insert(E.cur_flat.end(), E.codes[i].begin() + E.offsets[i], E.codes[i].begin() + E.offsets[i] + n); }
    }
    E.cur_progressive = default_progressive;
    if (E.seq_no > max_prog_rows || (1. - (double)E.cur_lens[E.seq_no - 1] / (double)E.cur_lens[0] > max_prog_length_diff)) E.cur_progressive = 0;   // :567-571
    job = HostJob{(int)E.seq_no, E.cur_lens.data(), E.cur_flat.data(), E.cur_progressive};
    return "";
}

// take the window's MSA back, trim it against the previous window (:612-700). Returns an error text or "".
inline std::string end_consume_window(EndState &E, JobResult &res) {
    std::unique_ptr<Window> w(new Window());
    w->seq_no = E.seq_no; w->column_no = w->stride = res.msa_len; w->seq_lens = E.cur_lens; w->m.swap(res.msa);
    for (int64_t i = 0; i < E.seq_no; ++i) if (E.empty[i])
        for (int64_t c = 0; c < w->column_no; ++c) if (w->at(i, c) != GAPB) { w->at(i, c) = GAPB; --w->seq_lens[i]; break; }   // :631-644
    for (int64_t i = 0; i < E.seq_no; ++i) { E.bases_remaining -= w->seq_lens[i]; E.offsets[i] += w->seq_lens[i]; }
    bool ok = true;
    if (!E.windows.empty()) {
        Window &prev = *E.windows.back();
        flip(*w);
        std::vector<float> pcs = column_scores(prev), cs = column_scores(*w);
        for (int64_t i = 0; i < E.seq_no; ++i) {
            const int64_t ov = std::min<int64_t>(w->seq_lens[i], E.row_overlaps[i]);
            if (ov > 0 && !trim(i, *w, cs, i, prev, pcs, ov)) ok = false;
        }
        fix_trimmed(*w); fix_trimmed(prev);
        flip(*w);
    }
    E.windows.push_back(std::move(w));
    if (E.bases_remaining <= 0) E.done = true;
    return ok ? "" : "inconsistent overlap while trimming windows";
}

inline barb200_msa *to_msa(int64_t seq_no, const std::vector<int> &lens, int64_t cols) {
    barb200_msa *m = (barb200_msa *)calloc(1, sizeof(barb200_msa));
    if (!m) return nullptr;
    m->seq_no = seq_no; m->column_no = cols;
    m->seq_lens = (int *)malloc(sizeof(int) * (size_t)(seq_no > 0 ? seq_no : 1));
    m->msa = (uint8_t *)malloc((size_t)(seq_no * cols > 0 ? seq_no * cols : 1));
    if (!m->seq_lens || !m->msa) { free(m->seq_lens); free(m->msa); free(m); return nullptr; }
    for (int64_t i = 0; i < seq_no; ++i) m->seq_lens[i] = lens[i];
    return m;
}

// stitch the windows of a finished end (:703-736); nullptr = out of memory
inline barb200_msa *end_stitch(const EndState &E) {
    int64_t cols = 0;
    if (E.seq_no == 1) cols = E.seq_lens[0]; else for (auto &w : E.windows) cols += w->column_no;
    barb200_msa *m = to_msa(E.seq_no, E.seq_lens, cols);
    if (!m) return nullptr;
    if (E.seq_no == 1) memcpy(m->msa, E.codes[0].data(), (size_t)cols);
    else for (int64_t i = 0; i < E.seq_no; ++i) {
        int64_t o = 0;
        for (auto &w : E.windows) { memcpy(m->msa + (size_t)(i * cols + o), &w->m[(size_t)(i * w->stride)], (size_t)w->column_no); o += w->column_no; }
    }
    return m;
}

// cross-end consistency (:781-793), serial exactly as in the reference: the order of trims matters. false = bad indexes / overlaps
inline bool consistency_trim(int64_t end_no, barb200_msa **msas, const std::vector<std::vector<int64_t>> &right_end_indexes,
                             const std::vector<std::vector<int64_t>> &right_end_row_indexes, const std::vector<std::vector<int64_t>> &overlaps) {
    std::vector<Window> ws((size_t)end_no); std::vector<std::vector<float>> cs((size_t)end_no);
    for (int64_t i = 0; i < end_no; ++i) {
        Window &w = ws[i];
        w.seq_no = msas[i]->seq_no; w.column_no = w.stride = msas[i]->column_no;
        w.seq_lens.assign(msas[i]->seq_lens, msas[i]->seq_lens + w.seq_no);
        w.m.assign(msas[i]->msa, msas[i]->msa + (size_t)(w.seq_no * w.column_no));
        cs[i] = column_scores(w);
    }
    for (int64_t i = 0; i < end_no; ++i)
        for (int64_t j = 0; j < ws[i].seq_no; ++j) {
            const int64_t re = right_end_indexes[i][j], rr = right_end_row_indexes[i][j];
            if (re > i || (re == i && rr > j)) {
                if (re < 0 || re >= end_no || rr < 0 || rr >= ws[re].seq_no) return false;
                if (!trim(j, ws[i], cs[i], rr, ws[re], cs[re], overlaps[i][j])) return false;
            }
        }
    for (int64_t i = 0; i < end_no; ++i) memcpy(msas[i]->msa, ws[i].m.data(), ws[i].m.size());
    return true;
}

}  // namespace barwin
}  // namespace barb200
