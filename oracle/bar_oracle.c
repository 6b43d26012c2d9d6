/*
 * bar_oracle.c -- TEST INFRASTRUCTURE ONLY (see poa_oracle.h). Plain-C restatement of the MSA-level part
 * of /root/reference/bar/impl/poaBarAligner.c: sliding-window POA with overlap trimming
 * (msa_make_partial_order_alignment, :463-749) and cross-end consistency trimming
 * (make_consistent_partial_order_alignments, :751-801; trim :376-434). The Flower/Cap plumbing
 * (:803-1299) is host glue around these and is not restated.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "poa_oracle.h"

#define GAP 5

/* nst_nt4_table, poaBarAligner.c:116-133: AaCcGgTt -> 0..3, '-' -> 5, anything else -> 4 */
static uint8_t to_byte(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

/* msa_to_byte(msa_to_base(b)) (poaBarAligner.c:136-161, 654-663): 0-3 stay, 5 and 27 are gaps, the rest N */
static uint8_t normalise(uint8_t b) { return b < 4 ? b : ((b == 5 || b == 27) ? GAP : 4); }

static const uint8_t rc_table[6] = {3, 2, 1, 0, 4, 5};      /* poaBarAligner.c:163 */

void oracle_msa_destruct(oracle_msa_t *m) { if (m) { free(m->seq_lens); free(m->msa); free(m); } }

#define AT(m, i, j) ((m)->msa[(size_t)(i) * (size_t)(m)->column_no + (size_t)(j)])

/* flip_msa_seq, poaBarAligner.c:302-317. NB column_no may have been reduced by msa_fix_trimmed while the
 * row stride stays, hence the explicit stride. */
typedef struct { oracle_msa_t m; int64_t stride; } wmsa_t;   /* a window: rows keep their original stride */
#define WAT(w, i, j) ((w)->m.msa[(size_t)(i) * (size_t)(w)->stride + (size_t)(j)])

static void flip(wmsa_t *w) {
    int64_t n = w->m.column_no, middle = n / 2;
    for (int64_t i = 0; i < w->m.seq_no; ++i) {
        for (int64_t j = 0; j < middle; ++j) {
            uint8_t buf = WAT(w, i, j);
            WAT(w, i, j) = rc_table[WAT(w, i, n - 1 - j)];
            WAT(w, i, n - 1 - j) = rc_table[buf];
        }
        if (n % 2 == 1) WAT(w, i, middle) = rc_table[WAT(w, i, middle)];
    }
}

/* make_column_scores, poaBarAligner.c:323-338 */
static float *column_scores(const wmsa_t *w) {
    float *s = (float *)calloc(w->m.column_no > 0 ? w->m.column_no : 1, sizeof(float));
    for (int64_t i = 0; i < w->m.column_no; ++i) {
        for (int64_t j = 0; j < w->m.seq_no; ++j) if (WAT(w, j, i) != GAP) s[i]++;
        if (s[i] >= 1.0) s[i]--;
    }
    return s;
}

/* sum_column_scores, poaBarAligner.c:344-354 */
static void cumulative(int64_t row, const wmsa_t *w, const float *cs, float *cu) {
    float c = 0.0; int64_t j = 0;
    for (int64_t i = 0; i < w->m.column_no; ++i) if (WAT(w, row, i) != GAP) { c += cs[i]; cu[j++] = c; }
    if (j != w->m.seq_lens[row]) { fprintf(stderr, "bar_oracle: seq_len mismatch in cumulative()\n"); exit(1); }
}

/* trim_msa_suffix, poaBarAligner.c:360-371 */
static void trim_suffix(wmsa_t *w, float *cs, int64_t row, int64_t suffix_start) {
    int64_t seq_index = 0;
    for (int64_t i = 0; i < w->m.column_no; ++i)
        if (WAT(w, row, i) != GAP && seq_index++ >= suffix_start) {
            WAT(w, row, i) = GAP;
            cs[i] = cs[i] > 1 ? cs[i] - 1 : 0;
        }
}

/* trim, poaBarAligner.c:376-434 */
static void trim(int64_t row1, wmsa_t *m1, float *cs1, int64_t row2, wmsa_t *m2, float *cs2, int64_t overlap) {
    if (overlap == 0) return;
    int64_t l1 = m1->m.seq_lens[row1], l2 = m2->m.seq_lens[row2];
    float *cu1 = (float *)malloc(sizeof(float) * (m1->m.column_no + 1)), *cu2 = (float *)malloc(sizeof(float) * (m2->m.column_no + 1));
    cumulative(row1, m1, cs1, cu1); cumulative(row2, m2, cs2, cu2);
    float max_cut = cu2[l2 - 1];
    if (overlap < l1) max_cut += cu1[l1 - overlap - 1];
    int64_t cut = 0;
    for (int64_t i = 0; i < overlap - 1; ++i) {
        float c = cu1[l1 - overlap + i] + cu2[l2 - i - 2];
        if (c > max_cut) { cut = i + 1; max_cut = c; }
    }
    float f = cu1[l1 - 1];
    if (overlap < l2) f += cu2[l2 - overlap - 1];
    if (f > max_cut) { max_cut = f; cut = overlap; }
    trim_suffix(m1, cs1, row1, l1 - overlap + cut);
    trim_suffix(m2, cs2, row2, l2 - cut);
    free(cu1); free(cu2);
}

/* msa_fix_trimmed, poaBarAligner.c:440-461 */
static void fix_trimmed(wmsa_t *w) {
    for (int64_t i = 0; i < w->m.seq_no; ++i) {
        w->m.seq_lens[i] = 0;
        for (int64_t j = 0; j < w->m.column_no; ++j) if (WAT(w, i, j) != GAP) ++w->m.seq_lens[i];
    }
    int64_t empty = 0; int still = 1;
    for (; empty < w->m.column_no; ++empty) {
        for (int64_t i = 0; i < w->m.seq_no && still; ++i) still = WAT(w, i, w->m.column_no - 1 - empty) == GAP;
        if (!still) break;
    }
    w->m.column_no -= empty;
}

static void wmsa_free(wmsa_t *w) { if (w) { free(w->m.seq_lens); free(w->m.msa); free(w); } }

/* msa_make_partial_order_alignment, poaBarAligner.c:463-749 */
oracle_msa_t *oracle_msa_make_partial_order_alignment(const oracle_params_t *p, char **seqs, const int *seq_lens,
        int64_t seq_no, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff) {
    oracle_msa_t *out = (oracle_msa_t *)calloc(1, sizeof(oracle_msa_t));
    out->seq_no = seq_no; out->seq_lens = (int *)malloc(sizeof(int) * seq_no);
    memcpy(out->seq_lens, seq_lens, sizeof(int) * seq_no);
    if (seq_no == 1) {                                         /* :471-483 */
        out->column_no = seq_lens[0];
        out->msa = (uint8_t *)malloc(out->column_no > 0 ? out->column_no : 1);
        for (int64_t i = 0; i < out->column_no; ++i) out->msa[i] = to_byte(seqs[0][i]);
        return out;
    }
    int64_t overlap_size = (int64_t)(0.5f * window_size);     /* :487-491 */
    if (overlap_size > 0) --overlap_size;
    int64_t bases_remaining = 0;
    int64_t *seq_offsets = (int64_t *)calloc(seq_no, sizeof(int64_t)), *row_overlaps = (int64_t *)calloc(seq_no, sizeof(int64_t));
    char *empty_seqs = (char *)calloc(seq_no, 1);
    for (int64_t i = 0; i < seq_no; ++i) bases_remaining += seq_lens[i];
    wmsa_t **windows = NULL; int64_t n_win = 0;
    wmsa_t *prev = NULL;
    while (bases_remaining > 0) {
        if (prev) {                                            /* :520-534 */
            for (int64_t i = 0; i < seq_no; ++i) {
                row_overlaps[i] = 0;
                for (int64_t j = prev->m.column_no - overlap_size; j < prev->m.column_no; ++j)
                    if (WAT(prev, i, j) != GAP) ++row_overlaps[i];
                seq_offsets[i] -= row_overlaps[i];
                bases_remaining += row_overlaps[i];
            }
        }
        wmsa_t *w = (wmsa_t *)calloc(1, sizeof(wmsa_t));
        w->m.seq_no = seq_no; w->m.seq_lens = (int *)malloc(sizeof(int) * seq_no);
        size_t tot = 0;
        for (int64_t i = 0; i < seq_no; ++i) {
            int64_t n = seq_lens[i] - seq_offsets[i]; if (n > window_size) n = window_size; if (n < 0) n = 0;
            w->m.seq_lens[i] = (int)n; tot += n > 0 ? n : 1;
        }
        uint8_t *flat = (uint8_t *)malloc(tot); int *lens = (int *)malloc(sizeof(int) * seq_no); size_t o = 0;
        int empty_count = 0;
        for (int64_t i = 0; i < seq_no; ++i) {                 /* :543-562 incl. the N hack for empty rows */
            if (w->m.seq_lens[i] == 0) { empty_seqs[i] = 1; w->m.seq_lens[i] = 1; flat[o] = to_byte('N'); ++empty_count; }
            else { empty_seqs[i] = 0; for (int j = 0; j < w->m.seq_lens[i]; ++j) flat[o + j] = to_byte(seqs[i][seq_offsets[i] + j]); }
            lens[i] = w->m.seq_lens[i]; o += lens[i];
        }
        oracle_params_t pp = *p;                               /* :567-571 */
        if (seq_no > max_prog_rows || (1. - (double)w->m.seq_lens[seq_no - 1] / (double)w->m.seq_lens[0] > max_prog_length_diff))
            pp.progressive_poa = 0;
        uint8_t *msa = NULL;
        int msa_len = oracle_poa_msa(&pp, (int)seq_no, lens, flat, &msa);
        free(flat); free(lens);
        w->m.msa = msa; w->m.column_no = msa_len; w->stride = msa_len;
        for (int64_t i = 0; i < seq_no && empty_count > 0; ++i) {   /* :631-644 */
            if (empty_seqs[i]) for (int64_t j = 0; j < w->m.column_no; ++j) if (normalise(WAT(w, i, j)) != GAP) {
                WAT(w, i, j) = GAP; --w->m.seq_lens[i]; --empty_count; break;
            }
        }
        for (int64_t i = 0; i < seq_no; ++i) {                 /* :654-663 */
            for (int64_t j = 0; j < w->m.column_no; ++j) WAT(w, i, j) = normalise(WAT(w, i, j));
            bases_remaining -= w->m.seq_lens[i];
            seq_offsets[i] += w->m.seq_lens[i];
        }
        if (prev) {                                            /* :668-689 */
            flip(w);
            float *pcs = column_scores(prev), *cs = column_scores(w);
            for (int64_t i = 0; i < seq_no; ++i) {
                int64_t ov = w->m.seq_lens[i] < row_overlaps[i] ? w->m.seq_lens[i] : row_overlaps[i];
                if (ov > 0) trim(i, w, cs, i, prev, pcs, ov);
            }
            fix_trimmed(w); fix_trimmed(prev);
            flip(w);
            free(pcs); free(cs);
        }
        windows = (wmsa_t **)realloc(windows, sizeof(wmsa_t *) * (n_win + 1));
        windows[n_win++] = w; prev = w;
    }
    out->column_no = 0;
    for (int64_t k = 0; k < n_win; ++k) out->column_no += windows[k]->m.column_no;
    out->msa = (uint8_t *)malloc((size_t)seq_no * (out->column_no > 0 ? out->column_no : 1));
    for (int64_t i = 0; i < seq_no; ++i) {                     /* :703-736 */
        int64_t o = 0;
        for (int64_t k = 0; k < n_win; ++k) for (int64_t c = 0; c < windows[k]->m.column_no; ++c) AT(out, i, o++) = WAT(windows[k], i, c);
    }
    for (int64_t k = 0; k < n_win; ++k) wmsa_free(windows[k]);
    free(windows); free(seq_offsets); free(row_overlaps); free(empty_seqs);
    return out;
}

/* make_consistent_partial_order_alignments, poaBarAligner.c:751-801 */
oracle_msa_t **oracle_make_consistent_partial_order_alignments(const oracle_params_t *p, int64_t end_no,
        const int64_t *end_lengths, char ***end_strings, int **end_string_lengths, int64_t **right_end_indexes,
        int64_t **right_end_row_indexes, int64_t **overlaps, int64_t window_size, int64_t max_prog_rows,
        double max_prog_length_diff) {
    oracle_msa_t **msas = (oracle_msa_t **)malloc(sizeof(oracle_msa_t *) * end_no);
    wmsa_t *ws = (wmsa_t *)calloc(end_no, sizeof(wmsa_t));
    float **cs = (float **)malloc(sizeof(float *) * end_no);
    for (int64_t i = 0; i < end_no; ++i) {
        msas[i] = oracle_msa_make_partial_order_alignment(p, end_strings[i], end_string_lengths[i], end_lengths[i],
                                                          window_size, max_prog_rows, max_prog_length_diff);
        ws[i].m = *msas[i]; ws[i].stride = msas[i]->column_no;
        cs[i] = column_scores(&ws[i]);
    }
    for (int64_t i = 0; i < end_no; ++i)
        for (int64_t j = 0; j < ws[i].m.seq_no; ++j) {
            int64_t re = right_end_indexes[i][j], rr = right_end_row_indexes[i][j];
            if (re > i || (re == i && rr > j)) trim(j, &ws[i], cs[i], rr, &ws[re], cs[re], overlaps[i][j]);
        }
    for (int64_t i = 0; i < end_no; ++i) free(cs[i]);
    free(cs); free(ws);
    return msas;
}
