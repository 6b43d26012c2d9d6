/*
 * bar_ref_harness.c -- TEST INFRASTRUCTURE ONLY. Flat (pointer + size) wrappers around the UNMODIFIED
 * reference msa_make_partial_order_alignment / make_consistent_partial_order_alignments
 * (/root/reference/bar/impl/poaBarAligner.c:463-801, compiled by oracle/Makefile from where it lies),
 * so tests can call them through ctypes and compare with oracle/bar_oracle.c and with the CUDA product.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "poaBarAligner.h"

typedef struct {
    int wb; float wf;
    int gap_open1, gap_ext1, gap_open2, gap_ext2;
    int mat[25];
    int k, w, min_w;
    int progressive_poa, disable_seeding;
} ref_params_t;

/* same values abpoaParamaters_constructFromCactusParams (poaBarAligner.c:24-81) builds from the XML */
static abpoa_para_t *make_para(const ref_params_t *p) {
    abpoa_para_t *abpt = abpoa_init_para();
    abpt->out_msa = 1; abpt->out_cons = 0;
    abpt->align_mode = ABPOA_GLOBAL_MODE;
    abpt->wb = p->wb; abpt->wf = p->wf;
    abpt->gap_open1 = p->gap_open1; abpt->gap_ext1 = p->gap_ext1;
    abpt->gap_open2 = p->gap_open2; abpt->gap_ext2 = p->gap_ext2;
    abpt->disable_seeding = p->disable_seeding;
    abpt->k = p->k; abpt->w = p->w; abpt->min_w = p->min_w;
    abpt->progressive_poa = p->progressive_poa;
    abpt->use_score_matrix = 0;
    abpoa_post_set_para(abpt);
    abpt->use_score_matrix = 1;
    memcpy(abpt->mat, p->mat, 25 * sizeof(int));
    abpt->min_mis = 0; abpt->max_mat = 0;
    for (int i = 0; i < 25; ++i) {
        if (abpt->mat[i] > abpt->max_mat) abpt->max_mat = abpt->mat[i];
        if (-abpt->mat[i] > abpt->min_mis) abpt->min_mis = -abpt->mat[i];
    }
    return abpt;
}

static char **dup_strings(char **s, const int *lens, int64_t n) {
    char **out = (char **)malloc(sizeof(char *) * n);
    for (int64_t i = 0; i < n; ++i) { out[i] = (char *)malloc(lens[i] + 1); memcpy(out[i], s[i], lens[i]); out[i][lens[i]] = 0; }
    return out;
}
static int *dup_ints(const int *a, int64_t n) { int *o = (int *)malloc(sizeof(int) * n); memcpy(o, a, sizeof(int) * n); return o; }

static uint8_t *flatten(Msa *m) {
    uint8_t *o = (uint8_t *)malloc((size_t)m->seq_no * (m->column_no > 0 ? m->column_no : 1));
    for (int64_t i = 0; i < m->seq_no; ++i) memcpy(o + (size_t)i * m->column_no, m->msa_seq[i], m->column_no);
    return o;
}

/* returns column_no; *msa_out = malloc'd seq_no*column_no bytes */
int64_t bar_ref_msa_make_partial_order_alignment(const ref_params_t *p, char **seqs, const int *seq_lens, int64_t seq_no,
        int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff, uint8_t **msa_out) {
    abpoa_para_t *abpt = make_para(p);
    Msa *m = msa_make_partial_order_alignment(dup_strings(seqs, seq_lens, seq_no), dup_ints(seq_lens, seq_no), seq_no,
                                              window_size, max_prog_rows, max_prog_length_diff, abpt);
    int64_t cols = m->column_no;
    *msa_out = flatten(m);
    msa_destruct(m);
    abpoa_free_para(abpt);
    return cols;
}

/* column_nos[end_no] and msa_outs[end_no] are filled */
void bar_ref_make_consistent_partial_order_alignments(const ref_params_t *p, int64_t end_no, int64_t *end_lengths,
        char ***end_strings, int **end_string_lengths, int64_t **right_end_indexes, int64_t **right_end_row_indexes,
        int64_t **overlaps, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff,
        int64_t *column_nos, uint8_t **msa_outs) {
    abpoa_para_t *abpt = make_para(p);
    char ***es = (char ***)malloc(sizeof(char **) * end_no);
    int **el = (int **)malloc(sizeof(int *) * end_no);
    for (int64_t i = 0; i < end_no; ++i) { es[i] = dup_strings(end_strings[i], end_string_lengths[i], end_lengths[i]); el[i] = dup_ints(end_string_lengths[i], end_lengths[i]); }
    Msa **msas = make_consistent_partial_order_alignments(end_no, end_lengths, es, el, right_end_indexes, right_end_row_indexes,
                                                          overlaps, window_size, max_prog_rows, max_prog_length_diff, abpt);
    for (int64_t i = 0; i < end_no; ++i) { column_nos[i] = msas[i]->column_no; msa_outs[i] = flatten(msas[i]); msa_destruct(msas[i]); }
    free(msas); free(es); free(el);
    abpoa_free_para(abpt);
}

void bar_ref_free(void *p) { free(p); }
