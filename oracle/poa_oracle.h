/*
 * poa_oracle.h -- TEST INFRASTRUCTURE ONLY. Plain-C restatement of the reference BAR/POA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load libpoa_oracle.so. The product (cactus_b200/) never links or calls anything declared here.
 *
 * Parity status: PINNED. Checked bit-for-bit against the unmodified reference abPOA v1.5.6 (AVX2 build,
 * oracle/_ref/libabpoa_ref.so) and the reference bar/impl/poaBarAligner.c (oracle/_ref/libbar_ref.so)
 * by tests/test_oracle_vs_ref.py, and against the committed golden vectors in tests/golden/ (generated
 * from those reference builds by scripts/make_golden.py) by tests/test_oracle_golden.py.
 */
#ifndef POA_ORACLE_H
#define POA_ORACLE_H
#include <stdint.h>

/* same field order as oracle/ref_harness.c:ref_params_t */
typedef struct {
    int wb; float wf;
    int gap_open1, gap_ext1, gap_open2, gap_ext2;
    int mat[25];
    int k, w, min_w;
    int progressive_poa, disable_seeding;
} oracle_params_t;

/* abpoa_msa (abpoa_align.c:401-471) under Cactus' settings: global mode, convex gap, adaptive band,
 * seeding disabled, read ids on, MSA output. seqs are 0..4 codes, concatenated. Returns msa_len;
 * *msa_out = malloc'd n_seq*msa_len row-major bytes (0-3 ACGT, 4 N, 5 gap). */
int oracle_poa_msa(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat, uint8_t **msa_out);

/* as above + per-alignment intermediates, same word layout as ref_harness.c:ref_poa_msa_trace */
int64_t *oracle_poa_msa_trace(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat, int64_t *n_words);

/* banded cell count only (sum over alignments and rows of dp_end-dp_beg+1) */
int64_t oracle_poa_cells(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat);

void oracle_free(void *p);

/* ---- BAR level (bar_oracle.c), restating bar/impl/poaBarAligner.c -------------------------------- */
typedef struct {
    int64_t seq_no, column_no;
    int *seq_lens;        /* [seq_no] */
    uint8_t *msa;         /* [seq_no * column_no], 0-3 ACGT, 4 N, 5 gap */
} oracle_msa_t;

/* msa_make_partial_order_alignment (poaBarAligner.c:463-749). seqs: ASCII strings. */
oracle_msa_t *oracle_msa_make_partial_order_alignment(const oracle_params_t *p, char **seqs, const int *seq_lens,
        int64_t seq_no, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff);
/* make_consistent_partial_order_alignments (poaBarAligner.c:751-801) */
oracle_msa_t **oracle_make_consistent_partial_order_alignments(const oracle_params_t *p, int64_t end_no,
        const int64_t *end_lengths, char ***end_strings, int **end_string_lengths, int64_t **right_end_indexes,
        int64_t **right_end_row_indexes, int64_t **overlaps, int64_t window_size, int64_t max_prog_rows,
        double max_prog_length_diff);
void oracle_msa_destruct(oracle_msa_t *m);

#endif
