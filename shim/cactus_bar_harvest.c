/*
 * cactus_bar_harvest.c -- instrumentation for a REFERENCE (CPU) Cactus build: records the inputs of every
 * make_consistent_partial_order_alignments / msa_make_partial_order_alignment call of a bar() run into one file, so that the BAR
 * workload of a real dataset (evolverMammals, 10way-mhc, yeast ... -- BASELINE.json configs[0,1,3,4]; the data is not available
 * offline) can be captured once on a box that has cactusTestData and replayed anywhere by
 * `bench.py --workload <file>` (GPU engine and CPU reference on exactly the same ends).
 *
 * Build hook (oracle/Makefile does exactly this for oracle/_ref/libflower_harvest.so): the reference's poaBarAligner.o is linked
 * TWICE -- once with the two entry points weakened (everything in Cactus keeps calling them by name and reaches the wrappers
 * below), once as a private copy whose own global symbols carry the prefix ref_ (objcopy --redefine-syms with the list of the
 * object's defined symbols), which the wrappers forward to:
 *     objcopy --weaken-symbol=msa_make_partial_order_alignment --weaken-symbol=make_consistent_partial_order_alignments poaBarAligner.o weak.o
 *     nm -g --defined-only poaBarAligner.o | awk '{print $3 " ref_" $3}' > syms;  objcopy --redefine-syms=syms poaBarAligner.o private.o
 * then run cactus_consolidated with BARB200_HARVEST=/path/to/dump. Without the variable the wrappers just forward.
 * Only top-level calls are recorded (inside the private copy make_consistent_... calls its own msa_make_...: those nested calls
 * are part of the recorded flower). The alignment itself is the unmodified reference's.
 *
 * Record layout (little endian): int64 magic 0x4852414232303042 ("B002RABH"), kind (1 consistent / 2 single end), end_no,
 * window_size, max_prog_rows; double max_prog_length_diff; per end: int64 n; per string: int64 length, right_end_index,
 * right_end_row_index, overlap; then the string bytes of the end, concatenated.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "poaBarAligner.h"

Msa *ref_msa_make_partial_order_alignment(char **seqs, int *seq_lens, int64_t seq_no, int64_t window_size, int64_t max_prog_rows,
                                                double max_prog_length_diff, abpoa_para_t *poa_parameters);
Msa **ref_make_consistent_partial_order_alignments(int64_t end_no, int64_t *end_lengths, char ***end_strings, int **end_string_lengths,
        int64_t **right_end_indexes, int64_t **right_end_row_indexes, int64_t **overlaps, int64_t window_size, int64_t max_prog_rows,
        double max_prog_length_diff, abpoa_para_t *poa_parameters);

static pthread_mutex_t harvest_mutex = PTHREAD_MUTEX_INITIALIZER;

static void put64(FILE *f, int64_t v) { fwrite(&v, sizeof(v), 1, f); }

static void harvest(int64_t kind, int64_t end_no, int64_t *end_lengths, char ***end_strings, int **end_string_lengths, int64_t **rei, int64_t **reri,
                    int64_t **ov, int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff) {
    const char *path = getenv("BARB200_HARVEST");
    if (path == NULL) return;
    pthread_mutex_lock(&harvest_mutex);
    FILE *f = fopen(path, "ab");
    if (f != NULL) {
        put64(f, 0x4852414232303042LL); put64(f, kind); put64(f, end_no); put64(f, window_size); put64(f, max_prog_rows);
        fwrite(&max_prog_length_diff, sizeof(double), 1, f);
        for (int64_t e = 0; e < end_no; e++) {
            put64(f, end_lengths[e]);
            for (int64_t i = 0; i < end_lengths[e]; i++) {
                put64(f, end_string_lengths[e][i]); put64(f, rei ? rei[e][i] : -1); put64(f, reri ? reri[e][i] : -1); put64(f, ov ? ov[e][i] : 0);
            }
            for (int64_t i = 0; i < end_lengths[e]; i++) fwrite(end_strings[e][i], 1, (size_t)end_string_lengths[e][i], f);
        }
        fclose(f);
    }
    pthread_mutex_unlock(&harvest_mutex);
}

Msa *msa_make_partial_order_alignment(char **seqs, int *seq_lens, int64_t seq_no, int64_t window_size, int64_t max_prog_rows,
                                      double max_prog_length_diff, abpoa_para_t *poa_parameters) {
    harvest(2, 1, &seq_no, &seqs, &seq_lens, NULL, NULL, NULL, window_size, max_prog_rows, max_prog_length_diff);
    return ref_msa_make_partial_order_alignment(seqs, seq_lens, seq_no, window_size, max_prog_rows, max_prog_length_diff, poa_parameters);
}

Msa **make_consistent_partial_order_alignments(int64_t end_no, int64_t *end_lengths, char ***end_strings, int **end_string_lengths,
        int64_t **right_end_indexes, int64_t **right_end_row_indexes, int64_t **overlaps, int64_t window_size, int64_t max_prog_rows,
        double max_prog_length_diff, abpoa_para_t *poa_parameters) {
    harvest(1, end_no, end_lengths, end_strings, end_string_lengths, right_end_indexes, right_end_row_indexes, overlaps, window_size, max_prog_rows,
            max_prog_length_diff);
    return ref_make_consistent_partial_order_alignments(end_no, end_lengths, end_strings, end_string_lengths, right_end_indexes, right_end_row_indexes,
                                                        overlaps, window_size, max_prog_rows, max_prog_length_diff, poa_parameters);
}
