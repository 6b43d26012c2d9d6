/*
 * cactus_bar_shim.c -- the thin C host shim between Cactus' BAR code and libbarb200 (include/barb200.h).
 *
 * Compiled AGAINST THE CACTUS TREE (bar/inc/poaBarAligner.h and what it includes) and linked into cactusBarLib.a in
 * place of the two reference functions it re-exports with their reference signatures:
 *
 *     Msa  *msa_make_partial_order_alignment(...)            bar/inc/poaBarAligner.h:76,  bar/impl/poaBarAligner.c:463-749
 *     Msa **make_consistent_partial_order_alignments(...)    bar/inc/poaBarAligner.h:108, bar/impl/poaBarAligner.c:751-801
 *
 * Everything above them -- bar() (bar/impl/bar.c:52), make_flower_alignment_poa (poaBarAligner.c:1115), the
 * Flower/End/Cap/CactusDisk API, CAF -- stays the reference's code and keeps calling these two symbols, so
 * cactus_consolidated and the Python/Toil pipeline run unchanged. See INTEGRATION.md for the build hook.
 *
 * Semantics kept: ownership (the returned Msa owns seqs / seq_lens exactly as the reference's does,
 * poaBarAligner.c:474-475, 708-717; released by the reference's msa_destruct), re-entrancy (called concurrently from
 * OpenMP teams, bar.c:90-94, poaBarAligner.c:772 -- calls are funnelled into one engine context, which serialises
 * device batches), error convention (no return codes: failure -> st_errAbort, i.e. message + exit(1), as every
 * other fatal condition in this code path).
 * There is no CPU fallback: if no CUDA device is usable the process aborts with the engine's message.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "poaBarAligner.h"
#include "barb200.h"

static pthread_mutex_t shim_mutex = PTHREAD_MUTEX_INITIALIZER;
/* one engine per distinct parameter set, created on first use and kept for the life of the process: another OpenMP thread
 * may be inside a call on any of them, so a context is never destroyed or replaced in place */
#define SHIM_MAX_CTX 16
static struct { barb200_params p; barb200_ctx *ctx; } shim_ctxs[SHIM_MAX_CTX];
static int shim_ctx_no = 0;

/* abpoa_para_t (as built by abpoaParamaters_constructFromCactusParams, poaBarAligner.c:24-81) -> barb200_params */
static void params_from_abpoa(const abpoa_para_t *abpt, barb200_params *p) {
    barb200_params_default(p);
    memcpy(p->mat, abpt->mat, 25 * sizeof(int));
    p->gap_open1 = abpt->gap_open1; p->gap_ext1 = abpt->gap_ext1;
    p->gap_open2 = abpt->gap_open2; p->gap_ext2 = abpt->gap_ext2;
    p->wb = abpt->wb; p->wf = abpt->wf;
    p->k = abpt->k; p->w = abpt->w; p->min_w = abpt->min_w;
    p->progressive_poa = abpt->progressive_poa;
    p->disable_seeding = abpt->disable_seeding;
    const char *dev = getenv("BARB200_DEVICE");          /* one cactus_consolidated process per GPU */
    if (dev) p->device = atoi(dev);
}

/* field-wise comparison (memcmp would read struct padding) */
static int params_equal(const barb200_params *a, const barb200_params *b) {
    return memcmp(a->mat, b->mat, sizeof(a->mat)) == 0 && a->gap_open1 == b->gap_open1 && a->gap_ext1 == b->gap_ext1 &&
           a->gap_open2 == b->gap_open2 && a->gap_ext2 == b->gap_ext2 && a->wb == b->wb && a->wf == b->wf && a->k == b->k &&
           a->w == b->w && a->min_w == b->min_w && a->progressive_poa == b->progressive_poa &&
           a->disable_seeding == b->disable_seeding && a->device == b->device;
}

static barb200_ctx *shim_context(abpoa_para_t *abpt) {
    barb200_params p;
    params_from_abpoa(abpt, &p);
    pthread_mutex_lock(&shim_mutex);
    for (int i = 0; i < shim_ctx_no; i++) {
        if (params_equal(&p, &shim_ctxs[i].p)) {
            barb200_ctx *ctx = shim_ctxs[i].ctx;
            pthread_mutex_unlock(&shim_mutex);
            return ctx;
        }
    }
    if (shim_ctx_no == SHIM_MAX_CTX) {
        pthread_mutex_unlock(&shim_mutex);
        st_errAbort("barb200: more than %d distinct POA parameter sets in one process", SHIM_MAX_CTX);
    }
    char err[256];
    barb200_ctx *ctx = barb200_create(&p, err, (int)sizeof(err));
    if (ctx == NULL) {
        pthread_mutex_unlock(&shim_mutex);
        st_errAbort("barb200: cannot create the GPU BAR engine: %s", err);
    }
    shim_ctxs[shim_ctx_no].p = p;
    shim_ctxs[shim_ctx_no].ctx = ctx;
    shim_ctx_no++;
    pthread_mutex_unlock(&shim_mutex);
    return ctx;
}

/* barb200_msa (flat matrix) -> the reference's Msa (row pointers); takes ownership of the caller's
 * seqs / seq_lens like the reference (seq_lens keeps the input lengths, poaBarAligner.c:708-717) */
static Msa *msa_from_engine(barb200_msa *m, char **seqs, int *seq_lens) {
    Msa *msa = st_malloc(sizeof(Msa));
    msa->seq_no = m->seq_no;
    msa->seqs = seqs;
    msa->seq_lens = seq_lens;
    msa->column_no = (int)m->column_no;
    msa->msa_seq = st_malloc(sizeof(uint8_t *) * (m->seq_no > 0 ? m->seq_no : 1));
    for (int64_t i = 0; i < m->seq_no; i++) {
        msa->msa_seq[i] = st_malloc(m->column_no > 0 ? m->column_no : 1);
        memcpy(msa->msa_seq[i], m->msa + i * m->column_no, m->column_no);
    }
    barb200_msa_destruct(m);
    return msa;
}

Msa *msa_make_partial_order_alignment(char **seqs, int *seq_lens, int64_t seq_no, int64_t window_size,
                                      int64_t max_prog_rows, double max_prog_length_diff, abpoa_para_t *poa_parameters) {
    barb200_ctx *ctx = shim_context(poa_parameters);
    barb200_msa *m = barb200_msa_make_partial_order_alignment(ctx, seqs, seq_lens, seq_no, window_size, max_prog_rows,
                                                              max_prog_length_diff);
    if (m == NULL) {
        st_errAbort("barb200: msa_make_partial_order_alignment failed: %s", barb200_last_error(ctx));
    }
    return msa_from_engine(m, seqs, seq_lens);
}

Msa **make_consistent_partial_order_alignments(int64_t end_no, int64_t *end_lengths, char ***end_strings,
        int **end_string_lengths, int64_t **right_end_indexes, int64_t **right_end_row_indexes, int64_t **overlaps,
        int64_t window_size, int64_t max_prog_rows, double max_prog_length_diff, abpoa_para_t *poa_parameters) {
    barb200_ctx *ctx = shim_context(poa_parameters);
    /* all ends of the flower in ONE batched device call (the reference loops, optionally with a nested OpenMP team) */
    barb200_msa **ms = barb200_make_consistent_partial_order_alignments(ctx, end_no, end_lengths, end_strings,
            end_string_lengths, right_end_indexes, right_end_row_indexes, overlaps, window_size, max_prog_rows,
            max_prog_length_diff);
    if (ms == NULL) {
        st_errAbort("barb200: make_consistent_partial_order_alignments failed: %s", barb200_last_error(ctx));
    }
    Msa **msas = st_malloc(sizeof(Msa *) * (end_no > 0 ? end_no : 1));
    for (int64_t i = 0; i < end_no; i++) {
        msas[i] = msa_from_engine(ms[i], end_strings[i], end_string_lengths[i]);
    }
    barb200_free(ms);
    return msas;
}
