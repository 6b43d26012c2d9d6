/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver around the UNMODIFIED reference abPOA (compiled by oracle/Makefile from
 * /root/reference/submodules/abPOA, AVX2 flavour exactly as Cactus ships it, include.mk:95-105).
 * It exists for two reasons:
 *   1. run abpoa_msa() with the parameters Cactus' BAR phase uses
 *      (bar/impl/poaBarAligner.c:24-112, 565-614) and hand back msa_base/msa_len;
 *   2. expose per-alignment intermediates the reference never returns (read_id_map, graph_cigar,
 *      dp_beg/dp_end per row, best score, banded cell count) by re-running the same
 *      per-sequence loop abpoa_anchor_poa executes when seeding is disabled
 *      (abpoa_align.c:208-309: align_sequence_to_subgraph(SRC,SINK) + add_subgraph_alignment),
 *      and asserting that its final MSA equals abpoa_msa()'s.
 * These dumps are the golden vectors the reference's own tests lack.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "abpoa.h"
#include "abpoa_seed.h"
#include "abpoa_seq.h"

typedef struct {
    int wb; float wf;
    int gap_open1, gap_ext1, gap_open2, gap_ext2;
    int mat[25];
    int k, w, min_w;
    int progressive_poa, disable_seeding;
} ref_params_t;

/* what copy_abpoa_params() + abpoaParamaters_constructFromCactusParams() produce
 * (bar/impl/poaBarAligner.c:24-112) */
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

/* Plain abpoa_msa. msa_out: malloc'd n_seq*msa_len bytes, row-major. Returns msa_len. */
int ref_poa_msa(const ref_params_t *p, int n_seq, const int *lens, const uint8_t *flat, uint8_t **msa_out) {
    abpoa_para_t *abpt = make_para(p);
    abpoa_t *ab = abpoa_init();
    uint8_t **seqs = (uint8_t **)malloc(sizeof(uint8_t *) * n_seq);
    int *l = (int *)malloc(sizeof(int) * n_seq);
    size_t off = 0;
    for (int i = 0; i < n_seq; ++i) { seqs[i] = (uint8_t *)flat + off; off += lens[i]; l[i] = lens[i]; }
    abpoa_msa(ab, abpt, n_seq, NULL, l, seqs, NULL, NULL);
    int msa_len = ab->abc->msa_len;
    uint8_t *out = (uint8_t *)malloc((size_t)n_seq * (msa_len > 0 ? msa_len : 1));
    for (int i = 0; i < n_seq; ++i) memcpy(out + (size_t)i * msa_len, ab->abc->msa_base[i], msa_len);
    *msa_out = out;
    abpoa_free(ab); abpoa_free_para(abpt);
    free(seqs); free(l);
    return msa_len;
}

/* Trace record layout (int64 words, then raw payload), see tests/_reflib.py for the reader:
 *   header: n_seq, msa_len, total_cells
 *   read_id_map[n_seq]
 *   per alignment a (n_seq of them, in guide-tree order):
 *      read_id, qlen, node_n(before alignment), n_cigar, best_score, n_rows(=node_n-1, 0 when graph empty)
 *      cigar[n_cigar] (u64 as i64), dp_beg[n_rows], dp_end[n_rows]
 *   msa bytes packed 8 per word at the end
 */
typedef struct { int64_t *w; size_t n, m; } wbuf_t;
static void wpush(wbuf_t *b, int64_t v) {
    if (b->n == b->m) { b->m = b->m ? b->m * 2 : 1024; b->w = (int64_t *)realloc(b->w, b->m * sizeof(int64_t)); }
    b->w[b->n++] = v;
}

int64_t *ref_poa_msa_trace(const ref_params_t *p, int n_seq, const int *lens, const uint8_t *flat, int64_t *n_words) {
    abpoa_para_t *abpt = make_para(p);
    abpoa_t *ab = abpoa_init();
    uint8_t **seqs = (uint8_t **)malloc(sizeof(uint8_t *) * n_seq);
    int *l = (int *)malloc(sizeof(int) * n_seq);
    size_t off = 0; int max_len = 0;
    for (int i = 0; i < n_seq; ++i) {
        seqs[i] = (uint8_t *)flat + off; off += lens[i]; l[i] = lens[i];
        if (lens[i] > max_len) max_len = lens[i];
    }
    /* abpoa_msa prologue, abpoa_align.c:405-440 */
    abpoa_reset(ab, abpt, 1024);
    ab->abs->n_seq += n_seq; abpoa_realloc_seq(ab->abs);
    int **weights = (int **)malloc(sizeof(int *) * n_seq);
    for (int i = 0; i < n_seq; ++i) {
        weights[i] = (int *)malloc(sizeof(int) * l[i]);
        for (int j = 0; j < l[i]; ++j) weights[i][j] = 1;
    }
    int *read_id_map = (int *)malloc(sizeof(int) * n_seq);
    for (int i = 0; i < n_seq; ++i) read_id_map[i] = i;
    if (!(abpt->disable_seeding && abpt->progressive_poa == 0)) {
        ab_u64_v par_anchors = {0, 0, 0}; int *par_c = (int *)calloc(n_seq, sizeof(int));
        abpoa_build_guide_tree_partition(seqs, l, n_seq, abpt, read_id_map, &par_anchors, par_c);
        if (par_anchors.n != 0) { fprintf(stderr, "ref_harness: anchors present; seeding must be disabled\n"); exit(1); }
        free(par_c); if (par_anchors.m) free(par_anchors.a);
    }
    wbuf_t b = {0, 0, 0};
    wpush(&b, n_seq); wpush(&b, 0); wpush(&b, 0);
    for (int i = 0; i < n_seq; ++i) wpush(&b, read_id_map[i]);
    int64_t total_cells = 0;
    for (int _i = 0; _i < n_seq; ++_i) {
        int i = read_id_map[_i], qlen = l[i];
        abpoa_res_t res; memset(&res, 0, sizeof(res));
        int node_n = ab->abg->node_n;
        int rc = abpoa_align_sequence_to_subgraph(ab, abpt, ABPOA_SRC_NODE_ID, ABPOA_SINK_NODE_ID, seqs[i], qlen, &res);
        int n_rows = rc < 0 ? 0 : node_n - 1;
        wpush(&b, i); wpush(&b, qlen); wpush(&b, node_n); wpush(&b, res.n_cigar);
        wpush(&b, rc < 0 ? 0 : res.best_score); wpush(&b, n_rows);
        for (int c = 0; c < res.n_cigar; ++c) wpush(&b, (int64_t)res.graph_cigar[c]);
        for (int r = 0; r < n_rows; ++r) wpush(&b, ab->abm->dp_beg[r]);
        for (int r = 0; r < n_rows; ++r) { wpush(&b, ab->abm->dp_end[r]); total_cells += ab->abm->dp_end[r] - ab->abm->dp_beg[r] + 1; }
        abpoa_add_subgraph_alignment(ab, abpt, ABPOA_SRC_NODE_ID, ABPOA_SINK_NODE_ID, seqs[i], weights[i], qlen, NULL, res, i, n_seq, 1);
        if (res.n_cigar) free(res.graph_cigar);
    }
    abpoa_generate_rc_msa(ab, abpt);
    int msa_len = ab->abc->msa_len;
    /* cross-check against the reference's own driver */
    uint8_t *chk = NULL; int chk_len = ref_poa_msa(p, n_seq, lens, flat, &chk);
    if (chk_len != msa_len) { fprintf(stderr, "ref_harness: driver mismatch msa_len %d vs %d\n", msa_len, chk_len); exit(1); }
    for (int i = 0; i < n_seq; ++i)
        if (memcmp(chk + (size_t)i * msa_len, ab->abc->msa_base[i], msa_len) != 0) { fprintf(stderr, "ref_harness: driver mismatch row %d\n", i); exit(1); }
    free(chk);
    b.w[1] = msa_len; b.w[2] = total_cells;
    size_t nbytes = (size_t)n_seq * msa_len, nw = (nbytes + 7) / 8;
    size_t base = b.n;
    for (size_t k = 0; k < nw; ++k) wpush(&b, 0);
    uint8_t *dst = (uint8_t *)(b.w + base);
    for (int i = 0; i < n_seq; ++i) memcpy(dst + (size_t)i * msa_len, ab->abc->msa_base[i], msa_len);
    for (int i = 0; i < n_seq; ++i) free(weights[i]);
    free(weights); free(read_id_map); free(seqs); free(l);
    abpoa_free(ab); abpoa_free_para(abpt);
    *n_words = (int64_t)b.n;
    return b.w;
}

void ref_free(void *p) { free(p); }

/* CPU baseline driver: n_jobs independent abpoa_msa calls, OpenMP over jobs with schedule(dynamic,1) -- the
 * reference's own unit of parallelism (bar/impl/bar.c:90-94, poaBarAligner.c:772). Each job does what one sliding
 * window costs in the shim: abpoa_init + copy of the parameters + abpoa_msa + abpoa_free (poaBarAligner.c:565-628).
 * Returns the wall time in seconds; msa_lens[n_jobs] (may be NULL) receives msa_len, checksum (may be NULL) a sum of
 * all MSA bytes so the work cannot be optimised away. */
#include <omp.h>
#include <malloc.h>
/* FNV-1a over (msa_len, bytes): the per-end alignment hash bench.py's parity gate compares with the GPU's */
static uint64_t msa_hash(const uint8_t *m, int64_t n, int msa_len) {
    uint64_t h = 1469598103934665603ULL ^ (uint64_t)(uint32_t)msa_len;
    h *= 1099511628211ULL;
    for (int64_t k = 0; k < n; ++k) { h ^= m[k]; h *= 1099511628211ULL; }
    return h;
}
/* Allocator stand-in. Cactus links jemalloc (include.mk:53-64,119-123) because every window mallocs and frees abPOA's
 * ~84 MB DP matrix; jemalloc keeps such extents cached, glibc mmap()s and munmap()s them every time (page faults +
 * mmap_sem contention across threads). jemalloc's autoconf build cannot run in this image, so mode 1 makes glibc retain
 * and reuse large blocks instead (no mmap for big requests, no trimming): the same effect on this workload. */
void ref_malloc_mode(int mode) {
    if (mode) { mallopt(M_MMAP_MAX, 0); mallopt(M_TRIM_THRESHOLD, 0x7fffffff); mallopt(M_TOP_PAD, 512 << 20); }
    else { mallopt(M_MMAP_MAX, 65536); mallopt(M_TRIM_THRESHOLD, 128 * 1024); mallopt(M_TOP_PAD, 128 * 1024); }
}
double ref_poa_msa_many(const ref_params_t *p, int64_t n_jobs, const int *n_seq, const int *lens, const uint8_t *flat,
                        int threads, int *msa_lens, uint64_t *checksum, uint64_t *hashes) {
    int64_t *len_off = (int64_t *)malloc(sizeof(int64_t) * (n_jobs + 1)), *seq_off = (int64_t *)malloc(sizeof(int64_t) * (n_jobs + 1));
    int64_t lo = 0, so = 0;
    for (int64_t j = 0; j < n_jobs; ++j) { len_off[j] = lo; seq_off[j] = so; for (int i = 0; i < n_seq[j]; ++i) so += lens[lo + i]; lo += n_seq[j]; }
    if (threads <= 0) threads = omp_get_max_threads();
    { abpoa_para_t *warm = make_para(p); abpoa_free_para(warm); }   /* global tables are initialised once, up front */
    uint64_t sum = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : sum)
    for (int64_t j = 0; j < n_jobs; ++j) {
        uint8_t *msa = NULL;
        int ml = ref_poa_msa(p, n_seq[j], lens + len_off[j], flat + seq_off[j], &msa);
        if (msa_lens) msa_lens[j] = ml;
        for (int64_t k = 0; k < (int64_t)n_seq[j] * ml; ++k) sum += msa[k];
        if (hashes) hashes[j] = msa_hash(msa, (int64_t)n_seq[j] * ml, ml);
        free(msa);
    }
    double t1 = omp_get_wtime();
    if (checksum) *checksum = sum;
    free(len_off); free(seq_off);
    return t1 - t0;
}
