/*
 * flower_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the reference's FLOWER-level BAR code -- make_flower_alignment_poa (bar/impl/poaBarAligner.c:1115-1237),
 * stPinchIterator_constructFromAlignedBlocks (:1291-1299) and bar() itself (bar/impl/bar.c:52-176) -- on flowers described
 * by flat arrays, and returns what they produce as a flat int64 stream, so that tests/ can require the SAME ordered
 * AlignmentBlock / stPinch streams and the same post-BAR flower from
 *     oracle/_ref/libflower_ref.so   the unmodified reference objects, and
 *     oracle/_ref/libflower_shim.so  the same objects with shim/cactus_bar_shim.c (+ pecan shim) linked in place of the
 *                                    reference's definitions (the CACTUS_BAR_B200 build of INTEGRATION.md).
 * Both libraries are built by oracle/Makefile from the sources where they lie under /root/reference.
 *
 * cactusParams_get_* are supplied here from a key/value table (the reference's cactus_params_parser.c needs libxml2, which this
 * image lacks): keys are the XML path below <cactusWorkflowConfig>, e.g. "bar/poa/partialOrderAlignmentWindow".
 *
 * A flower is described the way bar/tests/flowersShared.h builds one: sequences, stub ends (with a side), and adjacencies --
 * each a pair of caps on one sequence at coordinates lo < hi, on the side-0 view of end A and the side-1 view of end B, in the
 * positive-strand (caps at lo / hi, strand 1) or negative-strand (caps at hi / lo, strand 0) representation.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cactus.h"
#include "sonLib.h"
#include "poaBarAligner.h"
#include "stCaf.h"
#include "stPinchIterator.h"
#include <omp.h>

void bar(stList *flowers, CactusParams *params, CactusDisk *cactusDisk, stList *listOfEndAlignmentFiles);

/* ---- parameter table ------------------------------------------------------------------------------------------------ */
#define MAX_PARAMS 512
static struct { char *key, *val; } g_tab[MAX_PARAMS];
static int g_tab_n = 0;

void flower_harness_clear_params(void) {
    for (int i = 0; i < g_tab_n; ++i) { free(g_tab[i].key); free(g_tab[i].val); }
    g_tab_n = 0;
}
void flower_harness_set_param(const char *path, const char *value) {
    for (int i = 0; i < g_tab_n; ++i)
        if (strcmp(g_tab[i].key, path) == 0) { free(g_tab[i].val); g_tab[i].val = strdup(value); return; }
    if (g_tab_n == MAX_PARAMS) { fprintf(stderr, "flower_harness: parameter table full\n"); exit(1); }
    g_tab[g_tab_n].key = strdup(path); g_tab[g_tab_n].val = strdup(value); ++g_tab_n;
}
static const char *lookup(int num, va_list ap) {
    char path[512]; path[0] = 0;
    for (int i = 0; i < num; ++i) { if (i) strcat(path, "/"); strncat(path, va_arg(ap, const char *), 200); }
    for (int i = 0; i < g_tab_n; ++i) if (strcmp(g_tab[i].key, path) == 0) return g_tab[i].val;
    fprintf(stderr, "flower_harness: parameter %s not set\n", path);
    exit(1);
}
char *cactusParams_get_string(CactusParams *p, int num, ...) { va_list ap; va_start(ap, num); const char *v = lookup(num, ap); va_end(ap); return stString_copy(v); }
int64_t cactusParams_get_int(CactusParams *p, int num, ...) { va_list ap; va_start(ap, num); const char *v = lookup(num, ap); va_end(ap); return atoll(v); }
double cactusParams_get_float(CactusParams *p, int num, ...) { va_list ap; va_start(ap, num); const char *v = lookup(num, ap); va_end(ap); return atof(v); }
int64_t *cactusParams_get_ints(CactusParams *p, int64_t *length, int num, ...) {
    va_list ap; va_start(ap, num); const char *v = lookup(num, ap); va_end(ap);
    char *c = stString_copy(v); int64_t n = 0, cap = 8, *out = st_malloc(sizeof(int64_t) * cap);
    for (char *t = strtok(c, " "); t; t = strtok(NULL, " ")) { if (n == cap) { cap *= 2; out = realloc(out, sizeof(int64_t) * cap); } out[n++] = atoll(t); }
    free(c); *length = n; return out;
}
void cactusParams_set_root(CactusParams *p, int num, ...) { (void)p; (void)num; }
CactusParams *cactusParams_load(char *file_name) { (void)file_name; return NULL; }
void cactusParams_destruct(CactusParams *p) { (void)p; }

/* ---- output stream -------------------------------------------------------------------------------------------------- */
typedef struct { int64_t *w; size_t n, cap; } wbuf;
static void push(wbuf *b, int64_t v) {
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 1024; b->w = realloc(b->w, sizeof(int64_t) * b->cap); }
    b->w[b->n++] = v;
}

/* ---- flowers ------------------------------------------------------------------------------------------------------- */
typedef struct {
    Flower *flower;
    int64_t n_seq, n_caps;
    Name *seq_names;
    Name *cap_names;            /* [2 * n_adj]: cap A of adjacency i at 2i, cap B at 2i+1 */
} built;

typedef struct {
    CactusDisk *disk; EventTree *eventTree; Event **leaves; int64_t n_events;
    built *f; int64_t n_flowers, cap_flowers;
} session;

static End *view(End *end, int side) { return (end_getSide(end) ? 1 : 0) == side ? end : end_getReverse(end); }

void *flower_harness_begin(int64_t n_events) {
    session *S = st_calloc(1, sizeof(session));
    st_randomSeed(20260923);            /* the cPecan configuration draws st_random() numbers (tie breaks, multipleAligner.c:853) */
    S->disk = cactusDisk_construct();
    S->eventTree = eventTree_construct2(S->disk);
    Event *root = eventTree_getRootEvent(S->eventTree);
    Event *anc = event_construct3("ANC", 0.1, root, S->eventTree);
    S->n_events = n_events;
    S->leaves = st_malloc(sizeof(Event *) * (n_events > 0 ? n_events : 1));
    for (int64_t i = 0; i < n_events; ++i) { char h[32]; sprintf(h, "LEAF%d", (int)i); S->leaves[i] = event_construct3(h, 0.1, anc, S->eventTree); }
    return S;
}

/* adds one flower to the session's disk; returns its index */
int64_t flower_harness_add_flower(void *session_, int64_t n_seq, const char **seqs, const int *seq_event, int64_t n_ends, const int *end_side,
                                  int64_t n_adj, const int64_t *adj) {
    session *S = session_;
    if (S->n_flowers == S->cap_flowers) { S->cap_flowers = S->cap_flowers ? 2 * S->cap_flowers : 16; S->f = realloc(S->f, sizeof(built) * S->cap_flowers); }
    built *B = &S->f[S->n_flowers];
    memset(B, 0, sizeof(*B));
    B->flower = flower_construct(S->disk);
    group_construct2(B->flower);                      /* stCaf_setup expects a leaf group (caf/tests/filteringTest.c:71) */
    B->n_seq = n_seq; B->seq_names = st_malloc(sizeof(Name) * (n_seq > 0 ? n_seq : 1));
    Sequence **sq = st_malloc(sizeof(Sequence *) * (n_seq > 0 ? n_seq : 1));
    for (int64_t i = 0; i < n_seq; ++i) {
        char h[48]; sprintf(h, ">f%ds%d", (int)S->n_flowers, (int)i);
        sq[i] = sequence_construct(1, (int64_t)strlen(seqs[i]), seqs[i], h, S->leaves[seq_event[i]], S->disk);
        flower_addSequence(B->flower, sq[i]);
        B->seq_names[i] = sequence_getName(sq[i]);
    }
    End **ends = st_malloc(sizeof(End *) * (n_ends > 0 ? n_ends : 1));
    for (int64_t i = 0; i < n_ends; ++i) ends[i] = end_construct2(end_side[i] != 0, 1, B->flower);
    B->n_caps = 2 * n_adj; B->cap_names = st_malloc(sizeof(Name) * (B->n_caps > 0 ? B->n_caps : 1));
    for (int64_t i = 0; i < n_adj; ++i) {
        const int64_t *a = adj + 6 * i;           /* seq, lo, hi, positive representation?, end A, end B */
        Sequence *s = sq[a[0]];
        Cap *ca, *cb;
        if (a[3]) { ca = cap_construct2(view(ends[a[4]], 0), a[1], 1, s); cb = cap_construct2(view(ends[a[5]], 1), a[2], 1, s); }
        else { ca = cap_construct2(view(ends[a[4]], 0), a[2], 0, s); cb = cap_construct2(view(ends[a[5]], 1), a[1], 0, s); }
        cap_makeAdjacent(ca, cb);
        B->cap_names[2 * i] = cap_getName(ca); B->cap_names[2 * i + 1] = cap_getName(cb);
    }
    free(ends); free(sq);
    return S->n_flowers++;
}
static int64_t cap_index(built *B, Name n) { for (int64_t i = 0; i < B->n_caps; ++i) if (B->cap_names[i] == n) return i; return -1; }
static int64_t seq_index(built *B, Name n) { for (int64_t i = 0; i < B->n_seq; ++i) if (B->seq_names[i] == n) return i; return -1; }

/* Canonical dump of what BAR produced: every block of the flower hierarchy as the sorted list of its segments
 * (sequence index, lowest forward-strand coordinate, length, orientation within the block), with the block's arbitrary
 * orientation normalised (the first segment in sorted order is forward), and the blocks themselves sorted. Object names,
 * iteration order and block orientation depend on hash iteration / the order in which concurrently processed flowers drew
 * ids from the disk -- they differ between two runs of the UNMODIFIED reference -- so they are not part of the comparison. */
typedef struct { int64_t *w; size_t n; } rec;
static int rec_cmp(const void *a, const void *b) {
    const rec *x = a, *y = b;
    const size_t n = x->n < y->n ? x->n : y->n;
    for (size_t i = 0; i < n; ++i) if (x->w[i] != y->w[i]) return x->w[i] < y->w[i] ? -1 : 1;
    return x->n < y->n ? -1 : (x->n > y->n ? 1 : 0);
}
typedef struct { rec *r; size_t n, cap; } recs;
static void recs_add(recs *R, int64_t *w, size_t n) {
    if (R->n == R->cap) { R->cap = R->cap ? 2 * R->cap : 64; R->r = realloc(R->r, sizeof(rec) * R->cap); }
    R->r[R->n].w = w; R->r[R->n].n = n; ++R->n;
}
static void collect_blocks(built *B, Flower *f, recs *out) {
    Flower_EndIterator *eit = flower_getEndIterator(f); End *e;
    while ((e = flower_getNextEnd(eit)) != NULL) {
        if (!(end_isBlockEnd(e) && end_getSide(e))) continue;
        Block *b = end_getBlock(e);
        recs segs = {0, 0, 0};
        Block_InstanceIterator *sit = block_getInstanceIterator(b); Segment *sg;
        while ((sg = block_getNext(sit)) != NULL) {
            Sequence *sq = segment_getSequence(sg);
            if (sq == NULL) continue;
            Segment *fw = segment_getStrand(sg) ? sg : segment_getReverse(sg);
            int64_t *w = malloc(sizeof(int64_t) * 4);
            w[0] = seq_index(B, sequence_getName(sq)); w[1] = segment_getStart(fw); w[2] = segment_getLength(sg); w[3] = segment_getStrand(sg) ? 1 : 0;
            recs_add(&segs, w, 4);
        }
        block_destructInstanceIterator(sit);
        if (segs.n == 0) { free(segs.r); continue; }
        qsort(segs.r, segs.n, sizeof(rec), rec_cmp);
        const int64_t flip = segs.r[0].w[3] ? 0 : 1;
        int64_t *w = malloc(sizeof(int64_t) * (1 + 4 * segs.n));
        w[0] = block_getLength(b);
        for (size_t i = 0; i < segs.n; ++i) { memcpy(w + 1 + 4 * i, segs.r[i].w, sizeof(int64_t) * 4); w[1 + 4 * i + 3] ^= flip; free(segs.r[i].w); }
        recs_add(out, w, 1 + 4 * segs.n);
        free(segs.r);
    }
    flower_destructEndIterator(eit);
    Flower_GroupIterator *git = flower_getGroupIterator(f); Group *g;
    while ((g = flower_getNextGroup(git)) != NULL) if (!group_isLeaf(g)) collect_blocks(B, group_getNestedFlower(g), out);
    flower_destructGroupIterator(git);
}
static void dump_flower(built *B, Flower *f, wbuf *o, int depth) {
    (void)depth;
    recs all = {0, 0, 0};
    collect_blocks(B, f, &all);
    qsort(all.r, all.n, sizeof(rec), rec_cmp);
    push(o, (int64_t)all.n);
    for (size_t i = 0; i < all.n; ++i) { push(o, (int64_t)all.r[i].n); for (size_t k = 0; k < all.r[i].n; ++k) push(o, all.r[i].w[k]); free(all.r[i].w); }
    free(all.r);
}

/* make_flower_alignment_poa + stPinchIterator_constructFromAlignedBlocks on flower `index`. Stream:
 *   n_blocks, then per block: chain length c and c x (cap index, position, strand, length);
 *   n_pinches, then per pinch: cap index 1, cap index 2, start1, start2, length, strand.
 * The POA parameters come from abpoaParamaters_constructFromCactusParams, i.e. the "bar/poa/..." table entries;
 * max_seq_length / window / mask_filter / max_prog_rows / max_prog_length_diff from the "bar/..." entries bar() reads. */
int64_t *flower_harness_blocks(void *session_, int64_t index, int64_t *n_words) {
    session *S = session_;
    built *B = &S->f[index];
    wbuf o = {0, 0, 0};
    abpoa_para_t *abpt = abpoaParamaters_constructFromCactusParams(NULL);
    stList *blocks = make_flower_alignment_poa(B->flower, cactusParams_get_int(NULL, 2, "bar", "bandingLimit"),
            cactusParams_get_int(NULL, 3, "bar", "poa", "partialOrderAlignmentWindow"),
            cactusParams_get_int(NULL, 3, "bar", "poa", "partialOrderAlignmentMaskFilter"),
            cactusParams_get_int(NULL, 3, "bar", "poa", "partialOrderAlignmentProgressiveMaxRows"),
            cactusParams_get_float(NULL, 3, "bar", "poa", "partialOrderAlignmentProgressiveMaxLengthDiff"), abpt);
    push(&o, stList_length(blocks));
    for (int64_t i = 0; i < stList_length(blocks); ++i) {
        AlignmentBlock *b = stList_get(blocks, i);
        int64_t c = 0; for (AlignmentBlock *q = b; q; q = q->next) ++c;
        push(&o, c);
        for (AlignmentBlock *q = b; q; q = q->next) { push(&o, cap_index(B, q->subsequenceIdentifier)); push(&o, q->position); push(&o, q->strand); push(&o, q->length); }
    }
    stPinchIterator *it = stPinchIterator_constructFromAlignedBlocks(blocks);
    stPinchIterator_reset(it);
    size_t at = o.n; push(&o, 0);
    stPinch *pinch, fill; int64_t np = 0;
    while ((pinch = stPinchIterator_getNext(it, &fill)) != NULL) {
        push(&o, cap_index(B, pinch->name1)); push(&o, cap_index(B, pinch->name2)); push(&o, pinch->start1); push(&o, pinch->start2);
        push(&o, pinch->length); push(&o, pinch->strand); ++np;
    }
    o.w[at] = np;
    stPinchIterator_destruct(it);
    stList_destruct(blocks);
    abpoa_free_para(abpt);
    *n_words = (int64_t)o.n;
    return o.w;
}

/* bar() on ALL flowers of the session (bar.c:52-176: alignment, pinch iterator, stCaf_setup / anneal / melt / finish); threads > 0
 * sets the OpenMP team size of bar()'s loop over flowers. Returns the wall time of the bar() call in seconds. */
double flower_harness_bar(void *session_, int threads) {
    session *S = session_;
    stList *flowers = stList_construct();
    for (int64_t i = 0; i < S->n_flowers; ++i) stList_append(flowers, S->f[i].flower);
    if (threads > 0) omp_set_num_threads(threads);
    const double t0 = omp_get_wtime();
    bar(flowers, NULL, S->disk, NULL);
    const double dt = omp_get_wtime() - t0;         /* wall time of bar() itself: alignment + CAF of every flower */
    stList_destruct(flowers);
    return dt;
}

/* recursive dump of flower `index`'s hierarchy (ends, caps, blocks and their segments, groups), e.g. after flower_harness_bar */
int64_t *flower_harness_dump(void *session_, int64_t index, int64_t *n_words) {
    session *S = session_;
    wbuf o = {0, 0, 0};
    dump_flower(&S->f[index], S->f[index].flower, &o, 0);
    *n_words = (int64_t)o.n;
    return o.w;
}

void flower_harness_end(void *session_) {
    session *S = session_;
    cactusDisk_destruct(S->disk);
    for (int64_t i = 0; i < S->n_flowers; ++i) { free(S->f[i].seq_names); free(S->f[i].cap_names); }
    free(S->f); free(S->leaves); free(S);
}

void flower_harness_free(void *p) { free(p); }
