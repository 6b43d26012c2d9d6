/*
 * pecan_ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Flat (pointer + size) driver around the UNMODIFIED reference cPecan pair-HMM
 * (/root/reference/submodules/cPecan/impl/pairwiseAligner.c + stateMachine.c, compiled by oracle/Makefile from where
 * they lie): getPosteriorProbsWithBanding (pairwiseAligner.c:766-887) with the five-state machine
 * (stateMachine.c:482-521) and a posterior callback that records the PRE-FLOOR match posteriors
 * exp(f_M + b_M - total) (the reference's own callback floors them to integers, pairwiseAligner.c:665-699, which
 * would hide differences below 1e-7 and amplify others; SURVEY.md section 7 "cPecan").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "pairwiseAligner.h"
#include "multipleAligner.h"

typedef struct {
    double threshold;
    int64_t minDiagsBetweenTraceBack, traceBackDiagonals, diagonalExpansion;
} pecan_params_t;

typedef struct { int64_t *x, *y; double *p; int64_t n, m; Band *band; int64_t lXalY; double threshold; } sink_t;

static void sink_push(sink_t *s, int64_t x, int64_t y, double p) {
    if (s->n == s->m) {
        s->m = s->m ? 2 * s->m : 1024;
        s->x = realloc(s->x, sizeof(int64_t) * s->m); s->y = realloc(s->y, sizeof(int64_t) * s->m); s->p = realloc(s->p, sizeof(double) * s->m);
    }
    s->x[s->n] = x; s->y[s->n] = y; s->p[s->n] = p; s->n++;
}

/* same walk as diagonalCalculationPosteriorMatchProbs, raw doubles */
static void record_match_posteriors(StateMachine *sM, int64_t xay, DpMatrix *forwardDpMatrix, DpMatrix *backwardDpMatrix,
                                    const SymbolString sX, const SymbolString sY, double totalProbability,
                                    PairwiseAlignmentParameters *p, void *extraArgs) {
    sink_t *s = (sink_t *)extraArgs;
    DpDiagonal *f = dpMatrix_getDiagonal(forwardDpMatrix, xay), *b = dpMatrix_getDiagonal(backwardDpMatrix, xay);
    /* the diagonal's extent: from an identical band (DpDiagonal is opaque) */
    BandIterator *it = bandIterator_construct(s->band);
    Diagonal d = bandIterator_getNext(it);
    for (int64_t k = 0; k < xay; ++k) d = bandIterator_getNext(it);
    bandIterator_destruct(it);
    for (int64_t xmy = diagonal_getMinXmy(d); xmy <= diagonal_getMaxXmy(d); xmy += 2) {
        int64_t x = diagonal_getXCoordinate(xay, xmy), y = diagonal_getYCoordinate(xay, xmy);
        if (x > 0 && y > 0) {
            double post = exp((dpDiagonal_getCell(f, xmy)[sM->matchState] + dpDiagonal_getCell(b, xmy)[sM->matchState]) - totalProbability);
            if (post >= s->threshold) sink_push(s, x - 1, y - 1, post);
        }
    }
}

/* anchors: n_anchor (x, y) pairs, 0-based sequence coordinates, strictly increasing in both. Returns the number of
 * recorded pairs; *xs, *ys, *ps are malloc'd (release with pecan_ref_free). */
int64_t pecan_ref_posteriors(const char *sx, int64_t lx, const char *sy, int64_t ly, const int64_t *anchors, int64_t n_anchor,
                             int ragged_left, int ragged_right, const pecan_params_t *pp, int64_t **xs, int64_t **ys, double **ps) {
    StateMachine *sM = stateMachine5_construct(fiveState);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    p->threshold = pp->threshold; p->minDiagsBetweenTraceBack = pp->minDiagsBetweenTraceBack;
    p->traceBackDiagonals = pp->traceBackDiagonals; p->diagonalExpansion = pp->diagonalExpansion;
    stList *anchorPairs = stList_construct3(0, (void (*)(void *))stIntTuple_destruct);
    for (int64_t i = 0; i < n_anchor; ++i) stList_append(anchorPairs, stIntTuple_construct3(anchors[2 * i], anchors[2 * i + 1], p->diagonalExpansion));
    SymbolString sX = symbolString_construct(sx, lx), sY = symbolString_construct(sy, ly);
    sink_t s; memset(&s, 0, sizeof(s));
    s.band = band_construct(anchorPairs, lx, ly, p->diagonalExpansion); s.lXalY = lx + ly; s.threshold = p->threshold;
    getPosteriorProbsWithBanding(sM, anchorPairs, sX, sY, p, ragged_left, ragged_right, record_match_posteriors, &s);
    band_destruct(s.band);
    free(sX.sequence); free(sY.sequence);
    stList_destruct(anchorPairs);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    *xs = s.x; *ys = s.y; *ps = s.p;
    return s.n;
}

/* the reference's own integer triples through its public entry point (getAlignedPairsUsingAnchors,
 * pairwiseAligner.c) -- used to check the harness itself against the code path Cactus calls */
int64_t pecan_ref_aligned_pairs2(const char *sx, int64_t lx, const char *sy, int64_t ly, const int64_t *anchors, int64_t n_anchor,
                                 int ragged_left, int ragged_right, const pecan_params_t *pp, int64_t split_bigger, int64_t **trip) {
    StateMachine *sM = stateMachine5_construct(fiveState);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    if (split_bigger > 0) p->splitMatrixBiggerThanThis = split_bigger;
    p->threshold = pp->threshold; p->minDiagsBetweenTraceBack = pp->minDiagsBetweenTraceBack;
    p->traceBackDiagonals = pp->traceBackDiagonals; p->diagonalExpansion = pp->diagonalExpansion;
    stList *anchorPairs = stList_construct3(0, (void (*)(void *))stIntTuple_destruct);
    for (int64_t i = 0; i < n_anchor; ++i) stList_append(anchorPairs, stIntTuple_construct3(anchors[2 * i], anchors[2 * i + 1], p->diagonalExpansion));
    char *cx = malloc(lx + 1), *cy = malloc(ly + 1);
    memcpy(cx, sx, lx); cx[lx] = 0; memcpy(cy, sy, ly); cy[ly] = 0;
    stList *pairs = getAlignedPairsUsingAnchors(sM, cx, cy, anchorPairs, p, ragged_left, ragged_right);
    int64_t n = stList_length(pairs);
    int64_t *t = malloc(sizeof(int64_t) * 3 * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { stIntTuple *tp = stList_get(pairs, i); for (int k = 0; k < 3; ++k) t[3 * i + k] = stIntTuple_get(tp, k); }
    stList_destruct(pairs); stList_destruct(anchorPairs);
    free(cx); free(cy);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    *trip = t;
    return n;
}

int64_t pecan_ref_aligned_pairs(const char *sx, int64_t lx, const char *sy, int64_t ly, const int64_t *anchors, int64_t n_anchor,
                                int ragged_left, int ragged_right, const pecan_params_t *pp, int64_t **trip) {
    return pecan_ref_aligned_pairs2(sx, lx, sy, ly, anchors, n_anchor, ragged_left, ragged_right, pp, 0, trip);
}

/* makeAllPairwiseAlignments (multipleAligner.c:667-680): all pairs of n_seq fragments with the reference's default
 * PairwiseAlignmentParameters (MUM anchors for matrices > 500 x 500). tuples5 = n x 5 (score, seq1, pos1, seq2, pos2),
 * scores = npairs x 3 (similarity, seq1, seq2). */
int64_t pecan_ref_make_all_pairwise(int64_t n_seq, const char **seqs, const int64_t *left_end, const int64_t *right_end,
                                    int64_t **tuples5, int64_t **scores, int64_t *n_scores) {
    StateMachine *sM = stateMachine5_construct(fiveState);
    PairwiseAlignmentParameters *p = pairwiseAlignmentBandingParameters_construct();
    stList *seqFrags = stList_construct3(0, (void (*)(void *))seqFrag_destruct);
    for (int64_t i = 0; i < n_seq; ++i) stList_append(seqFrags, seqFrag_construct(seqs[i], left_end[i], right_end[i]));
    stList *sc = NULL;
    stList *mp = makeAllPairwiseAlignments(sM, seqFrags, p, &sc);
    int64_t n = stList_length(mp), m = stList_length(sc);
    int64_t *t = malloc(sizeof(int64_t) * 5 * (n > 0 ? n : 1)), *s = malloc(sizeof(int64_t) * 3 * (m > 0 ? m : 1));
    for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 5; ++k) t[5 * i + k] = stIntTuple_get(stList_get(mp, i), k);
    for (int64_t i = 0; i < m; ++i) for (int k = 0; k < 3; ++k) s[3 * i + k] = stIntTuple_get(stList_get(sc, i), k);
    stList_destruct(mp); stList_destruct(sc); stList_destruct(seqFrags);
    pairwiseAlignmentBandingParameters_destruct(p);
    stateMachine_destruct(sM);
    *tuples5 = t; *scores = s; *n_scores = m;
    return n;
}

void pecan_ref_free(void *p) { free(p); }
