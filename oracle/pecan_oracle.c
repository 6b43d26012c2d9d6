/*
 * pecan_oracle.c -- TEST INFRASTRUCTURE ONLY. Plain-C restatement of the cPecan pair-HMM path Cactus' BAR phase uses
 * when partialOrderAlignment="0" (SURVEY.md 8a rows a13/a14): banded forward / backward over x+y diagonals in log
 * space (double) with the reference's piecewise-cubic logAdd, intermediate tracebacks, posterior match probabilities.
 * Nothing in the product may link or call this file; tests/ use it as the checker next to the compiled reference
 * (oracle/_ref/libpecan_ref.so, oracle/pecan_ref_harness.c) that pins it.
 *
 * Follows (paths relative to /root/reference/submodules/cPecan/impl):
 *   band_construct / band_setCurrentDiagonal ........ pairwiseAligner.c:98-133, 193-244
 *   logAdd / lookup .................................. pairwiseAligner.c:297-317
 *   stateMachine5 constants, start/end probabilities . stateMachine.c:395-448, 482-521; emissions :269-292, N :351-366
 *   stateMachine5_cellCalculate (transition order) ... stateMachine.c:450-480
 *   diagonalCalculation (lower / middle / upper) ..... pairwiseAligner.c:619-634
 *   diagonalCalculationTotalProbability .............. pairwiseAligner.c:646-663
 *   getPosteriorProbsWithBanding ..................... pairwiseAligner.c:766-887
 *   posterior: exp(f_M + b_M - total), kept if >= threshold (pre-floor) ... pairwiseAligner.c:665-699
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double threshold;
    int64_t minDiagsBetweenTraceBack, traceBackDiagonals, diagonalExpansion;
} pecan_params_t;

#define NS 5
enum { M = 0, SX = 1, SY = 2, LX = 3, LY = 4 };
#define LZ (-INFINITY)

static const double T_MATCH_CONTINUE = -0.030064059121770816, T_MATCH_FROM_SHORT = -1.272871422049609,
                    T_MATCH_FROM_LONG = -5.673280173170473, T_SHORT_OPEN = -4.34381910900448,
                    T_SHORT_EXTEND = -0.3388262689231553, T_LONG_OPEN = -6.30810595366929,
                    T_LONG_EXTEND = -0.003442492794189331;
static const double E_MATCH = -2.1149196655034745, E_TRANSVERSION = -4.5691014376830479, E_TRANSITION = -3.9833860032220842,
                    E_GAP = -1.6094379124341003, E_GAP_N = -1.386294361, E_MATCH_N = -2.772588722;

static double lookup(double x) {
    if (x <= 1.00f) return ((-0.009350833524763f * x + 0.130659527668286f) * x + 0.498799810682272f) * x + 0.693203116424741f;
    if (x <= 2.50f) return ((-0.014532321752540f * x + 0.139942324101744f) * x + 0.495635523139337f) * x + 0.692140569840976f;
    if (x <= 4.50f) return ((-0.004605031767994f * x + 0.063427417320019f) * x + 0.695956496475118f) * x + 0.514272634594009f;
    return ((-0.000458661602210f * x + 0.009695946122598f) * x + 0.930734667215156f) * x + 0.168037164329057f;
}
static double logAdd(double x, double y) {
    if (x < y) return (x == LZ || y - x >= 7.5) ? y : lookup(y - x) + x;
    return (y == LZ || x - y >= 7.5) ? x : lookup(x - y) + y;
}
double oracle_pecan_logAdd(double x, double y) { return logAdd(x, y); }

static int sym(char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}
static double e_gap(int s) { return s == 4 ? E_GAP_N : E_GAP; }
static double e_match(int x, int y) {
    if (x == 4 || y == 4) return E_MATCH_N;
    if (x == y) return E_MATCH;
    return ((x ^ y) == 2) ? E_TRANSITION : E_TRANSVERSION;       /* A<->G (0,2), C<->T (1,3) are transitions */
}

/* ---- band ------------------------------------------------------------------------------------------------- */
static int64_t avoid_off_by_one(int64_t xay, int64_t xmy) { return (xay + xmy) % 2 == 0 ? xmy : xmy + 1; }
static int64_t bound(int64_t z, int64_t l) { return z < 0 ? 0 : (z > l ? l : z); }
static void set_diag(int64_t xay, int64_t xL, int64_t yL, int64_t xU, int64_t yU, int64_t *oL, int64_t *oR) {
    int64_t l = avoid_off_by_one(xay, xL - yL), r = avoid_off_by_one(xay, xU - yU);
    int64_t i;
    i = (xay + l) / 2; if (i < xL) l += 2 * (xL - i);
    i = (xay - l) / 2; if (yL < i) l += 2 * (i - yL);
    i = (xay + r) / 2; if (xU < i) r -= 2 * (i - xU);
    i = (xay - r) / 2; if (i < yU) r -= 2 * (yU - i);
    *oL = l; *oR = r;
}
/* diagonals xay = 0 .. lX+lY: [xmyL, xmyR]; anchors are 0-based (x, y) pairs */
void oracle_pecan_band(const int64_t *anchors, int64_t n_anchor, int64_t lX, int64_t lY, int64_t expansion, int64_t *xmyL, int64_t *xmyR) {
    int64_t ai = 0, xay = 0, pxay = 0, pxmy = 0, nxay = 0, nxmy = 0, xL = 0, yL = 0, xU = 0, yU = 0;
    while (xay <= lX + lY) {
        set_diag(xay, xL, yL, xU, yU, &xmyL[xay], &xmyR[xay]);
        if (nxay == xay++) {
            pxay = nxay; pxmy = nxmy;
            int64_t x = lX, y = lY;
            if (ai < n_anchor) { x = anchors[2 * ai] + 1; y = anchors[2 * ai + 1] + 1; ++ai; }
            nxay = x + y; nxmy = x - y;
            xL = bound((pxay + (pxmy - expansion)) / 2, lX);
            yL = bound((nxay - (nxmy - expansion)) / 2, lY);
            xU = bound((nxay + (nxmy + expansion)) / 2, lX);
            yU = bound((pxay - (pxmy + expansion)) / 2, lY);
        }
    }
}

/* ---- dp storage ------------------------------------------------------------------------------------------- */
typedef struct { int64_t n; const int64_t *L, *R; int64_t *off; double *v; } mat_t;
static void mat_init(mat_t *m, int64_t ndiag, const int64_t *L, const int64_t *R) {
    m->n = ndiag; m->L = L; m->R = R; m->off = malloc(sizeof(int64_t) * (ndiag + 2));
    int64_t o = 0;
    for (int64_t d = 0; d <= ndiag; ++d) { m->off[d] = o; o += ((R[d] - L[d]) / 2 + 1) * NS; }
    m->off[ndiag + 1] = o;
    m->v = malloc(sizeof(double) * (o > 0 ? o : 1));
}
static double *cell(const mat_t *m, int64_t d, int64_t xmy) {
    if (d < 0 || d > m->n || xmy < m->L[d] || xmy > m->R[d]) return NULL;
    return m->v + m->off[d] + ((xmy - m->L[d]) / 2) * NS;
}
static void fill(mat_t *m, int64_t d, const double *s) {
    for (int64_t xmy = m->L[d]; xmy <= m->R[d]; xmy += 2) { double *c = cell(m, d, xmy); for (int k = 0; k < NS; ++k) c[k] = s ? s[k] : LZ; }
}

/* stateMachine5_cellCalculate in its order of doTransition calls; fwd: to += from, bwd: from += to */
static void cell_calc(double *cur, double *lower, double *middle, double *upper, int cX, int cY, int backward) {
#define TR(fromc, f, t, e) do { if (backward) (fromc)[f] = logAdd((fromc)[f], cur[t] + (e)); else cur[t] = logAdd(cur[t], (fromc)[f] + (e)); } while (0)
    if (lower) {
        const double eP = e_gap(cX);
        TR(lower, M, SX, eP + T_SHORT_OPEN); TR(lower, SX, SX, eP + T_SHORT_EXTEND);
        TR(lower, M, LX, eP + T_LONG_OPEN); TR(lower, LX, LX, eP + T_LONG_EXTEND);
    }
    if (middle) {
        const double eP = e_match(cX, cY);
        TR(middle, M, M, eP + T_MATCH_CONTINUE); TR(middle, SX, M, eP + T_MATCH_FROM_SHORT); TR(middle, SY, M, eP + T_MATCH_FROM_SHORT);
        TR(middle, LX, M, eP + T_MATCH_FROM_LONG); TR(middle, LY, M, eP + T_MATCH_FROM_LONG);
    }
    if (upper) {
        const double eP = e_gap(cY);
        TR(upper, M, SY, eP + T_SHORT_OPEN); TR(upper, SY, SY, eP + T_SHORT_EXTEND);
        TR(upper, M, LY, eP + T_LONG_OPEN); TR(upper, LY, LY, eP + T_LONG_EXTEND);
    }
#undef TR
}
static int xsym(const int *sx, int64_t xay, int64_t xmy) { int64_t x = (xay + xmy) / 2; return x > 0 ? sx[x - 1] : 4; }
static int ysym(const int *sy, int64_t xay, int64_t xmy) { int64_t y = (xay - xmy) / 2; return y > 0 ? sy[y - 1] : 4; }
/* diagonalCalculation: cur = diagonal d of `to`, M1 = d-1, M2 = d-2 of `from` (same matrix except for the total-probability trick) */
static void diag_calc(mat_t *cm, int64_t d, mat_t *m1m, int has_m1, mat_t *m2m, int has_m2, const int *sx, const int *sy, int backward) {
    for (int64_t xmy = cm->L[d]; xmy <= cm->R[d]; xmy += 2) {
        double *cur = cell(cm, d, xmy);
        double *lower = has_m1 ? cell(m1m, d - 1, xmy - 1) : NULL, *middle = has_m2 ? cell(m2m, d - 2, xmy) : NULL,
               *upper = has_m1 ? cell(m1m, d - 1, xmy + 1) : NULL;
        cell_calc(cur, lower, middle, upper, xsym(sx, d, xmy), ysym(sy, d, xmy), backward);
    }
}
static double dot(const mat_t *a, const mat_t *b, int64_t d) {
    double tot = LZ;
    for (int64_t xmy = a->L[d]; xmy <= a->R[d]; xmy += 2) {
        const double *c1 = cell(a, d, xmy), *c2 = cell(b, d, xmy);
        double t = c1[0] + c2[0];
        for (int k = 1; k < NS; ++k) t = logAdd(t, c1[k] + c2[k]);
        tot = logAdd(tot, t);
    }
    return tot;
}

typedef struct { int64_t *x, *y; double *p; int64_t n, m; } sink_t;
static void push(sink_t *s, int64_t x, int64_t y, double p) {
    if (s->n == s->m) { s->m = s->m ? 2 * s->m : 1024; s->x = realloc(s->x, 8 * s->m); s->y = realloc(s->y, 8 * s->m); s->p = realloc(s->p, 8 * s->m); }
    s->x[s->n] = x; s->y[s->n] = y; s->p[s->n] = p; s->n++;
}

int64_t oracle_pecan_posteriors(const char *csx, int64_t lX, const char *csy, int64_t lY, const int64_t *anchors, int64_t n_anchor,
                                int ragged_left, int ragged_right, const pecan_params_t *pp, int64_t **xs, int64_t **ys, double **ps) {
    sink_t out; memset(&out, 0, sizeof(out));
    const int64_t D = lX + lY;
    *xs = NULL; *ys = NULL; *ps = NULL;
    if (D == 0) return 0;
    int *sx = malloc(sizeof(int) * (lX + 1)), *sy = malloc(sizeof(int) * (lY + 1));
    for (int64_t i = 0; i < lX; ++i) sx[i] = sym(csx[i]);
    for (int64_t i = 0; i < lY; ++i) sy[i] = sym(csy[i]);
    int64_t *L = malloc(8 * (D + 1)), *R = malloc(8 * (D + 1));
    oracle_pecan_band(anchors, n_anchor, lX, lY, pp->diagonalExpansion, L, R);
    mat_t F, B, T;
    mat_init(&F, D, L, R); mat_init(&B, D, L, R); mat_init(&T, D, L, R);
    const double start[NS] = {0, LZ, LZ, LZ, LZ}, rstart[NS] = {LZ, LZ, LZ, 0, 0};
    const double endp[NS] = {T_MATCH_CONTINUE, T_MATCH_FROM_SHORT, T_MATCH_FROM_SHORT, T_MATCH_FROM_LONG, T_MATCH_FROM_LONG};
    const double rend[NS] = {T_LONG_OPEN, T_LONG_OPEN, T_LONG_OPEN, T_LONG_EXTEND, T_LONG_EXTEND};
    fill(&F, 0, ragged_left ? rstart : start);
    int64_t tracedBackTo = 0;
    for (int64_t d = 1; d <= D; ++d) {
        fill(&F, d, NULL);
        diag_calc(&F, d, &F, 1, &F, d >= 2, sx, sy, 0);
        const int atEnd = d == D;
        const int tbPoint = d >= tracedBackTo + pp->minDiagsBetweenTraceBack && (R[d] - L[d]) / 2 + 1 <= pp->diagonalExpansion * 2 + 1;
        if (!(atEnd || tbPoint)) continue;
        fill(&B, d, (atEnd && ragged_right) ? rend : endp);
        if (d > tracedBackTo + 1) fill(&B, d - 1, NULL);
        const int64_t tracedBackFrom = d - (atEnd ? 0 : pp->traceBackDiagonals + 1);
        double total = LZ; int64_t ncalc = 0;
        for (int64_t d2 = d; d2 > tracedBackTo; --d2) {
            if (d2 > tracedBackTo + 2) fill(&B, d2 - 2, NULL);
            /* the backward diagonal d2-2 only exists (was created just above) beyond tracedBackTo + 2 */
            if (d2 > tracedBackTo + 1) diag_calc(&B, d2, &B, 1, &B, d2 > tracedBackTo + 2, sx, sy, 1);
            if (d2 <= tracedBackFrom) {
                if (ncalc++ % 10 == 0) {
                    double t = dot(&F, &B, d2);
                    if (d2 + 1 <= D && d2 - 1 >= 0) {          /* matches through d2: forward d2-1 -> match -> backward d2+1 */
                        /* (the backward diagonal d2+1 exists unless d2 is the last diagonal walked from) */
                        if (d2 + 1 <= d) {
                            fill(&T, d2 + 1, NULL);
                            diag_calc(&T, d2 + 1, &F, 0, &F, 1, sx, sy, 0);
                            t = logAdd(t, dot(&T, &B, d2 + 1));
                        }
                    }
                    total = t;
                }
                for (int64_t xmy = L[d2]; xmy <= R[d2]; xmy += 2) {
                    const int64_t x = (d2 + xmy) / 2, y = (d2 - xmy) / 2;
                    if (x > 0 && y > 0) {
                        const double post = exp((cell(&F, d2, xmy)[M] + cell(&B, d2, xmy)[M]) - total);
                        if (post >= pp->threshold) push(&out, x - 1, y - 1, post);
                    }
                }
            }
        }
        tracedBackTo = tracedBackFrom;
    }
    free(F.off); free(F.v); free(B.off); free(B.v); free(T.off); free(T.v); free(L); free(R); free(sx); free(sy);
    *xs = out.x; *ys = out.y; *ps = out.p;
    return out.n;
}

void oracle_pecan_free(void *p) { free(p); }

/* ---- getAlignedPairsUsingAnchors (pairwiseAligner.c:1477-1495): split at large anchor gaps (getSplitPoints :1241-1292),
 * posteriors of every sub-matrix (:1308-1363), floor(p * PAIR_ALIGNMENT_PROB_1) (addPosteriorProb :665-674), coordinates
 * shifted back (:1294-1306, 1457-1464). Output: n x 3 int64 (score, x, y) in the reference's order (each region's
 * pairs in REVERSE order of emission, because they are moved over with stList_pop). ------------------- */
typedef struct { int64_t *v; int64_t n, m; } vec_t;
static void vpush4(vec_t *s, int64_t a, int64_t b, int64_t c, int64_t d) {
    if (s->n + 4 > s->m) { s->m = s->m ? 2 * s->m : 64; s->v = realloc(s->v, 8 * s->m); }
    s->v[s->n++] = a; s->v[s->n++] = b; s->v[s->n++] = c; s->v[s->n++] = d;
}
static int split_p(int64_t *x1, int64_t *y1, int64_t x2, int64_t y2, int64_t x3, int64_t y3, vec_t *sp, int64_t bigger, int skip) {
    int64_t lX2 = x3 - x2, lY2 = y3 - y2;
    if (lX2 * lY2 > bigger) {
        int64_t maxLen = sqrt(bigger);
        int64_t hX = lX2 / 2 > maxLen ? maxLen : lX2 / 2, hY = lY2 / 2 > maxLen ? maxLen : lY2 / 2;
        if (!skip) vpush4(sp, *x1, *y1, x2 + hX, y2 + hY);
        *x1 = x3 - hX; *y1 = y3 - hY;
        return 1;
    }
    return 0;
}
/* returns the number of split regions; *out = malloc'd n x 4 (x1, y1, x2, y2) */
int64_t oracle_pecan_split_points(const int64_t *anchors, int64_t n_anchor, int64_t lX, int64_t lY, int64_t bigger, int ragged_left, int ragged_right, int64_t **out) {
    vec_t sp; memset(&sp, 0, sizeof(sp));
    int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    for (int64_t i = 0; i < n_anchor; ++i) {
        int64_t x3 = anchors[2 * i], y3 = anchors[2 * i + 1];
        split_p(&x1, &y1, x2, y2, x3, y3, &sp, bigger, ragged_left && i == 0);
        x2 = x3 + 1; y2 = y3 + 1;
    }
    if (!split_p(&x1, &y1, x2, y2, lX, lY, &sp, bigger, ragged_left && n_anchor == 0) || !ragged_right) vpush4(&sp, x1, y1, lX, lY);
    *out = sp.v;
    return sp.n / 4;
}

int64_t oracle_pecan_aligned_pairs(const char *csx, int64_t lX, const char *csy, int64_t lY, const int64_t *anchors, int64_t n_anchor,
                                   int ragged_left, int ragged_right, const pecan_params_t *pp, int64_t split_bigger, int64_t **trip, double **post) {
    int64_t *sp = NULL;
    const int64_t nsp = oracle_pecan_split_points(anchors, n_anchor, lX, lY, split_bigger, ragged_left, ragged_right, &sp);
    int64_t n = 0, m = 0, *t = NULL, j = 0; double *po = NULL;
    for (int64_t i = 0; i < nsp; ++i) {
        const int64_t x1 = sp[4 * i], y1 = sp[4 * i + 1], x2 = sp[4 * i + 2], y2 = sp[4 * i + 3];
        int64_t *sub = malloc(16 * (n_anchor > 0 ? n_anchor : 1)), ns = 0;
        while (j < n_anchor) {
            const int64_t x = anchors[2 * j], y = anchors[2 * j + 1];
            if (x + y >= x2 + y2) break;
            sub[2 * ns] = x - x1; sub[2 * ns + 1] = y - y1; ++ns; ++j;
        }
        int64_t *xs, *ys; double *ps;
        const int64_t k = oracle_pecan_posteriors(csx + x1, x2 - x1, csy + y1, y2 - y1, sub, ns, ragged_left || i > 0, ragged_right || i < nsp - 1, pp, &xs, &ys, &ps);
        if (n + k > m) { m = 2 * (n + k) + 16; t = realloc(t, 24 * m); po = realloc(po, 8 * m); }
        for (int64_t q = k - 1; q >= 0; --q) {      /* alignedPairCoordinateCorrectionFn POPS the sub-list: reversed per region (:1457-1464) */
            double p = ps[q];
            po[n] = p;
            if (p > 1.0) p = 1.0;
            t[3 * n] = (int64_t)floor(p * 10000000.0); t[3 * n + 1] = xs[q] + x1; t[3 * n + 2] = ys[q] + y1; ++n;
        }
        free(xs); free(ys); free(ps); free(sub);
    }
    free(sp);
    if (!t) { t = malloc(24); po = malloc(8); }
    *trip = t; *post = po;
    return n;
}
