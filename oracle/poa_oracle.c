/*
 * poa_oracle.c -- TEST INFRASTRUCTURE ONLY (see poa_oracle.h). Scalar plain-C restatement of what
 * abpoa_msa() computes under the parameters Cactus' BAR phase uses. Nothing here is SIMD, nothing is
 * shared with the CUDA product; every function cites the reference lines it follows
 * (paths relative to /root/reference/submodules/abPOA/src unless stated otherwise).
 *
 * Fixed settings (bar/impl/poaBarAligner.c:24-112): align_mode=GLOBAL, gap_mode=CONVEX (both opens > 0),
 * wb >= 0 (adaptive band on), disable_seeding=1, use_read_ids=1, out_msa=1, out_cons=0, inc_path_score=0,
 * put_gap_on_right=put_gap_at_end=0, rev_cigar=0, amb_strand=0, use_qv=0, m=5.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <limits.h>
#include "poa_oracle.h"

#define SRC 0
#define SINK 1
#define OP_M 0x1
#define OP_E1 0x2
#define OP_E2 0x4
#define OP_E 0x6
#define OP_F1 0x8
#define OP_F2 0x10
#define OP_F 0x18
#define OP_ALL 0x1f
#define CMATCH 0
#define CINS 1
#define CDEL 2

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX3(a, b, c) MAX2(MAX2(a, b), c)

static void die(const char *msg) { fprintf(stderr, "poa_oracle: %s\n", msg); exit(1); }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) die("out of memory"); return p; }
static void *xrealloc(void *q, size_t s) { void *p = realloc(q, s ? s : 1); if (!p) die("out of memory"); return p; }

/* ------------------------------------------------------------------------------------------------
 * graph (include/abpoa.h:96-116, abpoa_graph.c)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t base;
    int in_n, in_m, *in_id, *in_w;
    int out_n, out_m, *out_id, *out_w;
    uint64_t *rid;                 /* [out_m * W] read-id bitset per out edge (abpoa_graph.c:525-544) */
    int aln_n, aln_m, *aln_id;     /* aligned (same MSA column) nodes */
} node_t;

typedef struct {
    node_t *node; int node_n, node_m;
    int W;                         /* words per read-id set: 1 + ((tot_read_n-1) >> 6), abpoa_graph.c:692 */
    int *index_to_node, *node_to_index, *remain, *maxL, *maxR, *msa_rank;
    int arr_m;
} graph_t;

static void graph_init(graph_t *g, int tot_read_n) {
    memset(g, 0, sizeof(*g));
    g->W = 1 + ((tot_read_n - 1) >> 6);
    g->node_m = 1024; g->node = (node_t *)xcalloc(g->node_m, sizeof(node_t));
    g->node_n = 2;                                     /* SRC, SINK (abpoa_graph.c:103-113) */
}

static void graph_free(graph_t *g) {
    for (int i = 0; i < g->node_n; ++i) {
        free(g->node[i].in_id); free(g->node[i].in_w); free(g->node[i].out_id); free(g->node[i].out_w);
        free(g->node[i].rid); free(g->node[i].aln_id);
    }
    free(g->node); free(g->index_to_node); free(g->node_to_index); free(g->remain);
    free(g->maxL); free(g->maxR); free(g->msa_rank);
}

/* abpoa_add_graph_node, abpoa_graph.c:471-478 */
static int add_node(graph_t *g, uint8_t base) {
    if (g->node_n == g->node_m) {
        g->node = (node_t *)xrealloc(g->node, sizeof(node_t) * g->node_m * 2);
        memset(g->node + g->node_m, 0, sizeof(node_t) * g->node_m);
        g->node_m *= 2;
    }
    g->node[g->node_n].base = base;
    return g->node_n++;
}

/* abpoa_add_graph_edge with w=1, add_read_id=1, add_read_weight=0 (abpoa_graph.c:480-556) */
static void add_edge(graph_t *g, int from, int to, int check_edge, int read_id) {
    node_t *f = &g->node[from], *t = &g->node[to];
    int exist = 0, out_i = -1;
    if (check_edge) {
        for (int i = 0; i < t->in_n; ++i) if (t->in_id[i] == from) { t->in_w[i] += 1; break; }
        for (int i = 0; i < f->out_n; ++i) if (f->out_id[i] == to) { f->out_w[i] += 1; exist = 1; out_i = i; break; }
    }
    if (!exist) {
        if (t->in_n == t->in_m) {
            t->in_m = t->in_m ? t->in_m * 2 : 2;
            t->in_id = (int *)xrealloc(t->in_id, sizeof(int) * t->in_m);
            t->in_w = (int *)xrealloc(t->in_w, sizeof(int) * t->in_m);
        }
        t->in_id[t->in_n] = from; t->in_w[t->in_n] = 1; t->in_n++;
        if (f->out_n == f->out_m) {
            int m = f->out_m ? f->out_m * 2 : 2;
            f->out_id = (int *)xrealloc(f->out_id, sizeof(int) * m);
            f->out_w = (int *)xrealloc(f->out_w, sizeof(int) * m);
            f->rid = (uint64_t *)xrealloc(f->rid, sizeof(uint64_t) * (size_t)m * g->W);
            memset(f->rid + (size_t)f->out_m * g->W, 0, sizeof(uint64_t) * (size_t)(m - f->out_m) * g->W);
            f->out_m = m;
        }
        f->out_id[f->out_n] = to; f->out_w[f->out_n] = 1; out_i = f->out_n; f->out_n++;
    }
    f->rid[(size_t)out_i * g->W + (read_id >> 6)] |= 1ULL << (read_id & 63);   /* abpoa_graph.c:465-469 */
}

static void aln_push(node_t *n, int id) {
    if (n->aln_n == n->aln_m) { n->aln_m = n->aln_m ? n->aln_m * 2 : 4; n->aln_id = (int *)xrealloc(n->aln_id, sizeof(int) * n->aln_m); }
    n->aln_id[n->aln_n++] = id;
}

/* abpoa_add_graph_aligned_node, abpoa_graph.c:455-463 */
static void add_aligned(graph_t *g, int node_id, int new_id) {
    node_t *n = &g->node[node_id];
    for (int i = 0; i < n->aln_n; ++i) {
        aln_push(&g->node[n->aln_id[i]], new_id);
        aln_push(&g->node[new_id], n->aln_id[i]);
    }
    aln_push(&g->node[node_id], new_id);
    aln_push(&g->node[new_id], node_id);
}

/* abpoa_get_aligned_id, abpoa_graph.c:439-448 */
static int aligned_with_base(graph_t *g, int node_id, uint8_t base) {
    node_t *n = &g->node[node_id];
    for (int i = 0; i < n->aln_n; ++i) if (g->node[n->aln_id[i]].base == base) return n->aln_id[i];
    return -1;
}

/* abpoa_BFS_set_node_index, abpoa_graph.c:221-266: FIFO BFS from SRC; a node is enqueued when its
 * in-degree reaches 0 AND all nodes aligned to it are at 0 too, and then drags those along. */
static void bfs_index(graph_t *g) {
    int n = g->node_n, head = 0, tail = 0, index = 0;
    int *indeg = (int *)xcalloc(n, sizeof(int)), *q = (int *)xcalloc(n, sizeof(int));
    for (int i = 0; i < n; ++i) indeg[i] = g->node[i].in_n;
    q[tail++] = SRC;
    while (head < tail) {
        int cur = q[head++];
        g->index_to_node[index] = cur; g->node_to_index[cur] = index++;
        if (cur == SINK) { free(indeg); free(q); return; }
        node_t *c = &g->node[cur];
        for (int i = 0; i < c->out_n; ++i) {
            int o = c->out_id[i];
            if (--indeg[o] == 0) {
                node_t *on = &g->node[o]; int ok = 1;
                for (int j = 0; j < on->aln_n; ++j) if (indeg[on->aln_id[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                q[tail++] = o;
                for (int j = 0; j < on->aln_n; ++j) q[tail++] = on->aln_id[j];
            }
        }
    }
    die("Failed to set node index");
}

/* abpoa_sort_in_out_ids, abpoa_graph.c:192-219: the exact (unstable) exchange sort, weight descending */
static void sort_edges(graph_t *g) {
    for (int i = 0; i < g->node_n; ++i) {
        node_t *n = &g->node[i];
        for (int j = 0; j < n->in_n - 1; ++j)
            for (int k = j + 1; k < n->in_n; ++k)
                if (n->in_w[j] < n->in_w[k]) {
                    int t = n->in_id[j]; n->in_id[j] = n->in_id[k]; n->in_id[k] = t;
                    t = n->in_w[j]; n->in_w[j] = n->in_w[k]; n->in_w[k] = t;
                }
        for (int j = 0; j < n->out_n - 1; ++j)
            for (int k = j + 1; k < n->out_n; ++k)
                if (n->out_w[j] < n->out_w[k]) {
                    int t = n->out_id[j]; n->out_id[j] = n->out_id[k]; n->out_id[k] = t;
                    t = n->out_w[j]; n->out_w[j] = n->out_w[k]; n->out_w[k] = t;
                    for (int w = 0; w < g->W; ++w) {
                        uint64_t r = n->rid[(size_t)j * g->W + w];
                        n->rid[(size_t)j * g->W + w] = n->rid[(size_t)k * g->W + w];
                        n->rid[(size_t)k * g->W + w] = r;
                    }
                }
    }
}

/* abpoa_BFS_set_node_remain, abpoa_graph.c:268-309: reverse BFS from SINK (remain=-1);
 * remain[v] = remain[first max-weight out neighbour] + 1 */
static void bfs_remain(graph_t *g) {
    int n = g->node_n, head = 0, tail = 0;
    int *outdeg = (int *)xcalloc(n, sizeof(int)), *q = (int *)xcalloc(n, sizeof(int));
    for (int i = 0; i < n; ++i) { outdeg[i] = g->node[i].out_n; g->remain[i] = 0; }
    q[tail++] = SINK; g->remain[SINK] = -1;
    while (head < tail) {
        int cur = q[head++]; node_t *c = &g->node[cur];
        if (cur != SINK) {
            int max_w = -1, max_id = -1;
            for (int i = 0; i < c->out_n; ++i) if (c->out_w[i] > max_w) { max_w = c->out_w[i]; max_id = c->out_id[i]; }
            g->remain[cur] = g->remain[max_id] + 1;
        }
        if (cur == SRC) { free(outdeg); free(q); return; }
        for (int i = 0; i < c->in_n; ++i) if (--outdeg[c->in_id[i]] == 0) q[tail++] = c->in_id[i];
    }
    die("Failed to set node remain");
}

/* abpoa_topological_sort, abpoa_graph.c:322-357 */
static void topo_sort(graph_t *g) {
    int n = g->node_n;
    if (n > g->arr_m) {
        g->arr_m = n * 2;
        g->index_to_node = (int *)xrealloc(g->index_to_node, sizeof(int) * g->arr_m);
        g->node_to_index = (int *)xrealloc(g->node_to_index, sizeof(int) * g->arr_m);
        g->remain = (int *)xrealloc(g->remain, sizeof(int) * g->arr_m);
        g->maxL = (int *)xrealloc(g->maxL, sizeof(int) * g->arr_m);
        g->maxR = (int *)xrealloc(g->maxR, sizeof(int) * g->arr_m);
        g->msa_rank = (int *)xrealloc(g->msa_rank, sizeof(int) * g->arr_m);
    }
    bfs_index(g);
    sort_edges(g);
    for (int i = 0; i < n; ++i) { g->maxR[i] = 0; g->maxL[i] = n; }
    bfs_remain(g);
}

/* ------------------------------------------------------------------------------------------------
 * guide tree (abpoa_seed.c:36-46, 85-156, 232-325, 705-722)
 * ---------------------------------------------------------------------------------------------- */
static uint64_t hash64(uint64_t key, uint64_t mask) {       /* abpoa_seed.c:36-46 (minimap2's invertible hash) */
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

typedef struct { uint64_t x, y; } mm_t;
typedef struct { mm_t *a; size_t n, m; } mm_v;
static void mm_push(mm_v *v, mm_t e) {
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 256; v->a = (mm_t *)xrealloc(v->a, sizeof(mm_t) * v->m); }
    v->a[v->n++] = e;
}

/* (w,k)-minimizers of the forward strand, no homopolymer compression (mm_sketch with is_hpc=0,
 * both_strand=0; abpoa_seed.c:85-156). x = hash<<8 | span, y = rid<<32 | pos<<1. */
static void sketch(const uint8_t *s, int len, int w, int k, uint32_t rid, mm_v *out) {
    uint64_t mask = (1ULL << 2 * k) - 1, kmer = 0;
    mm_t buf[256], min = {UINT64_MAX, UINT64_MAX};
    int l = 0, buf_pos = 0, min_pos = 0;
    for (int j = 0; j < w; ++j) buf[j].x = buf[j].y = UINT64_MAX;
    for (int i = 0; i < len; ++i) {
        int c = s[i];
        mm_t info = {UINT64_MAX, UINT64_MAX};
        if (c < 4) {
            int span = l + 1 < k ? l + 1 : k;
            kmer = (kmer << 2 | (uint64_t)c) & mask;
            ++l;
            if (l >= k) { info.x = hash64(kmer, mask) << 8 | (uint64_t)span; info.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1; }
        } else l = 0;
        buf[buf_pos] = info;
        if (l == w + k - 1 && min.x != UINT64_MAX) {        /* first full window: flush k-mers tying with min */
            for (int j = buf_pos + 1; j < w; ++j) if (min.x == buf[j].x && buf[j].y != min.y) mm_push(out, buf[j]);
            for (int j = 0; j < buf_pos; ++j) if (min.x == buf[j].x && buf[j].y != min.y) mm_push(out, buf[j]);
        }
        if (info.x <= min.x) {                               /* new minimum */
            if (l >= w + k && min.x != UINT64_MAX) mm_push(out, min);
            min = info; min_pos = buf_pos;
        } else if (buf_pos == min_pos) {                     /* old minimum left the window */
            if (l >= w + k - 1 && min.x != UINT64_MAX) mm_push(out, min);
            min.x = UINT64_MAX;
            for (int j = buf_pos + 1; j < w; ++j) if (min.x >= buf[j].x) { min = buf[j]; min_pos = j; }
            for (int j = 0; j <= buf_pos; ++j) if (min.x >= buf[j].x) { min = buf[j]; min_pos = j; }
            if (l >= w + k - 1 && min.x != UINT64_MAX) {
                for (int j = buf_pos + 1; j < w; ++j) if (min.x == buf[j].x && min.y != buf[j].y) mm_push(out, buf[j]);
                for (int j = 0; j <= buf_pos; ++j) if (min.x == buf[j].x && min.y != buf[j].y) mm_push(out, buf[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min.x != UINT64_MAX) mm_push(out, min);
}

static int mm_cmp_x(const void *a, const void *b) {
    uint64_t x = ((const mm_t *)a)->x, y = ((const mm_t *)b)->x;
    return x < y ? -1 : x > y;
}

/* abpoa_build_guide_tree_partition (abpoa_seed.c:705-722) with seeding disabled: identity order unless
 * progressive && n_seq > 2, then abpoa_build_guide_tree (abpoa_seed.c:232-325): min-count Jaccard on
 * minimizer multisets, greedy order by summed similarity to the sequences already chosen. */
static void guide_tree(const oracle_params_t *p, uint8_t **seqs, const int *lens, int n, int *order) {
    for (int i = 0; i < n; ++i) order[i] = i;
    if (!(p->progressive_poa && n > 2)) return;
    mm_v mm = {0, 0, 0};
    for (int i = 0; i < n; ++i) sketch(seqs[i], lens[i], p->w, p->k, (uint32_t)i, &mm);
    if (mm.n == 0) { free(mm.a); return; }
    qsort(mm.a, mm.n, sizeof(mm_t), mm_cmp_x);               /* only the grouping by x matters */
    int *hit = (int *)xcalloc((size_t)n * (n + 1) / 2, sizeof(int)), *cnt = (int *)xcalloc(n, sizeof(int));
    for (size_t s = 0; s < mm.n;) {
        size_t e = s; memset(cnt, 0, sizeof(int) * n);
        while (e < mm.n && mm.a[e].x == mm.a[s].x) {
            int r = (int)(mm.a[e].y >> 32); ++cnt[r]; ++hit[(size_t)r * (r + 1) / 2 + r]; ++e;
        }
        for (int r1 = 0; r1 < n - 1; ++r1) for (int r2 = r1 + 1; r2 < n; ++r2)
            hit[(size_t)r2 * (r2 + 1) / 2 + r1] += MIN2(cnt[r1], cnt[r2]);
        s = e;
    }
    double *jac = (double *)xcalloc((size_t)n * (n - 1) / 2, sizeof(double)), max_jac = -1.0, jc;
    int max_i = -1, max_j = -1;
    for (int i = 1; i < n; ++i) for (int j = 0; j < i; ++j) {
        int shared = hit[(size_t)i * (i + 1) / 2 + j];
        int tot = hit[(size_t)i * (i + 1) / 2 + i] + hit[(size_t)j * (j + 1) / 2 + j] - shared;
        if (tot == 0) jc = 0; else if (tot < 0) { die("guide tree (1)"); jc = 0; } else jc = (0.0 + shared) / tot;
        jac[(size_t)i * (i - 1) / 2 + j] = jc;
        if (jc > max_jac) { max_jac = jc; max_i = i; max_j = j; }
    }
    int n_in = 2; order[0] = max_j; order[1] = max_i;
    while (n_in < n) {
        max_jac = -1.0; max_i = n;
        for (int r1 = 0; r1 < n; ++r1) {
            jc = 0.0;
            for (int i = 0; i < n_in; ++i) {
                int r2 = order[i];
                if (r1 == r2) { jc = -1.0; break; }
                else if (r1 > r2) jc += jac[(size_t)r1 * (r1 - 1) / 2 + r2];
                else jc += jac[(size_t)r2 * (r2 - 1) / 2 + r1];
            }
            if (jc > max_jac) { max_jac = jc; max_i = r1; }
        }
        if (max_i == n) die("guide tree (2)");
        order[n_in++] = max_i;
    }
    free(hit); free(cnt); free(jac); free(mm.a);
}

/* ------------------------------------------------------------------------------------------------
 * banded convex-gap sequence-to-graph DP + traceback (abpoa_align_simd.c)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int n_rows;                  /* rows 0..n_rows-1 = graph indices 0..node_n-2 (SINK row is never filled) */
    int *beg, *end;              /* dp_beg/dp_end per row */
    int64_t *off;                /* start of the row's band in each plane */
    int32_t *H, *E1, *E2, *F1, *F2;
    int32_t inf_min;
} dp_t;

static inline int32_t cell(const dp_t *d, const int32_t *plane, int row, int j) {
    if (j < d->beg[row] || j > d->end[row]) return d->inf_min;   /* out-of-band lanes hold inf_min (:1035-1036) */
    return plane[d->off[row] + (j - d->beg[row])];
}

typedef struct { uint64_t *c; int n, m; } cigar_t;
/* abpoa_push_cigar, abpoa_align.h:58-78 */
static void push_cigar(cigar_t *cg, int op, int len, int32_t node_id, int32_t query_id) {
    uint64_t l = (uint64_t)len;
    if (cg->n == 0 || op != CINS || op != (int)(cg->c[cg->n - 1] & 0xf)) {
        if (cg->n == cg->m) { cg->m = cg->m ? cg->m << 1 : 4; cg->c = (uint64_t *)xrealloc(cg->c, sizeof(uint64_t) * cg->m); }
        uint64_t n_id = (uint64_t)(int64_t)node_id, q_id = (uint64_t)(int64_t)query_id;
        if (op == CMATCH) cg->c[cg->n++] = n_id << 34 | q_id << 4 | (uint64_t)op;
        else if (op == CINS) cg->c[cg->n++] = q_id << 34 | l << 4 | (uint64_t)op;
        else cg->c[cg->n++] = n_id << 34 | l << 4 | (uint64_t)op;
    } else cg->c[cg->n - 1] += l << 4;
}

/* simd_abpoa_align_sequence_to_subgraph(SRC, SINK) for the convex-gap global banded case
 * (abpoa_align_simd.c:1250-1332 -> :1201-1231). Whole graph => every index is in index_map. */
static int align_to_graph(const oracle_params_t *p, graph_t *g, const uint8_t *q, int L, cigar_t *cg,
                          int32_t *best_score_out, dp_t *d) {
    const int *mat = p->mat;
    const int32_t o1 = p->gap_open1, e1 = p->gap_ext1, o2 = p->gap_open2, e2 = p->gap_ext2;
    const int32_t oe1 = o1 + e1, oe2 = o2 + e2;
    int max_mat = 0, min_mis = 0;
    for (int i = 0; i < 25; ++i) { if (mat[i] > max_mat) max_mat = mat[i]; if (-mat[i] > min_mis) min_mis = -mat[i]; }
    const int R = g->node_n - 1;                              /* rows to fill: index 0 .. end_index-1 (:1205) */
    const int gn = g->node_n;
    /* lane count the reference would pick: int16 (16 lanes at AVX2) or int32 (8 lanes), :1293-1302;
     * it leaks into dp_beg through the lane-group snap at :957-959 */
    int len = L > gn ? L : gn;
    int max_score = MAX2(L * max_mat, len * e1 + o1);
    const int pn = (max_score <= INT16_MAX - min_mis - oe1 - oe2) ? 16 : 8;
    const int32_t inf_min = MAX3(INT32_MIN + min_mis, INT32_MIN + oe1, INT32_MIN + oe2) + 512 * MAX2(e1, e2);  /* :1299 */
    const int w = p->wb + (int)(p->wf * L);                  /* :474 */

    d->n_rows = R; d->inf_min = inf_min;
    d->beg = (int *)xcalloc(R, sizeof(int)); d->end = (int *)xcalloc(R, sizeof(int));
    d->off = (int64_t *)xcalloc(R + 1, sizeof(int64_t));
    size_t cap = (size_t)1 << 20, used = 0;
    d->H = (int32_t *)xcalloc(cap, 4); d->E1 = (int32_t *)xcalloc(cap, 4); d->E2 = (int32_t *)xcalloc(cap, 4);
    d->F1 = (int32_t *)xcalloc(cap, 4); d->F2 = (int32_t *)xcalloc(cap, 4);
#define ENSURE(nw) do { if (used + (size_t)(nw) > cap) { while (used + (size_t)(nw) > cap) cap *= 2; \
        d->H = (int32_t *)xrealloc(d->H, cap * 4); d->E1 = (int32_t *)xrealloc(d->E1, cap * 4); d->E2 = (int32_t *)xrealloc(d->E2, cap * 4); \
        d->F1 = (int32_t *)xrealloc(d->F1, cap * 4); d->F2 = (int32_t *)xrealloc(d->F2, cap * 4); } } while (0)

    const int rem_sink = g->remain[SINK];
    /* ---- row 0 (simd_abpoa_cg_first_row / _first_dp, :617-688) ---- */
    {
        g->maxL[SRC] = g->maxR[SRC] = 0;
        for (int i = 0; i < g->node[SRC].out_n; ++i) { int o = g->node[SRC].out_id[i]; g->maxL[o] = g->maxR[o] = 1; }
        int dd = L - (g->remain[SRC] - rem_sink - 1);
        int end = MIN2(L, MAX2(g->maxR[SRC], dd) + w);        /* GET_AD_DP_END, abpoa_align.h:34-35 */
        d->beg[0] = 0; d->end[0] = end; d->off[0] = 0;
        ENSURE(end + 1);
        d->H[0] = 0; d->E1[0] = -oe1; d->E2[0] = -oe2; d->F1[0] = d->F2[0] = inf_min;
        for (int j = 1; j <= end; ++j) {
            d->F1[j] = -o1 - e1 * j; d->F2[j] = -o2 - e2 * j;
            d->H[j] = MAX2(d->F1[j], d->F2[j]);
            d->E1[j] = d->E2[j] = inf_min;
        }
        used = (size_t)end + 1; d->off[1] = (int64_t)used;
    }
    /* ---- rows 1..R-1 in topological index order (simd_abpoa_cg_dp, :935-1074) ---- */
    for (int r = 1; r < R; ++r) {
        int v = g->index_to_node[r]; node_t *nv = &g->node[v];
        int dd = L - (g->remain[v] - rem_sink - 1);
        int beg = MAX2(0, MIN2(g->maxL[v], dd) - w), end = MIN2(L, MAX2(g->maxR[v], dd) + w);
        int min_pre_beg = INT_MAX;
        for (int k = 0; k < nv->in_n; ++k) { int pr = g->node_to_index[nv->in_id[k]]; if (min_pre_beg > d->beg[pr]) min_pre_beg = d->beg[pr]; }
        if (beg / pn < min_pre_beg / pn) beg = min_pre_beg;   /* lane-group snap, :957-959 */
        d->beg[r] = beg; d->end[r] = end;
        int wr = end - beg + 1; if (wr < 0) die("negative band");
        ENSURE(wr);
        int32_t *H = d->H + used, *E1 = d->E1 + used, *E2 = d->E2 + used, *F1 = d->F1 + used, *F2 = d->F2 + used;
        const int *srow = mat + 5 * nv->base;
        /* M/E from every predecessor, then H' = max(M + s, E1, E2) (:967-1050) */
        for (int j = beg; j <= end; ++j) {
            int32_t m = inf_min, x1 = inf_min, x2 = inf_min;
            for (int k = 0; k < nv->in_n; ++k) {
                int pr = g->node_to_index[nv->in_id[k]];
                int32_t hp = cell(d, d->H, pr, j - 1); if (hp > m) m = hp;
                int32_t a = cell(d, d->E1, pr, j); if (a > x1) x1 = a;
                int32_t b = cell(d, d->E2, pr, j); if (b > x2) x2 = b;
            }
            int32_t s = j == 0 ? 0 : srow[q[j - 1]];           /* query profile column 0 is 0 (:536) */
            int32_t h = m + s;
            H[j - beg] = MAX3(h, x1, x2); E1[j - beg] = x1; E2[j - beg] = x2;
        }
        /* F along the row: F[j] = max(H'[j-1] - oe, F[j-1] - e) (:1038-1059); nothing to the left of beg */
        int32_t f1 = inf_min, f2 = inf_min;
        for (int j = beg; j <= end; ++j) {
            int32_t hme = H[j - beg];
            F1[j - beg] = f1; F2[j - beg] = f2;
            int32_t h = MAX3(hme, f1, f2);                     /* :1067 */
            H[j - beg] = h;
            E1[j - beg] = MAX2(E1[j - beg] - e1, h - oe1);     /* E for the next row, :1070-1071 */
            E2[j - beg] = MAX2(E2[j - beg] - e2, h - oe2);
            f1 = MAX2(hme - oe1, f1 - e1); f2 = MAX2(hme - oe2, f2 - e2);
        }
        d->off[r] = (int64_t)used; used += (size_t)wr; d->off[r + 1] = (int64_t)used;
        /* simd_abpoa_max_in_row + simd_abpoa_ada_max_i (:1107-1130) */
        int32_t mx = inf_min; int left = -1, right = -1;
        for (int j = beg; j <= end; ++j) {
            if (H[j - beg] > mx) { mx = H[j - beg]; left = right = j; } else if (H[j - beg] == mx) right = j;
        }
        for (int i = 0; i < nv->out_n; ++i) {
            int o = nv->out_id[i];
            if (right + 1 > g->maxR[o]) g->maxR[o] = right + 1;
            if (left + 1 < g->maxL[o]) g->maxL[o] = left + 1;
        }
    }
    /* ---- best cell over SINK's predecessors (simd_abpoa_global_get_max, :1092-1105) ---- */
    int32_t best = inf_min; int best_i = 0, best_j = 0;
    for (int k = 0; k < g->node[SINK].in_n; ++k) {
        int row = g->node_to_index[g->node[SINK].in_id[k]];
        int col = L > d->end[row] ? d->end[row] : L;
        int32_t sc = d->H[d->off[row] + (col - d->beg[row])];
        if (sc > best) { best = sc; best_i = row; best_j = col; }
    }
    *best_score_out = best;
    /* ---- traceback (simd_abpoa_cg_backtrack, :309-458) ---- */
    int i = best_i, j = best_j, cur_op = OP_ALL;
    if (best_j < L) push_cigar(cg, CINS, L - best_j, -1, L - 1);
    while (i > 0 && j > 0) {
        int id = g->index_to_node[i]; node_t *ni = &g->node[id];
        int32_t s = mat[5 * ni->base + q[j - 1]];
        int hit = 0;
        int32_t hij = cell(d, d->H, i, j);
        if (cur_op & OP_M) {
            for (int k = 0; k < ni->in_n; ++k) {
                int pi = g->node_to_index[ni->in_id[k]];
                if (j - 1 < d->beg[pi] || j - 1 > d->end[pi]) continue;
                if (cell(d, d->H, pi, j - 1) + s == hij) {
                    push_cigar(cg, CMATCH, 1, id, j - 1);
                    i = pi; --j; hit = 1; cur_op = OP_ALL; break;
                }
            }
        }
        if (!hit && (cur_op & OP_E)) {
            for (int k = 0; k < ni->in_n; ++k) {
                int pi = g->node_to_index[ni->in_id[k]];
                if (j < d->beg[pi] || j > d->end[pi]) continue;
                int32_t ph = cell(d, d->H, pi, j);
                if (cur_op & OP_E1) {
                    int32_t pe1 = cell(d, d->E1, pi, j);
                    int ok = (cur_op & OP_M) ? (hij == pe1) : (cell(d, d->E1, i, j) == pe1 - e1);
                    if (ok) {
                        cur_op = (ph - oe1 == pe1) ? (OP_M | OP_F) : OP_E1;
                        push_cigar(cg, CDEL, 1, id, j - 1); i = pi; hit = 1; break;
                    }
                }
                if (cur_op & OP_E2) {
                    int32_t pe2 = cell(d, d->E2, pi, j);
                    int ok = (cur_op & OP_M) ? (hij == pe2) : (cell(d, d->E2, i, j) == pe2 - e2);
                    if (ok) {
                        cur_op = (ph - oe2 == pe2) ? (OP_M | OP_F) : OP_E2;
                        push_cigar(cg, CDEL, 1, id, j - 1); i = pi; hit = 1; break;
                    }
                }
            }
        }
        if (!hit && (cur_op & OP_F)) {
            if (cur_op & OP_F1) {
                int32_t f = cell(d, d->F1, i, j);
                if (!(cur_op & OP_M) || hij == f) {
                    if (cell(d, d->H, i, j - 1) - oe1 == f) { cur_op = OP_M | OP_E; hit = 1; }
                    else if (cell(d, d->F1, i, j - 1) - e1 == f) { cur_op = OP_F1; hit = 1; }
                }
            }
            if (!hit && (cur_op & OP_F2)) {
                int32_t f = cell(d, d->F2, i, j);
                if (!(cur_op & OP_M) || hij == f) {
                    if (cell(d, d->H, i, j - 1) - oe2 == f) { cur_op = OP_M | OP_E; hit = 1; }
                    else if (cell(d, d->F2, i, j - 1) - e2 == f) { cur_op = OP_F2; hit = 1; }
                }
            }
            if (hit) { push_cigar(cg, CINS, 1, id, j - 1); --j; }
        }
        if (!hit && (cur_op & OP_M)) {
            for (int k = 0; k < ni->in_n; ++k) {
                int pi = g->node_to_index[ni->in_id[k]];
                if (j - 1 < d->beg[pi] || j - 1 > d->end[pi]) continue;
                if (cell(d, d->H, pi, j - 1) + s == hij) {
                    push_cigar(cg, CMATCH, 1, id, j - 1);
                    i = pi; --j; hit = 1; cur_op = OP_ALL; break;
                }
            }
        }
        if (!hit) die("Error in cg_backtrack");
    }
    if (j > 0) push_cigar(cg, CINS, j, -1, j - 1);
    for (int a = 0; a < cg->n >> 1; ++a) { uint64_t t = cg->c[a]; cg->c[a] = cg->c[cg->n - 1 - a]; cg->c[cg->n - 1 - a] = t; }
    return 0;
}

static void dp_free(dp_t *d) {
    free(d->beg); free(d->end); free(d->off); free(d->H); free(d->E1); free(d->E2); free(d->F1); free(d->F2);
    memset(d, 0, sizeof(*d));
}

/* abpoa_add_subgraph_alignment(SRC, SINK, inc_both_ends=1), abpoa_graph.c:689-774 */
static void fuse_alignment(graph_t *g, const uint8_t *seq, int seq_l, const cigar_t *cg, int read_id) {
    if (g->node_n == 2) {                                     /* abpoa_add_graph_sequence, :573-593 */
        int last = SRC;
        for (int i = 0; i < seq_l; ++i) { int cur = add_node(g, seq[i]); add_edge(g, last, cur, 0, read_id); last = cur; }
        add_edge(g, last, SINK, 0, read_id);
        topo_sort(g);
        return;
    }
    if (cg->n == 0) return;
    int query_id = -1, last_new = 0, last_id = SRC;
    for (int i = 0; i < cg->n; ++i) {
        int op = (int)(cg->c[i] & 0xf);
        if (op == CMATCH) {
            int node_id = (int)((cg->c[i] >> 34) & 0x3fffffff);
            query_id++;
            if (g->node[node_id].base != seq[query_id]) {
                int a = aligned_with_base(g, node_id, seq[query_id]);
                if (a != -1) { add_edge(g, last_id, a, 1 - last_new, read_id); last_id = a; last_new = 0; }
                else {
                    int nid = add_node(g, seq[query_id]);
                    add_edge(g, last_id, nid, 0, read_id);
                    last_id = nid; last_new = 1;
                    add_aligned(g, node_id, nid);
                }
            } else { add_edge(g, last_id, node_id, 1 - last_new, read_id); last_id = node_id; last_new = 0; }
        } else if (op == CINS) {
            int len = (int)((cg->c[i] >> 4) & 0x3fffffff);
            query_id += len;
            for (int j = len - 1; j >= 0; --j) {
                int nid = add_node(g, seq[query_id - j]);
                add_edge(g, last_id, nid, 0, read_id);
                last_id = nid; last_new = 1;
            }
        }
    }
    add_edge(g, last_id, SINK, 1 - last_new, read_id);
    topo_sort(g);
}

/* abpoa_DFS_set_msa_rank (abpoa_graph.c:359-410) + abpoa_generate_rc_msa (abpoa_output.c:149-176) */
static int build_msa(graph_t *g, int n_seq, uint8_t **msa_out) {
    int n = g->node_n, sp = 0, rank = 0;
    int *indeg = (int *)xcalloc(n, sizeof(int)), *st = (int *)xcalloc(n + 1, sizeof(int));
    for (int i = 0; i < n; ++i) indeg[i] = g->node[i].in_n;
    st[sp++] = SRC; g->msa_rank[SRC] = -1;
    int done = 0;
    while (sp > 0) {
        int cur = st[--sp]; node_t *c = &g->node[cur];
        if (g->msa_rank[cur] < 0) {
            g->msa_rank[cur] = rank;
            for (int i = 0; i < c->aln_n; ++i) g->msa_rank[c->aln_id[i]] = rank;
            rank++;
        }
        if (cur == SINK) { done = 1; break; }
        for (int i = 0; i < c->out_n; ++i) {
            int o = c->out_id[i];
            if (--indeg[o] == 0) {
                node_t *on = &g->node[o]; int ok = 1;
                for (int j = 0; j < on->aln_n; ++j) if (indeg[on->aln_id[j]] != 0) { ok = 0; break; }
                if (!ok) continue;
                st[sp++] = o; g->msa_rank[o] = -1;
                for (int j = 0; j < on->aln_n; ++j) { st[sp++] = on->aln_id[j]; g->msa_rank[on->aln_id[j]] = -1; }
            }
        }
    }
    if (!done) die("Error in set_msa_rank");
    free(indeg); free(st);
    int msa_len = g->msa_rank[SINK] - 1;
    uint8_t *msa = (uint8_t *)xcalloc((size_t)n_seq * (msa_len > 0 ? msa_len : 1), 1);
    memset(msa, 5, (size_t)n_seq * (msa_len > 0 ? msa_len : 1));
    for (int i = 2; i < n; ++i) {
        node_t *nd = &g->node[i];
        int r = g->msa_rank[i];
        for (int j = 0; j < nd->aln_n; ++j) r = MAX2(r, g->msa_rank[nd->aln_id[j]]);
        for (int w = 0; w < g->W; ++w) for (int e = 0; e < nd->out_n; ++e) {      /* abpoa_set_msa_seq, :105-122 */
            uint64_t bits = nd->rid[(size_t)e * g->W + w];
            while (bits) { int b = __builtin_ctzll(bits); msa[(size_t)(w * 64 + b) * msa_len + (r - 1)] = nd->base; bits &= bits - 1; }
        }
    }
    *msa_out = msa;
    return msa_len;
}

typedef struct { int64_t *w; size_t n, m; } wbuf_t;
static void wpush(wbuf_t *b, int64_t v) {
    if (!b) return;
    if (b->n == b->m) { b->m = b->m ? b->m * 2 : 1024; b->w = (int64_t *)xrealloc(b->w, b->m * sizeof(int64_t)); }
    b->w[b->n++] = v;
}

/* abpoa_msa -> abpoa_anchor_poa without anchors (progressive) or abpoa_poa (abpoa_align.c:208-352, 401-471) */
static int run_msa(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat, uint8_t **msa_out,
                   wbuf_t *tr, int64_t *cells_out) {
    if (p->gap_open1 <= 0 || p->gap_open2 <= 0 || p->wb < 0) die("only convex gap + adaptive band are restated");
    uint8_t **seqs = (uint8_t **)xcalloc(n_seq, sizeof(uint8_t *));
    size_t off = 0;
    for (int i = 0; i < n_seq; ++i) { seqs[i] = (uint8_t *)flat + off; off += lens[i]; if (lens[i] <= 0) die("empty sequence"); }
    int *order = (int *)xcalloc(n_seq, sizeof(int));
    guide_tree(p, seqs, lens, n_seq, order);
    graph_t g; graph_init(&g, n_seq);
    int64_t cells = 0;
    wpush(tr, n_seq); wpush(tr, 0); wpush(tr, 0);
    for (int i = 0; i < n_seq; ++i) wpush(tr, order[i]);
    for (int _i = 0; _i < n_seq; ++_i) {
        int i = order[_i], L = lens[i];
        cigar_t cg = {0, 0, 0}; dp_t d; memset(&d, 0, sizeof(d));
        int32_t best = 0; int node_n = g.node_n, n_rows = 0;
        if (g.node_n > 2) { align_to_graph(p, &g, seqs[i], L, &cg, &best, &d); n_rows = d.n_rows; }
        wpush(tr, i); wpush(tr, L); wpush(tr, node_n); wpush(tr, cg.n); wpush(tr, best); wpush(tr, n_rows);
        for (int c = 0; c < cg.n; ++c) wpush(tr, (int64_t)cg.c[c]);
        for (int r = 0; r < n_rows; ++r) wpush(tr, d.beg[r]);
        for (int r = 0; r < n_rows; ++r) { wpush(tr, d.end[r]); cells += d.end[r] - d.beg[r] + 1; }
        fuse_alignment(&g, seqs[i], L, &cg, i);
        free(cg.c); dp_free(&d);
    }
    uint8_t *msa = NULL; int msa_len = 0;
    if (g.node_n > 2) msa_len = build_msa(&g, n_seq, &msa); else msa = (uint8_t *)xcalloc(1, 1);
    if (tr) {
        tr->w[1] = msa_len; tr->w[2] = cells;
        size_t nbytes = (size_t)n_seq * msa_len, nw = (nbytes + 7) / 8, base = tr->n;
        for (size_t k = 0; k < nw; ++k) wpush(tr, 0);
        memcpy(tr->w + base, msa, nbytes);
    }
    if (cells_out) *cells_out = cells;
    if (msa_out) *msa_out = msa; else free(msa);
    graph_free(&g); free(seqs); free(order);
    return msa_len;
}

int oracle_poa_msa(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat, uint8_t **msa_out) {
    return run_msa(p, n_seq, lens, flat, msa_out, NULL, NULL);
}

int64_t *oracle_poa_msa_trace(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat, int64_t *n_words) {
    wbuf_t b = {0, 0, 0};
    run_msa(p, n_seq, lens, flat, NULL, &b, NULL);
    *n_words = (int64_t)b.n;
    return b.w;
}

int64_t oracle_poa_cells(const oracle_params_t *p, int n_seq, const int *lens, const uint8_t *flat) {
    int64_t cells = 0;
    run_msa(p, n_seq, lens, flat, NULL, NULL, &cells);
    return cells;
}

void oracle_free(void *p) { free(p); }

/* CPU baseline driver for the port (used only when oracle/_ref is absent): OpenMP over jobs, schedule(dynamic,1) */
#include <omp.h>
double oracle_poa_msa_many(const oracle_params_t *p, int64_t n_jobs, const int *n_seq, const int *lens, const uint8_t *flat,
                           int threads, int *msa_lens, uint64_t *checksum, uint64_t *hashes) {
    int64_t *len_off = (int64_t *)xcalloc(n_jobs + 1, sizeof(int64_t)), *seq_off = (int64_t *)xcalloc(n_jobs + 1, sizeof(int64_t));
    int64_t lo = 0, so = 0;
    for (int64_t j = 0; j < n_jobs; ++j) { len_off[j] = lo; seq_off[j] = so; for (int i = 0; i < n_seq[j]; ++i) so += lens[lo + i]; lo += n_seq[j]; }
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t sum = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : sum)
    for (int64_t j = 0; j < n_jobs; ++j) {
        uint8_t *msa = NULL;
        int ml = oracle_poa_msa(p, n_seq[j], lens + len_off[j], flat + seq_off[j], &msa);
        if (msa_lens) msa_lens[j] = ml;
        for (int64_t k = 0; k < (int64_t)n_seq[j] * ml; ++k) sum += msa[k];
        if (hashes) {          /* FNV-1a over (msa_len, bytes), as oracle/ref_harness.c */
            uint64_t h = (1469598103934665603ULL ^ (uint64_t)(uint32_t)ml) * 1099511628211ULL;
            for (int64_t k = 0; k < (int64_t)n_seq[j] * ml; ++k) { h ^= msa[k]; h *= 1099511628211ULL; }
            hashes[j] = h;
        }
        free(msa);
    }
    double t1 = omp_get_wtime();
    if (checksum) *checksum = sum;
    free(len_off); free(seq_off);
    return t1 - t0;
}
