// poa_graph.cuh -- the partial-order graph: fuse an alignment, topological (BFS) order, edge sort, max_remain,
// row tables for the DP, MSA rank and row-column fill, and the DP traceback.
//
// These are the strictly sequential parts of abPOA's per-sequence loop; on the device they run on one thread
// of the CTA that owns the job (the column-parallel DP sweep lives in poa_kernel.cu). Written as
// __host__ __device__ so tests/hosttest can exercise the very same code on the CPU of a GPU-less builder;
// the product only ever calls them from the kernel.
//
// Behaviour (not code) follows the reference abPOA v1.5.6 as Cactus drives it; file:line citations are
// relative to /root/reference/submodules/abPOA/src/.
#pragma once
#include <stdio.h>
#include "poa_types.h"

namespace barb200 {

// ---- edge lists ---------------------------------------------------------------------------------------
HD void graph_reset(Graph &g, int n_seq) {
    // abpoa_reset + abpoa_init_graph: just SRC and SINK (abpoa_graph.c:103-113, 783-795)
    g.node_n = 2; g.W = 1 + ((n_seq - 1) >> 6);
    g.in_used = 0; g.out_used = 0; g.err = JOB_OK;
    for (int i = 0; i < 2; ++i) {
        g.base[i] = 0; g.aln_n[i] = 0;
        g.in_off[i] = 0; g.in_n[i] = 0; g.in_cap[i] = 0;
        g.out_off[i] = 0; g.out_n[i] = 0; g.out_cap[i] = 0;
    }
}

HD int graph_add_node(Graph &g, uint8_t b) {          // abpoa_add_graph_node, abpoa_graph.c:471-478
    if (g.node_n >= g.node_cap) { g.err = JOB_ERR_NODE_CAP; return g.node_cap - 1; }
    int id = g.node_n++;
    g.base[id] = b; g.aln_n[id] = 0;
    g.in_off[id] = 0; g.in_n[id] = 0; g.in_cap[id] = 0;
    g.out_off[id] = 0; g.out_n[id] = 0; g.out_cap[id] = 0;
    return id;
}

// make room for one more in edge of node t; returns false on pool exhaustion
HD bool grow_in(Graph &g, int t) {
    if (g.in_n[t] < g.in_cap[t]) return true;
    int ncap = g.in_cap[t] ? g.in_cap[t] * 2 : 2;
    if (g.in_used + ncap > g.in_pool) { g.err = JOB_ERR_EDGE_CAP; return false; }
    int noff = g.in_used; g.in_used += ncap;
    for (int i = 0; i < g.in_n[t]; ++i) { g.in_id[noff + i] = g.in_id[g.in_off[t] + i]; g.in_w[noff + i] = g.in_w[g.in_off[t] + i]; }
    g.in_off[t] = noff; g.in_cap[t] = ncap;
    return true;
}

HD bool grow_out(Graph &g, int f) {
    if (g.out_n[f] < g.out_cap[f]) return true;
    int ncap = g.out_cap[f] ? g.out_cap[f] * 2 : 2;
    if (g.out_used + ncap > g.out_pool) { g.err = JOB_ERR_EDGE_CAP; return false; }
    int noff = g.out_used; g.out_used += ncap;
    const int W = g.W;
    for (int i = 0; i < g.out_n[f]; ++i) {
        g.out_id[noff + i] = g.out_id[g.out_off[f] + i]; g.out_w[noff + i] = g.out_w[g.out_off[f] + i];
        for (int w = 0; w < W; ++w) g.out_rid[(int64_t)(noff + i) * W + w] = g.out_rid[(int64_t)(g.out_off[f] + i) * W + w];
    }
    for (int i = g.out_n[f]; i < ncap; ++i) for (int w = 0; w < W; ++w) g.out_rid[(int64_t)(noff + i) * W + w] = 0;
    g.out_off[f] = noff; g.out_cap[f] = ncap;
    return true;
}

// abpoa_add_graph_edge with w = 1, add_read_id = 1 (abpoa_graph.c:480-556)
HD void graph_add_edge(Graph &g, int from, int to, int check_edge, int read_id) {
    int out_i = -1;
    if (check_edge) {
        const int io = g.in_off[to], in = g.in_n[to];
        for (int i = 0; i < in; ++i) if (g.in_id[io + i] == from) { g.in_w[io + i] += 1; break; }
        const int oo = g.out_off[from], on = g.out_n[from];
        for (int i = 0; i < on; ++i) if (g.out_id[oo + i] == to) { g.out_w[oo + i] += 1; out_i = i; break; }
    }
    if (out_i < 0) {
        if (!grow_in(g, to) || !grow_out(g, from)) return;
        int ip = g.in_off[to] + g.in_n[to]; g.in_id[ip] = from; g.in_w[ip] = 1; g.in_n[to]++;
        out_i = g.out_n[from];
        int op = g.out_off[from] + out_i; g.out_id[op] = to; g.out_w[op] = 1; g.out_n[from]++;
    }
    g.out_rid[(int64_t)(g.out_off[from] + out_i) * g.W + (read_id >> 6)] |= 1ULL << (read_id & 63);
}

HD void aln_push(Graph &g, int node, int id) {
    if (g.aln_n[node] >= 4) { g.err = JOB_ERR_ALIGNED_CAP; return; }
    g.aln_id[node * 4 + g.aln_n[node]] = id; g.aln_n[node]++;
}

// abpoa_add_graph_aligned_node, abpoa_graph.c:455-463
HD void graph_add_aligned(Graph &g, int node_id, int new_id) {
    const int n = g.aln_n[node_id];
    for (int i = 0; i < n; ++i) { int a = g.aln_id[node_id * 4 + i]; aln_push(g, a, new_id); aln_push(g, new_id, a); }
    aln_push(g, node_id, new_id); aln_push(g, new_id, node_id);
}

// abpoa_get_aligned_id, abpoa_graph.c:439-448
HD int graph_aligned_with_base(const Graph &g, int node_id, uint8_t b) {
    for (int i = 0; i < g.aln_n[node_id]; ++i) { int a = g.aln_id[node_id * 4 + i]; if (g.base[a] == b) return a; }
    return -1;
}

// ---- topological sort ---------------------------------------------------------------------------------
// abpoa_BFS_set_node_index, abpoa_graph.c:221-266. FIFO over nodes whose in-degree reached zero; a node is
// only released together with all nodes aligned to it (same MSA column), which follow it in the queue.
HD void graph_bfs_index(Graph &g) {
    const int n = g.node_n;
    int *indeg = g.tmp0, *q = g.tmp1;
    for (int i = 0; i < n; ++i) indeg[i] = g.in_n[i];
    int head = 0, tail = 0, index = 0;
    q[tail++] = SRC_ID;
    while (head < tail) {
        const int cur = q[head++];
        g.index_to_node[index] = cur; g.node_to_index[cur] = index++;
        if (cur == SINK_ID) return;
        const int oo = g.out_off[cur], on = g.out_n[cur];
        for (int i = 0; i < on; ++i) {
            const int o = g.out_id[oo + i];
            if (--indeg[o] == 0) {
                bool ok = true;
                const int an = g.aln_n[o];
                for (int j = 0; j < an; ++j) if (indeg[g.aln_id[o * 4 + j]] != 0) { ok = false; break; }
                if (!ok) continue;
                q[tail++] = o;
                for (int j = 0; j < an; ++j) q[tail++] = g.aln_id[o * 4 + j];
            }
        }
    }
    g.err = JOB_ERR_TOPO;
}

// abpoa_sort_in_out_ids for ONE node (abpoa_graph.c:192-219): the reference's exchange sort, weight
// descending, not stable -- the resulting order breaks ties in the DP traceback, so it is reproduced verbatim.
HD void graph_sort_node_edges(Graph &g, int v) {
    {
        const int o = g.in_off[v], n = g.in_n[v];
        for (int j = 0; j < n - 1; ++j) for (int k = j + 1; k < n; ++k)
            if (g.in_w[o + j] < g.in_w[o + k]) {
                int t = g.in_id[o + j]; g.in_id[o + j] = g.in_id[o + k]; g.in_id[o + k] = t;
                t = g.in_w[o + j]; g.in_w[o + j] = g.in_w[o + k]; g.in_w[o + k] = t;
            }
    }
    {
        const int o = g.out_off[v], n = g.out_n[v], W = g.W;
        for (int j = 0; j < n - 1; ++j) for (int k = j + 1; k < n; ++k)
            if (g.out_w[o + j] < g.out_w[o + k]) {
                int t = g.out_id[o + j]; g.out_id[o + j] = g.out_id[o + k]; g.out_id[o + k] = t;
                t = g.out_w[o + j]; g.out_w[o + j] = g.out_w[o + k]; g.out_w[o + k] = t;
                for (int w = 0; w < W; ++w) {
                    uint64_t r = g.out_rid[(int64_t)(o + j) * W + w];
                    g.out_rid[(int64_t)(o + j) * W + w] = g.out_rid[(int64_t)(o + k) * W + w];
                    g.out_rid[(int64_t)(o + k) * W + w] = r;
                }
            }
    }
}

// abpoa_BFS_set_node_remain, abpoa_graph.c:268-309: reverse BFS from SINK (-1);
// remain[v] = remain[first heaviest out neighbour] + 1
HD void graph_bfs_remain(Graph &g) {
    const int n = g.node_n;
    int *outdeg = g.tmp0, *q = g.tmp1;
    for (int i = 0; i < n; ++i) { outdeg[i] = g.out_n[i]; g.remain[i] = 0; }
    int head = 0, tail = 0;
    q[tail++] = SINK_ID; g.remain[SINK_ID] = -1;
    while (head < tail) {
        const int cur = q[head++];
        if (cur != SINK_ID) {
            int max_w = -1, max_id = -1;
            const int oo = g.out_off[cur], on = g.out_n[cur];
            for (int i = 0; i < on; ++i) if (g.out_w[oo + i] > max_w) { max_w = g.out_w[oo + i]; max_id = g.out_id[oo + i]; }
            g.remain[cur] = g.remain[max_id] + 1;
        }
        if (cur == SRC_ID) return;
        const int io = g.in_off[cur], in = g.in_n[cur];
        for (int i = 0; i < in; ++i) if (--outdeg[g.in_id[io + i]] == 0) q[tail++] = g.in_id[io + i];
    }
    g.err = JOB_ERR_TOPO;
}

// Row-major tables the DP sweeps: for topological index r the node's base, its remain term and its
// predecessor ROWS in stored in_id order (what simd_abpoa_init_var collects, abpoa_align_simd.c:550-558).
HD void graph_build_row(const Graph &g, RowTables &rt, int r, int off) {
    const int v = g.index_to_node[r];
    const int io = g.in_off[v], in = g.in_n[v];
    for (int k = 0; k < in; ++k) rt.pre_row[off + k] = g.node_to_index[g.in_id[io + k]];
    RowRec rec;
    rec.base_npre = g.base[v] | (in << 8);
    rec.rd = g.remain[v] - g.remain[SINK_ID] - 1;
    rec.pre_off = off;
    rec.pre0 = in ? g.node_to_index[g.in_id[io]] : -1;
    rt.rec[r] = rec;
}
HD void graph_build_rows(const Graph &g, RowTables &rt) {
    int off = 0;
    for (int r = 0; r < g.node_n; ++r) { graph_build_row(g, rt, r, off); off += g.in_n[g.index_to_node[r]]; }
}
HD int row_npre(const RowTables &rt, int r) { return rt.rec[r].base_npre >> 8; }
HD int row_base(const RowTables &rt, int r) { return rt.rec[r].base_npre & 0xff; }

// abpoa_topological_sort (abpoa_graph.c:322-357), serial form
HD void graph_topo_sort_serial(Graph &g, RowTables &rt) {
    graph_bfs_index(g);
    if (g.err) return;
    for (int v = 0; v < g.node_n; ++v) graph_sort_node_edges(g, v);
    graph_bfs_remain(g);
    if (g.err) return;
    graph_build_rows(g, rt);
}

// ---- fusing an alignment into the graph ------------------------------------------------------------------
// abpoa_add_graph_sequence (abpoa_graph.c:573-593): first sequence becomes a chain SRC -> b0 -> ... -> SINK
HD void graph_add_first_sequence(Graph &g, const uint8_t *seq, int len, int read_id) {
    int last = SRC_ID;
    for (int i = 0; i < len; ++i) { int cur = graph_add_node(g, seq[i]); graph_add_edge(g, last, cur, 0, read_id); last = cur; if (g.err) return; }
    graph_add_edge(g, last, SINK_ID, 0, read_id);
}

// abpoa_add_subgraph_alignment(SRC, SINK, inc_both_ends = 1), abpoa_graph.c:689-774
HD void graph_fuse_alignment(Graph &g, const uint8_t *seq, const uint64_t *cigar, int n_cigar, int read_id) {
    if (n_cigar == 0) return;
    int query_id = -1, last_new = 0, last_id = SRC_ID;
    for (int i = 0; i < n_cigar && !g.err; ++i) {
        const int op = (int)(cigar[i] & 0xf);
        if (op == CMATCH) {
            const int node_id = (int)((cigar[i] >> 34) & 0x3fffffff);
            ++query_id;
            const uint8_t b = seq[query_id];
            if (g.base[node_id] != b) {
                int a = graph_aligned_with_base(g, node_id, b);
                if (a != -1) { graph_add_edge(g, last_id, a, 1 - last_new, read_id); last_id = a; last_new = 0; }
                else {
                    int nid = graph_add_node(g, b);
                    graph_add_edge(g, last_id, nid, 0, read_id);
                    last_id = nid; last_new = 1;
                    graph_add_aligned(g, node_id, nid);
                }
            } else { graph_add_edge(g, last_id, node_id, 1 - last_new, read_id); last_id = node_id; last_new = 0; }
        } else if (op == CINS) {
            const int len = (int)((cigar[i] >> 4) & 0x3fffffff);
            query_id += len;
            for (int j = len - 1; j >= 0 && !g.err; --j) {
                int nid = graph_add_node(g, seq[query_id - j]);
                graph_add_edge(g, last_id, nid, 0, read_id);
                last_id = nid; last_new = 1;
            }
        }   // CDEL: nothing (abpoa_graph.c:759-762)
    }
    if (!g.err) graph_add_edge(g, last_id, SINK_ID, 1 - last_new, read_id);
}

// ---- MSA ----------------------------------------------------------------------------------------------------
// abpoa_DFS_set_msa_rank (abpoa_graph.c:359-410): LIFO version of the walk above; aligned nodes share a rank.
// Returns msa_len = rank(SINK) - 1 (abpoa_output.c:157).
HD int graph_msa_rank(Graph &g) {
    const int n = g.node_n;
    int *indeg = g.tmp0, *st = g.tmp1;
    for (int i = 0; i < n; ++i) indeg[i] = g.in_n[i];
    int sp = 0, rank = 0;
    st[sp++] = SRC_ID; g.msa_rank[SRC_ID] = -1;
    while (sp > 0) {
        const int cur = st[--sp];
        if (g.msa_rank[cur] < 0) {
            g.msa_rank[cur] = rank;
            for (int i = 0; i < g.aln_n[cur]; ++i) g.msa_rank[g.aln_id[cur * 4 + i]] = rank;
            ++rank;
        }
        if (cur == SINK_ID) return g.msa_rank[SINK_ID] - 1;
        const int oo = g.out_off[cur], on = g.out_n[cur];
        for (int i = 0; i < on; ++i) {
            const int o = g.out_id[oo + i];
            if (--indeg[o] == 0) {
                bool ok = true;
                const int an = g.aln_n[o];
                for (int j = 0; j < an; ++j) if (indeg[g.aln_id[o * 4 + j]] != 0) { ok = false; break; }
                if (!ok) continue;
                st[sp++] = o; g.msa_rank[o] = -1;
                for (int j = 0; j < an; ++j) { int a = g.aln_id[o * 4 + j]; st[sp++] = a; g.msa_rank[a] = -1; }
            }
        }
    }
    g.err = JOB_ERR_TOPO;
    return 0;
}

// abpoa_generate_rc_msa / abpoa_set_msa_seq for ONE node (abpoa_output.c:105-122, 167-176):
// every read whose bit is set on one of the node's out edges gets the node's base in column rank-1.
HD void graph_msa_fill_node(const Graph &g, int v, uint8_t *msa, int64_t stride) {
    int rank = g.msa_rank[v];
    for (int j = 0; j < g.aln_n[v]; ++j) rank = imax(rank, g.msa_rank[g.aln_id[v * 4 + j]]);
    const int oo = g.out_off[v], on = g.out_n[v], W = g.W;
    const uint8_t b = g.base[v];
    for (int e = 0; e < on; ++e) for (int w = 0; w < W; ++w) {
        uint64_t bits = g.out_rid[(int64_t)(oo + e) * W + w];
        while (bits) {
#if defined(__CUDA_ARCH__)
            const int bit = __ffsll((long long)bits) - 1;
#else
            const int bit = __builtin_ctzll(bits);
#endif
            msa[(int64_t)(w * 64 + bit) * stride + (rank - 1)] = b;
            bits &= bits - 1;
        }
    }
}

// ---- traceback ------------------------------------------------------------------------------------------------
// plane 0: H; 1: E1; 2: E2 (the two E values are stored as 16-bit distances below H, poa_types.h)
HD int plane_cell(const DpState &d, int inf_min, int row, int plane, int j) {
    const int beg = d.info[row].beg, end = d.info[row].end;
    if (j < beg || j > end) return inf_min;                 // out-of-band lanes hold inf_min (abpoa_align_simd.c:1035-1036)
    const int *rowp = d.planes + d.row_off[row];
    const int h = rowp[plane_index(beg, end, 0, j)];
    if (plane == 0) return h;
    const unsigned code = (unsigned)rowp[plane_index(beg, end, 1, j)];
    return e_decode(h, plane == 1 ? (int)(code & 0xffffu) : (int)(code >> 16), inf_min);
}

// H' = max(M + s, E1, E2) of cell (i, k), k inside row i's band: the value the F recurrences start from
// (abpoa_align_simd.c:1033-1050), recomputed from the predecessor rows' stored H / E.
HD int row_h_prime(const RowTables &rt, const DpState &d, const PoaParams &P, const uint8_t *q, int L, int i, int k) {
    const int inf = P.inf_min;
    const int p0 = rt.rec[i].pre_off, p1 = p0 + row_npre(rt, i);
    int m = inf, x1 = inf, x2 = inf;
    for (int t = p0; t < p1; ++t) {
        const int pi = rt.pre_row[t];
        m = imax(m, plane_cell(d, inf, pi, 0, k - 1)); x1 = imax(x1, plane_cell(d, inf, pi, 1, k)); x2 = imax(x2, plane_cell(d, inf, pi, 2, k));
    }
    const int s = (k >= 1 && k <= L) ? P.mat[5 * row_base(rt, i) + q[k - 1]] : 0;
    return imax(imax(m + s, x1), x2);
}

// F1 / F2 of row i for the columns [beg, j] into the traceback's scratch (d.fc): F[k] = max_{beg <= t < k} H'[t] - oe - (k-1-t) e,
// in the DP sweep's own "A space" form (poa_kernel.cu: row_pass1 / row_pass2) so that finite values are the ones the sweep
// would have stored. Row 0 has the closed form of simd_abpoa_cg_first_dp (abpoa_align_simd.c:669-688). Serial form (one thread): the host build and
// the serial_phases debug mode; the kernel's warp traceback (poa_cta.cuh: warp_backtrack_step) never builds F rows.
HD void row_f_cache(const RowTables &rt, DpState &d, const PoaParams &P, const uint8_t *q, int L, int i, int j) {
    if (d.fc_row == i && d.fc_hi >= j) return;
    const int inf = P.inf_min, e1 = P.e1, e2 = P.e2, o1 = P.o1, o2 = P.o2;
    const int beg = d.info[i].beg;
    int *f1 = d.fc, *f2 = d.fc + d.fc_cap;
    int c1 = inf + beg * e1 + o1, c2 = inf + beg * e2 + o2;
    for (int k = beg; k <= j; ++k) {
        if (i == 0) { f1[k] = k == 0 ? inf : -o1 - e1 * k; f2[k] = k == 0 ? inf : -o2 - e2 * k; continue; }
        f1[k] = c1 - o1 - k * e1; f2[k] = c2 - o2 - k * e2;
        if (k < j) { const int hp = row_h_prime(rt, d, P, q, L, i, k); c1 = imax(c1, hp + k * e1); c2 = imax(c2, hp + k * e2); }
    }
    d.fc_row = i; d.fc_hi = j;
}
// F1 (which = 0) / F2 (which = 1) of the cached row at column k (out-of-band -> inf_min, like every plane)
HD int row_f(const DpState &d, int inf_min, int which, int k) {
    const int beg = d.info[d.fc_row].beg;
    if (k < beg || k > d.fc_hi) return inf_min;
    return d.fc[which * d.fc_cap + k];
}

HD void push_cigar(DpState &d, Graph &g, int op, int len, int node_id, int query_id) {   // abpoa_push_cigar, abpoa_align.h:58-78
    const uint64_t l = (uint64_t)len;
    if (d.n_cigar == 0 || op != CINS || op != (int)(d.cigar[d.n_cigar - 1] & 0xf)) {
        if (d.n_cigar >= d.cigar_cap) { g.err = JOB_ERR_CIGAR_CAP; return; }
        const uint64_t n_id = (uint64_t)(int64_t)node_id, q_id = (uint64_t)(int64_t)query_id;
        if (op == CMATCH) d.cigar[d.n_cigar++] = n_id << 34 | q_id << 4 | (uint64_t)op;
        else if (op == CINS) d.cigar[d.n_cigar++] = q_id << 34 | l << 4 | (uint64_t)op;
        else d.cigar[d.n_cigar++] = n_id << 34 | l << 4 | (uint64_t)op;
    } else d.cigar[d.n_cigar - 1] += l << 4;
}

// simd_abpoa_global_get_max (abpoa_align_simd.c:1092-1105): best cell over SINK's predecessors in stored
// order, column min(L, dp_end), strictly greater wins.
HD void dp_best_cell(const Graph &g, const RowTables &rt, DpState &d, const PoaParams &P, int L) {
    const int sink_row = g.node_n - 1;
    int best = P.inf_min, bi = 0, bj = 0;
    const int sp0 = rt.rec[sink_row].pre_off, sp1 = sp0 + row_npre(rt, sink_row);
    for (int k = sp0; k < sp1; ++k) {
        const int row = rt.pre_row[k];
        const int col = L > d.info[row].end ? d.info[row].end : L;
        const int sc = plane_cell(d, P.inf_min, row, 0, col);
        if (sc > best) { best = sc; bi = row; bj = col; }
    }
    d.best_score = best; d.best_i = bi; d.best_j = bj;
}

// One iteration of simd_abpoa_cg_backtrack's loop (abpoa_align_simd.c:319-450) with put_gap_on_right =
// put_gap_at_end = 0: op priority M over predecessors in stored order, then E1/E2 per predecessor, then F1, F2,
// then M again; cur_op carries which gap state the walk is in. Moves (i, j), returns the cigar op (node `id`,
// query index j_before - 1) or -1 when no op explains the cell (the reference aborts, :448).
HD int backtrack_step(const Graph &g, const RowTables &rt, DpState &d, const PoaParams &P, const uint8_t *q, int L,
                      int &i, int &j, int &cur_op) {
    const int inf = P.inf_min, e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int s = P.mat[5 * row_base(rt, i) + q[j - 1]];
    const int p0 = rt.rec[i].pre_off, p1 = p0 + row_npre(rt, i);
    const int hij = plane_cell(d, inf, i, 0, j);
    if (cur_op & OP_M) {
        for (int k = p0; k < p1; ++k) {
            const int pi = rt.pre_row[k];
            if (j - 1 < d.info[pi].beg || j - 1 > d.info[pi].end) continue;
            if (plane_cell(d, inf, pi, 0, j - 1) + s == hij) { i = pi; --j; cur_op = OP_ALL; return CMATCH; }
        }
    }
    if (cur_op & OP_E) {
        for (int k = p0; k < p1; ++k) {
            const int pi = rt.pre_row[k];
            if (j < d.info[pi].beg || j > d.info[pi].end) continue;
            const int ph = plane_cell(d, inf, pi, 0, j);
            if (cur_op & OP_E1) {
                const int pe1 = plane_cell(d, inf, pi, 1, j);
                const bool ok = (cur_op & OP_M) ? (hij == pe1) : (plane_cell(d, inf, i, 1, j) == pe1 - e1);
                if (ok) { cur_op = (ph - oe1 == pe1) ? (OP_M | OP_F) : OP_E1; i = pi; return CDEL; }
            }
            if (cur_op & OP_E2) {
                const int pe2 = plane_cell(d, inf, pi, 2, j);
                const bool ok = (cur_op & OP_M) ? (hij == pe2) : (plane_cell(d, inf, i, 2, j) == pe2 - e2);
                if (ok) { cur_op = (ph - oe2 == pe2) ? (OP_M | OP_F) : OP_E2; i = pi; return CDEL; }
            }
        }
    }
    if (cur_op & OP_F) {
        bool hit = false;
        row_f_cache(rt, d, P, q, L, i, j);            // F1 / F2 of row i up to column j (not stored by the sweep)
        if (cur_op & OP_F1) {
            const int f = row_f(d, inf, 0, j);
            if (!(cur_op & OP_M) || hij == f) {
                if (plane_cell(d, inf, i, 0, j - 1) - oe1 == f) { cur_op = OP_M | OP_E; hit = true; }
                else if (row_f(d, inf, 0, j - 1) - e1 == f) { cur_op = OP_F1; hit = true; }
            }
        }
        if (!hit && (cur_op & OP_F2)) {
            const int f = row_f(d, inf, 1, j);
            if (!(cur_op & OP_M) || hij == f) {
                if (plane_cell(d, inf, i, 0, j - 1) - oe2 == f) { cur_op = OP_M | OP_E; hit = true; }
                else if (row_f(d, inf, 1, j - 1) - e2 == f) { cur_op = OP_F2; hit = true; }
            }
        }
        if (hit) { --j; return CINS; }
    }
    // (the reference retries M here, :429-446; with put_gap_on_right = 0 that is the test that already failed above)
    return -1;
}

// The rule the kernel's warp traceback applies (poa_cta.cuh: warp_backtrack_step), stated serially: M and E exactly as in
// backtrack_step; an insertion is taken WHOLE and without F planes. With F[j] = max_t H[i][j-t] - oe - (t-1) e (:1060-1075) and
// H[i][j] >= F[j]: H[i][j] == F1[j] iff some t >= 1 has H[i][j-t] == H[i][j] + oe1 + (t-1) e1, the smallest such t is the length of
// the insertion the reference walks ("open" is tested before "extend" at every column, :401-415), and since H[i][j] >= F2[j] bounds
// H[i][j-t] <= H[i][j] + oe2 + (t-1) e2 the F1 equation has no solution beyond (oe2 - oe1) / (e1 - e2) + 1 columns when e1 > e2.
// Returns the op, `len` = its length (1 unless CINS), moves (i, j, cur_op) past the whole op. The host build runs it in lock step
// with the reference's rule on every traceback (dp_backtrack below): the CPU suite checks the claim on all of its inputs.
HD int backtrack_step_without_f(const RowTables &rt, const DpState &d, const PoaParams &P, const uint8_t *q, int &i, int &j, int &cur_op, int &len) {
    const int inf = P.inf_min, e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    len = 1;
    const int s = P.mat[5 * row_base(rt, i) + q[j - 1]];
    const int p0 = rt.rec[i].pre_off, p1 = p0 + row_npre(rt, i);
    const int hij = plane_cell(d, inf, i, 0, j);
    if (cur_op & OP_M)
        for (int k = p0; k < p1; ++k) {
            const int pi = rt.pre_row[k];
            if (j - 1 < d.info[pi].beg || j - 1 > d.info[pi].end) continue;
            if (plane_cell(d, inf, pi, 0, j - 1) + s == hij) { i = pi; --j; cur_op = OP_ALL; return CMATCH; }
        }
    if (cur_op & OP_E)
        for (int k = p0; k < p1; ++k) {
            const int pi = rt.pre_row[k];
            if (j < d.info[pi].beg || j > d.info[pi].end) continue;
            const int ph = plane_cell(d, inf, pi, 0, j), pe1 = plane_cell(d, inf, pi, 1, j), pe2 = plane_cell(d, inf, pi, 2, j);
            const bool ok1 = (cur_op & OP_E1) && ((cur_op & OP_M) ? hij == pe1 : plane_cell(d, inf, i, 1, j) == pe1 - e1);
            const bool ok2 = (cur_op & OP_E2) && ((cur_op & OP_M) ? hij == pe2 : plane_cell(d, inf, i, 2, j) == pe2 - e2);
            if (ok1) { cur_op = (ph - oe1 == pe1) ? (OP_M | OP_F) : OP_E1; i = pi; return CDEL; }
            if (ok2) { cur_op = (ph - oe2 == pe2) ? (OP_M | OP_F) : OP_E2; i = pi; return CDEL; }
        }
    if (cur_op & OP_F) {
        if (!(cur_op & OP_M)) return -1;                     // F states always carry M: insertions are never left half walked
        const int avail = j - d.info[i].beg;
        int t1_max;
        if (e1 > e2) t1_max = oe2 >= oe1 ? (oe2 - oe1) / (e1 - e2) + 1 : 0;
        else t1_max = (e1 < e2 || oe1 <= oe2) ? avail : 0;
        if (t1_max > avail) t1_max = avail;
        for (int which = 0; which < 2; ++which) {
            if (!(cur_op & (which ? OP_F2 : OP_F1))) continue;
            const int oe = which ? oe2 : oe1, e = which ? e2 : e1, tmax = which ? avail : t1_max;
            for (int t = 1; t <= tmax; ++t)
                if (plane_cell(d, inf, i, 0, j - t) == hij + oe + (t - 1) * e) { len = t; j -= t; cur_op = OP_M | OP_E; return CINS; }
        }
    }
    return -1;
}

// simd_abpoa_cg_backtrack (abpoa_align_simd.c:309-458), serial form. Emits the graph cigar in forward order.
HD void dp_backtrack(Graph &g, const RowTables &rt, DpState &d, const PoaParams &P, const uint8_t *q, int L) {
    d.n_cigar = 0; d.fc_row = -1; d.fc_hi = -1;
    int i = d.best_i, j = d.best_j, cur_op = OP_ALL;
    if (j < L) push_cigar(d, g, CINS, L - j, -1, L - 1);
#if !defined(__CUDA_ARCH__)
    int chk_op = -1, chk_left = 0, chk_i = 0, chk_j = 0, chk_cur = 0;     // the F-free rule's prediction (below)
#endif
    while (i > 0 && j > 0 && !g.err) {
        const int id = g.index_to_node[i], jq = j - 1;
#if !defined(__CUDA_ARCH__)
        // host build: the F-free rule must walk the same ops (an insertion = its `len` single steps of the reference's rule)
        if (!(cur_op == OP_F1 || cur_op == OP_F2)) {
            int ci = i, cj = j, cop = cur_op, clen = 1;
            chk_op = backtrack_step_without_f(rt, d, P, q, ci, cj, cop, clen);
            chk_left = chk_op == CINS ? clen : 1; chk_i = ci; chk_j = cj; chk_cur = cop;
        }
#endif
        const int op = backtrack_step(g, rt, d, P, q, L, i, j, cur_op);
#if !defined(__CUDA_ARCH__)
        if (op >= 0) {
            bool same = op == chk_op && chk_left >= 1;
            if (same && --chk_left == 0) same = i == chk_i && j == chk_j && cur_op == chk_cur;
            if (!same) {
                fprintf(stderr, "barb200: the F-free traceback rule disagrees with the reference's at row %d col %d (op %d vs %d)\n", i, j, op, chk_op);
                g.err = JOB_ERR_BACKTRACK; return;
            }
        }
#endif
        if (op < 0) {
#if defined(__CUDA_ARCH__)
            printf("barb200: backtrack stuck at row %d col %d cur_op %d (band %d..%d, H %d) best %d,%d n_cigar %d\n", i, j, cur_op,
                   d.info[i].beg, d.info[i].end, plane_cell(d, P.inf_min, i, 0, j), d.best_i, d.best_j, d.n_cigar);
#endif
            g.err = JOB_ERR_BACKTRACK; return;
        }
        push_cigar(d, g, op, 1, id, jq);
    }
    if (g.err) return;
    if (j > 0) push_cigar(d, g, CINS, j, -1, j - 1);
    for (int a = 0; a < d.n_cigar >> 1; ++a) { uint64_t t = d.cigar[a]; d.cigar[a] = d.cigar[d.n_cigar - 1 - a]; d.cigar[d.n_cigar - 1 - a] = t; }
}

// lane count the reference's SIMD path would pick for this alignment (16 = int16 lanes at AVX2, 8 = int32);
// it leaks into dp_beg through the lane-group snap (abpoa_align_simd.c:957-959, 1293-1302)
HD int reference_lane_count(const PoaParams &P, int L, int node_n) {
    const int len = L > node_n ? L : node_n;
    const int max_score = imax(L * P.max_mat, len * P.e1 + P.o1);
    return (max_score <= 32767 - P.min_mis - (P.o1 + P.e1) - (P.o2 + P.e2)) ? 16 : 8;
}

}  // namespace barb200
