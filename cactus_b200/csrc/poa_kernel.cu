// poa_kernel.cu -- the fused, batched partial-order-alignment kernel for sm_100a.
//
// One CTA owns one job (= one abpoa_msa call of the reference: K sequences -> one MSA) from start to finish:
// it keeps the job's partial-order graph in its slot of device memory, aligns the sequences to it one after the
// other and never returns to the host in between. CTAs are persistent: each pulls the next job index from a
// global counter, so a launch of (148 SMs x resident CTAs) blocks streams through thousands of jobs while the
// strictly sequential phases of some jobs (traceback, graph fusion, topological sort) overlap the bandwidth-bound
// DP sweeps of the others on the same SM.
//
// DP sweep (the hot loop; replaces simd_abpoa_cg_dp + first row + row max + adaptive band,
// abPOA src/abpoa_align_simd.c:617-688, 935-1130):
//   * graph rows in topological order, strictly one after the other (the adaptive band of a row needs the argmax
//     columns of all predecessor rows), columns of a row in parallel: thread t owns 4 adjacent columns per pass;
//   * predecessor row values come from a double-buffered shared-memory copy of the previous row when the
//     predecessor is the row just computed (the common case in a near-linear graph), otherwise from the planes in
//     global memory (L2);
//   * the max-plus recurrence of the two insertion states F1/F2 along the row is turned into a plain prefix
//     maximum by the substitution A[k] = H'[k] - oe + (k+1)*e  =>  F[j] = max_{k<j} A[k] - j*e, done as
//     4 serial cells per thread, a warp shuffle scan over the 32 thread aggregates and a redux over the warp
//     aggregates staged in shared memory;
//   * all five int32 planes (H, E1, E2, F1, F2) of the row's band are written once, coalesced 16 B per thread,
//     to global memory for the traceback: 20 B/cell of HBM write traffic is what bounds the kernel.
// Integer DP: no tensor cores. int32 everywhere with the reference's own "minus infinity" so that finite cells
// are bit-identical to abPOA's AVX2 path.
#include <cuda_runtime.h>
#include <stdint.h>
#include "poa_graph.cuh"
#include "poa_kernel.cuh"

namespace barb200 {

#define FULL 0xffffffffu

struct KShared {
    Graph g; RowTables rt; DpState d;
    int job, msa_len_s, abort_s;
    int smat[25];
    int wF[2][2][32];      // [pass parity][plane F1/F2][warp] block scan staging
    int wM[3][32];         // row max / leftmost / rightmost per warp
};

__device__ __forceinline__ int4 ld4(const int *p) { return *reinterpret_cast<const int4 *>(p); }
__device__ __forceinline__ void st4(int *p, int4 v) { *reinterpret_cast<int4 *>(p) = v; }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }

// ---------------------------------------------------------------------------------------------------------
// banded convex-gap DP of query q[1..L] against the sorted graph. All threads of the CTA.
// Returns the number of banded cells (sum of dp_end-dp_beg+1), or -1 if the planes outgrew the slot.
// ---------------------------------------------------------------------------------------------------------
__device__ long long dp_sweep(KShared &S, const PoaParams &P, const uint8_t *__restrict__ qg, int L,
                              uint8_t *sq, int *rowbuf, int rb_stride, bool use_smem) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const Graph &g = S.g; const RowTables &rt = S.rt; DpState &d = S.d;
    const int node_n = g.node_n, R = node_n - 1;
    const int NEG = P.inf_min, e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int w = P.wb + (int)(P.wf * L);                                    // abpoa_align_simd.c:474
    const int pn = reference_lane_count(P, L, node_n);

    // query bytes to shared memory; sq[j] = q_j for j = 1..L, sq[0] unused (column 0 scores 0, :536)
    for (int j = tid; j <= L + 4; j += T) sq[j] = (j >= 1 && j <= L) ? qg[j - 1] : 4;
    int *bufH[2], *bufE1[2], *bufE2[2];
    for (int b = 0; b < 2; ++b) {
        bufH[b] = rowbuf + (b * 3 + 0) * rb_stride + 4;     // +4: index -1 is addressable, 16 B alignment kept
        bufE1[b] = rowbuf + (b * 3 + 1) * rb_stride + 4;
        bufE2[b] = rowbuf + (b * 3 + 2) * rb_stride + 4;
    }
    __syncthreads();

    // ---- row 0 (simd_abpoa_cg_first_dp, :617-688) ----
    int prev_beg = 0, prev_end, prev_left = 0, prev_right = 0;
    long long cur_off = 0, prev_off = 0, cells = 0;
    {
        const int dd = L - rt.row_rd[0];
        prev_end = min(L, max(0, dd) + w);
        const int wr4 = (prev_end | 3) + 1;
        if (5LL * wr4 > d.plane_cap) return -1;
        int *row = d.planes;
        for (int c0 = tid * 4; c0 <= prev_end; c0 += T * 4) {
            int h[4], x1[4], x2[4], f1[4], f2[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = c0 + t;
                if (j == 0) { h[t] = 0; x1[t] = -oe1; x2[t] = -oe2; f1[t] = NEG; f2[t] = NEG; }
                else if (j <= prev_end) { f1[t] = -P.o1 - e1 * j; f2[t] = -P.o2 - e2 * j; h[t] = max(f1[t], f2[t]); x1[t] = NEG; x2[t] = NEG; }
                else { h[t] = x1[t] = x2[t] = f1[t] = f2[t] = NEG; }
            }
            st4(row + c0, make_int4(h[0], h[1], h[2], h[3]));
            st4(row + wr4 + c0, make_int4(x1[0], x1[1], x1[2], x1[3]));
            st4(row + 2 * wr4 + c0, make_int4(x2[0], x2[1], x2[2], x2[3]));
            st4(row + 3 * wr4 + c0, make_int4(f1[0], f1[1], f1[2], f1[3]));
            st4(row + 4 * wr4 + c0, make_int4(f2[0], f2[1], f2[2], f2[3]));
            if (use_smem) {
                st4(bufH[0] + c0, make_int4(h[0], h[1], h[2], h[3]));
                st4(bufE1[0] + c0, make_int4(x1[0], x1[1], x1[2], x1[3]));
                st4(bufE2[0] + c0, make_int4(x2[0], x2[1], x2[2], x2[3]));
            }
        }
        if (tid == 0) { d.dp_beg[0] = 0; d.dp_end[0] = prev_end; d.row_off[0] = 0; d.row_left[0] = 0; d.row_right[0] = 0; }
        cur_off = 5LL * wr4; cells = prev_end + 1;
    }
    __syncthreads();

    int cur = 1;                               // shared-memory buffer the current row is written to
    for (int r = 1; r < R; ++r, cur ^= 1) {
        // ---- band of the row (GET_AD_DP_BEGIN/END + lane-group snap, :946-960) ----
        const int b = rt.row_base[r], p0 = rt.pre_off[r], p1 = rt.pre_off[r + 1];
        const int dd = L - rt.row_rd[r];
        int maxL = node_n, maxR = 0, min_pre_beg = 0x7fffffff;
        for (int k = p0; k < p1; ++k) {
            const int p = rt.pre_row[k];
            int pl, pr, pb;
            if (p == r - 1) { pl = prev_left; pr = prev_right; pb = prev_beg; }
            else { pl = d.row_left[p]; pr = d.row_right[p]; pb = d.dp_beg[p]; }
            maxL = min(maxL, pl + 1); maxR = max(maxR, pr + 1); min_pre_beg = min(min_pre_beg, pb);
        }
        int beg = max(0, min(maxL, dd) - w);
        const int end = min(L, max(maxR, dd) + w);
        if (beg / pn < min_pre_beg / pn) beg = min_pre_beg;
        const int beg4 = beg & ~3, wr4 = (end | 3) - beg4 + 1;
        if (cur_off + 5LL * wr4 > d.plane_cap) return -1;     // uniform across the CTA
        int *rowp = d.planes + cur_off - beg4;                  // rowp[plane*wr4 + j]
        const int *mrow = S.smat + 5 * b;

        int carry1 = NEG + beg * e1, carry2 = NEG + beg * e2;   // prefix-max carry in "A space"
        const int id1 = carry1, id2 = carry2;                   // identities of the two scans
        int tmax = NEG - 1000, tleft = 0x7fffffff, tright = -1; // thread-local row max bookkeeping
        const int npass = (wr4 + 4 * T - 1) / (4 * T);
        for (int pass = 0; pass < npass; ++pass) {
            const int c0 = beg4 + (pass * T + tid) * 4;
            const bool active = c0 <= end;
            int hme[4], x1[4], x2[4], A1[4], A2[4];
            int agg1 = id1, agg2 = id2;
            if (active) {
                int m[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) { m[t] = NEG; x1[t] = NEG; x2[t] = NEG; }
                for (int k = p0; k < p1; ++k) {
                    const int p = rt.pre_row[k];
                    int pb, pe; const int *Hs, *E1s, *E2s;
                    if (p == r - 1) { pb = prev_beg; pe = prev_end; } else { pb = d.dp_beg[p]; pe = d.dp_end[p]; }
                    if (use_smem && p == r - 1) { Hs = bufH[cur ^ 1]; E1s = bufE1[cur ^ 1]; E2s = bufE2[cur ^ 1]; }
                    else {
                        const int pb4 = pb & ~3, pw4 = (pe | 3) - pb4 + 1;
                        Hs = d.planes + (p == r - 1 ? prev_off : d.row_off[p]) - pb4; E1s = Hs + pw4; E2s = E1s + pw4;
                    }
                    if (c0 - 1 >= pb && c0 + 3 <= pe) {       // whole group inside the predecessor's band
                        const int4 h = ld4(Hs + c0); const int hl = Hs[c0 - 1];
                        const int4 a = ld4(E1s + c0), c = ld4(E2s + c0);
                        m[0] = max(m[0], hl); m[1] = max(m[1], h.x); m[2] = max(m[2], h.y); m[3] = max(m[3], h.z);
                        x1[0] = max(x1[0], a.x); x1[1] = max(x1[1], a.y); x1[2] = max(x1[2], a.z); x1[3] = max(x1[3], a.w);
                        x2[0] = max(x2[0], c.x); x2[1] = max(x2[1], c.y); x2[2] = max(x2[2], c.z); x2[3] = max(x2[3], c.w);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int j = c0 + t;
                            if (j - 1 >= pb && j - 1 <= pe) m[t] = max(m[t], Hs[j - 1]);
                            if (j >= pb && j <= pe) { x1[t] = max(x1[t], E1s[j]); x2[t] = max(x2[t], E2s[j]); }
                        }
                    }
                }
                const uint32_t q4 = *reinterpret_cast<const uint32_t *>(sq + c0);   // q_{c0..c0+3}
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = c0 + t;
                    const int s = (j == 0) ? 0 : mrow[(q4 >> (8 * t)) & 0xff];
                    if (j < beg || j > end) { hme[t] = NEG; x1[t] = NEG; x2[t] = NEG; }
                    else hme[t] = max3(m[t] + s, x1[t], x2[t]);                      // H' = max(M + s, E1, E2), :1033,1050
                    A1[t] = hme[t] - oe1 + (j + 1) * e1;
                    A2[t] = hme[t] - oe2 + (j + 1) * e2;
                    agg1 = max(agg1, A1[t]); agg2 = max(agg2, A2[t]);
                }
            }
            // ---- exclusive prefix maximum over the row: warp shuffle scan + redux over warp aggregates ----
            int inc1 = agg1, inc2 = agg2;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int n1 = __shfl_up_sync(FULL, inc1, off), n2 = __shfl_up_sync(FULL, inc2, off);
                if (lane >= off) { inc1 = max(inc1, n1); inc2 = max(inc2, n2); }
            }
            int ex1 = __shfl_up_sync(FULL, inc1, 1), ex2 = __shfl_up_sync(FULL, inc2, 1);
            if (lane == 0) { ex1 = id1; ex2 = id2; }
            if (lane == 31) { S.wF[pass & 1][0][warp] = inc1; S.wF[pass & 1][1][warp] = inc2; }
            __syncthreads();
            const int wv1 = lane < nwarps ? S.wF[pass & 1][0][lane] : id1, wv2 = lane < nwarps ? S.wF[pass & 1][1][lane] : id2;
            const int before1 = __reduce_max_sync(FULL, lane < warp ? wv1 : id1), before2 = __reduce_max_sync(FULL, lane < warp ? wv2 : id2);
            const int tot1 = __reduce_max_sync(FULL, wv1), tot2 = __reduce_max_sync(FULL, wv2);
            int P1 = max3(carry1, before1, ex1), P2 = max3(carry2, before2, ex2);
            carry1 = max(carry1, tot1); carry2 = max(carry2, tot2);
            if (active) {
                int h[4], f1[4], f2[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = c0 + t;
                    f1[t] = P1 - j * e1; f2[t] = P2 - j * e2;                       // F[j] = max_{k<j} A[k] - j*e
                    P1 = max(P1, A1[t]); P2 = max(P2, A2[t]);
                    if (j < beg || j > end) { h[t] = NEG; f1[t] = NEG; f2[t] = NEG; }
                    else {
                        h[t] = max3(hme[t], f1[t], f2[t]);                           // :1067
                        x1[t] = max(x1[t] - e1, h[t] - oe1);                         // E for the next rows, :1070-1071
                        x2[t] = max(x2[t] - e2, h[t] - oe2);
                        if (h[t] > tmax) { tmax = h[t]; tleft = j; tright = j; } else if (h[t] == tmax) tright = j;
                    }
                }
                st4(rowp + c0, make_int4(h[0], h[1], h[2], h[3]));
                st4(rowp + wr4 + c0, make_int4(x1[0], x1[1], x1[2], x1[3]));
                st4(rowp + 2 * wr4 + c0, make_int4(x2[0], x2[1], x2[2], x2[3]));
                st4(rowp + 3 * wr4 + c0, make_int4(f1[0], f1[1], f1[2], f1[3]));
                st4(rowp + 4 * wr4 + c0, make_int4(f2[0], f2[1], f2[2], f2[3]));
                if (use_smem) {
                    st4(bufH[cur] + c0, make_int4(h[0], h[1], h[2], h[3]));
                    st4(bufE1[cur] + c0, make_int4(x1[0], x1[1], x1[2], x1[3]));
                    st4(bufE2[cur] + c0, make_int4(x2[0], x2[1], x2[2], x2[3]));
                }
            }
        }
        // ---- left/right-most argmax of H over the band (simd_abpoa_max_in_row, :1107-1119) ----
        {
            const int wmax = __reduce_max_sync(FULL, tmax);
            const int wl = __reduce_min_sync(FULL, tmax == wmax ? tleft : 0x7fffffff);
            const int wr = __reduce_max_sync(FULL, tmax == wmax ? tright : -1);
            if (lane == 0) { S.wM[0][warp] = wmax; S.wM[1][warp] = wl; S.wM[2][warp] = wr; }
        }
        __syncthreads();
        {
            const int v = lane < nwarps ? S.wM[0][lane] : NEG - 1000;
            const int bmax = __reduce_max_sync(FULL, v);
            const int l = (lane < nwarps && v == bmax) ? S.wM[1][lane] : 0x7fffffff;
            const int rr = (lane < nwarps && v == bmax) ? S.wM[2][lane] : -1;
            prev_left = __reduce_min_sync(FULL, l); prev_right = __reduce_max_sync(FULL, rr);
        }
        prev_beg = beg; prev_end = end;
        if (tid == 0) { d.dp_beg[r] = beg; d.dp_end[r] = end; d.row_off[r] = cur_off; d.row_left[r] = prev_left; d.row_right[r] = prev_right; }
        prev_off = cur_off; cur_off += 5LL * wr4; cells += end - beg + 1;
    }
    __syncthreads();
    return cells;
}

__device__ __forceinline__ void carve(KShared &S, const BatchArgs &A, int slot) {
    const SlotLayout &Y = A.lay;
    uint8_t *b = A.slots + (int64_t)slot * Y.slot_bytes;
    Graph &g = S.g; RowTables &rt = S.rt; DpState &d = S.d;
    g.node_cap = Y.node_cap; g.in_pool = Y.in_pool; g.out_pool = Y.out_pool;
    g.base = b + Y.o_base; g.aln_n = b + Y.o_aln_n; g.aln_id = (int *)(b + Y.o_aln_id);
    g.in_off = (int *)(b + Y.o_in_off); g.in_n = (int *)(b + Y.o_in_n); g.in_cap = (int *)(b + Y.o_in_cap);
    g.out_off = (int *)(b + Y.o_out_off); g.out_n = (int *)(b + Y.o_out_n); g.out_cap = (int *)(b + Y.o_out_cap);
    g.in_id = (int *)(b + Y.o_in_id); g.in_w = (int *)(b + Y.o_in_w);
    g.out_id = (int *)(b + Y.o_out_id); g.out_w = (int *)(b + Y.o_out_w); g.out_rid = (uint64_t *)(b + Y.o_out_rid);
    g.index_to_node = (int *)(b + Y.o_index_to_node); g.node_to_index = (int *)(b + Y.o_node_to_index);
    g.remain = (int *)(b + Y.o_remain); g.msa_rank = (int *)(b + Y.o_msa_rank);
    g.tmp0 = (int *)(b + Y.o_tmp0); g.tmp1 = (int *)(b + Y.o_tmp1);
    rt.row_base = b + Y.o_row_base; rt.row_rd = (int *)(b + Y.o_row_rd);
    rt.pre_off = (int *)(b + Y.o_pre_off); rt.pre_row = (int *)(b + Y.o_pre_row);
    d.planes = A.planes + (int64_t)slot * Y.plane_cap; d.plane_cap = Y.plane_cap;
    d.row_off = (int64_t *)(b + Y.o_row_off); d.dp_beg = (int *)(b + Y.o_dp_beg); d.dp_end = (int *)(b + Y.o_dp_end);
    d.row_left = (int *)(b + Y.o_row_left); d.row_right = (int *)(b + Y.o_row_right);
    d.cigar = (uint64_t *)(b + Y.o_cigar); d.cigar_cap = Y.cigar_cap; d.n_cigar = 0;
}

// abpoa_topological_sort (abpoa_graph.c:322-357): BFS index (serial), edge sort (one node per thread),
// max_remain + row tables (serial).
__device__ void topo_sort_cta(KShared &S) {
    Graph &g = S.g;
    if (threadIdx.x == 0) graph_bfs_index(g);
    __syncthreads();
    if (g.err) return;
    for (int v = threadIdx.x; v < g.node_n; v += blockDim.x) graph_sort_node_edges(g, v);
    __syncthreads();
    if (threadIdx.x == 0) { graph_bfs_remain(g); if (!g.err) graph_build_rows(g, S.rt); }
    __syncthreads();
}

#define PHASE_TICK(ph) do { if (A.phase_clk && threadIdx.x == 0) { unsigned long long _n = clock64(); A.phase_clk[(size_t)blockIdx.x * PH_N + (ph)] += _n - t_last; t_last = _n; } } while (0)

extern "C" __global__ void __launch_bounds__(512, 2) poa_msa_kernel(const BatchArgs A) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ KShared S;
    const int tid = threadIdx.x;
    if (tid == 0) carve(S, A, blockIdx.x);
    if (tid < 25) S.smat[tid] = A.P.mat[tid];
    // dynamic shared memory: query bytes, then the double-buffered previous-row cache
    const int qbytes = (A.q_cols + 8 + 15) & ~15;
    uint8_t *sq = dyn_smem;
    int *rowbuf = reinterpret_cast<int *>(dyn_smem + qbytes);
    const int rb_stride = A.smem_cols + 8;
    unsigned long long t_last = A.phase_clk ? clock64() : 0ULL, t_start = t_last;
    __syncthreads();

    while (true) {
        if (tid == 0) S.job = atomicAdd(A.next_job, 1);
        __syncthreads();
        const int job = S.job;
        if (job >= A.n_jobs) break;
        const JobDesc jd = A.jobs[job];
        const int K = jd.n_seq;
        const int *lens = A.lens + jd.len_off, *order = A.order + jd.len_off;
        const int64_t *soff = A.soff + jd.len_off;
        const uint8_t *seqs = A.seqs + jd.seq_off;
        if (tid == 0) { graph_reset(S.g, K); S.abort_s = 0; }
        __syncthreads();
        long long cells = 0;
        for (int a = 0; a < K; ++a) {
            const int read = order[a], L = lens[read];
            const uint8_t *q = seqs + soff[read];
            if (a == 0) {
                if (tid == 0) graph_add_first_sequence(S.g, q, L, read);
                PHASE_TICK(PH_FUSE);
            } else {
                // queries longer than the shared-memory row cache read their predecessors from global memory (L2)
                const bool fits = L + 1 <= A.smem_cols;
                long long c = dp_sweep(S, A.P, q, L, sq, rowbuf, rb_stride, fits);
                if (c < 0) { if (tid == 0) S.g.err = JOB_ERR_PLANE_CAP; c = 0; }
                cells += c;
                __syncthreads();
                PHASE_TICK(PH_DP);
                if (tid == 0 && !S.g.err) { dp_best_cell(S.g, S.rt, S.d, A.P, L); dp_backtrack(S.g, S.rt, S.d, A.P, q, L); }
                PHASE_TICK(PH_BACKTRACK);
                if (tid == 0 && !S.g.err) graph_fuse_alignment(S.g, q, S.d.cigar, S.d.n_cigar, read);
                PHASE_TICK(PH_FUSE);
            }
            __syncthreads();
            if (S.g.err) break;
            topo_sort_cta(S);
            PHASE_TICK(PH_TOPO);
            if (S.g.err) break;
        }
        __syncthreads();
        // ---- MSA (abpoa_generate_rc_msa, abpoa_output.c:149-176) ----
        if (tid == 0 && !S.g.err) {
            S.msa_len_s = graph_msa_rank(S.g);
            if (!S.g.err && S.msa_len_s > jd.msa_stride) S.g.err = JOB_ERR_MSA_CAP;
        }
        __syncthreads();
        if (!S.g.err) {
            const int ml = S.msa_len_s;
            uint8_t *msa = A.msa + jd.msa_off;
            for (int64_t i = tid; i < (int64_t)K * ml; i += blockDim.x) msa[(i / ml) * jd.msa_stride + (i % ml)] = GAP_CODE;
            __syncthreads();
            for (int v = 2 + tid; v < S.g.node_n; v += blockDim.x) graph_msa_fill_node(S.g, v, msa, jd.msa_stride);
        }
        __syncthreads();
        if (tid == 0) { A.status[job] = S.g.err; A.msa_len[job] = S.g.err ? 0 : S.msa_len_s; A.cells[job] = cells; }
        PHASE_TICK(PH_MSA);
        __syncthreads();
    }
    if (A.phase_clk && tid == 0) A.phase_clk[(size_t)blockIdx.x * PH_N + PH_TOTAL] += clock64() - t_start;
}

}  // namespace barb200
