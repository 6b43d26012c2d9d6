// poa_kernel.cu -- the fused, batched partial-order-alignment kernel for sm_100a.
//
// One CTA owns one job (= one abpoa_msa call of the reference: K sequences -> one MSA) from start to finish:
// it keeps the job's partial-order graph in its slot of device memory, aligns the sequences to it one after the
// other and never returns to the host in between. CTAs are persistent: each pulls the next job index from a
// global counter, so a launch of (148 SMs x resident CTAs) blocks streams through thousands of jobs while the
// strictly sequential phases of some jobs (traceback, graph fusion, topological sort) overlap the DP sweeps of
// the other CTAs resident on the same SM.
//
// DP sweep (the hot loop; replaces simd_abpoa_cg_dp + first row + row max + adaptive band,
// abPOA src/abpoa_align_simd.c:617-688, 935-1130):
//   * graph rows in topological order, strictly one after the other (the adaptive band of a row needs the argmax
//     columns of all predecessor rows); the columns of a row in parallel;
//   * column ownership is FIXED: thread t owns columns [16t, 16t+16) of every row. The H / E1 / E2 values of the
//     row just computed therefore stay in the owning thread's registers and feed the next row without touching
//     memory when the predecessor is the previous row (the common case in a near-linear graph); only H[16t-1]
//     comes from the neighbour thread (warp shuffle; one shared-memory word per warp boundary). Predecessors
//     further back are read from the planes in global memory (L2), coalesced 16 B per thread;
//   * the max-plus recurrence of the two insertion states F1/F2 along the row is turned into a plain prefix
//     maximum by the substitution A[k] = H'[k] - oe + (k+1)*e  =>  F[j] = max_{k<j} A[k] - j*e: 16 serial cells per
//     thread, a warp shuffle scan over the 32 thread aggregates, a redux over the warp aggregates staged in
//     shared memory;
//   * per cell the sweep writes 8 bytes for the traceback and for later rows -- H (int32) and the two E values as 16-bit
//     distances below H; F1 / F2 are not stored, the traceback recomputes the few row prefixes it needs
//     (poa_types.h: DpState) -- in a thread-blocked layout that makes the 256-bit stores of a warp one contiguous run:
//     8 B/cell of HBM write traffic (the reference streams five int32 planes, 20 B/cell) is the kernel's only DRAM stream;
//   * two block barriers per row.
// Integer DP: no tensor cores. int32 everywhere with the reference's own "minus infinity" so that finite cells
// are bit-identical to abPOA's AVX2 path.
#include <cuda_runtime.h>
#include <stdint.h>
#include "poa_graph.cuh"
#include "poa_cta.cuh"
#include "poa_kernel.cuh"

namespace barb200 {

#define FULL 0xffffffffu

struct KShared {
    Graph g; RowTables rt; DpState d;
    int job, msa_len_s, abort_s;
    int smat[5 * 8];       // [graph base][query code 0..4, 5 = "no base": column 0 / beyond the query -> 0]
    int wF[2][2][32];      // [row parity][plane F1/F2][warp] block scan staging
    int wM[2][4][32];      // [row parity][max, leftmost, rightmost, H of the warp's last column][warp]
};

__device__ __forceinline__ int4 ld4cg(const int *p) { return __ldcg(reinterpret_cast<const int4 *>(p)); }
__device__ __forceinline__ void st4(int *p, int a, int b, int c, int d) { *reinterpret_cast<int4 *>(p) = make_int4(a, b, c, d); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }
// 256-bit global accesses (sm_100: STG.E.256 / LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void st8(int *p, int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7) : "memory");
}
struct int8v { int v[8]; };
__device__ __forceinline__ int8v ld8cg(const int *p) {
    int8v r;
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p) : "memory");
    return r;
}


// ---- the two per-thread passes over a row's 16 cells, in a branch-free form -------------------------------------
// "A space" of the F scans: A1'[k] = H'[k] + k*e1 (the constant -o1 is applied when F is read back), so
// F1[j] = max_{k<j} A1'[k] - o1 - j*e1 and the scan identity is inf_min + beg*e1 + o1. MASKED = some of the thread's
// columns lie outside [beg, end] (they must come out as exactly inf_min).
template <bool MASKED>
__device__ __forceinline__ void row_pass1(int (&H)[CPT], int (&E1)[CPT], int (&E2)[CPT], const int *mrow, const uint32_t (&qc)[2],
                                          int j0, int beg, int end, int NEG, int je1, int je2, int e1, int e2, int &agg1, int &agg2) {
#pragma unroll
    for (int e = 0; e < CPT; ++e) {
        const int s = mrow[(qc[e >> 3] >> ((e & 7) * 4)) & 7];
        int h = __vimax3_s32(H[e] + s, E1[e], E2[e]);                    // H' = max(M + s, E1, E2), :1033,1050
        if (MASKED) {
            const bool inb = (unsigned)(j0 + e - beg) <= (unsigned)(end - beg);
            h = inb ? h : NEG; E1[e] = inb ? E1[e] : NEG; E2[e] = inb ? E2[e] : NEG;
        }
        H[e] = h;
        agg1 = __viaddmax_s32(h, je1 + e * e1, agg1); agg2 = __viaddmax_s32(h, je2 + e * e2, agg2);
    }
}

// MODE 0: all of the warp's active columns are inside the band; 1: some lie left (or left and right) of it -- every value
// of an outside cell is forced to inf_min; 2: some lie right of it only -- nothing in the band depends on those cells
// and the traceback never reads the F planes outside the band, so only H / E1 / E2 are forced (they feed later rows).
template <int MODE>
__device__ __forceinline__ void row_pass2(int (&H)[CPT], int (&E1)[CPT], int (&E2)[CPT], int P1, int P2, int j0, int beg, int end, int NEG,
                                          int je1, int je2, int e1, int e2, int o1, int o2, int *tp, int &tmax) {
    const int oe1 = o1 + e1, oe2 = o2 + e2;
    // (256-bit stores of whole 32-byte sectors: 128-bit halves cost ~20 % of the kernel's throughput in L2 write merging)
#pragma unroll
    for (int oc = 0; oc < 2; ++oc) {
        int dd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = oc * 8 + u, u1 = je1 + e * e1, u2 = je2 + e * e2;
            const int f1 = P1 - o1 - u1, f2 = P2 - o2 - u2;                // F[j] = max_{k<j} A'[k] - o - j*e (used, not stored)
            P1 = __viaddmax_s32(H[e], u1, P1); P2 = __viaddmax_s32(H[e], u2, P2);
            int h = __vimax3_s32(H[e], f1, f2);                            // :1067
            int x1 = __viaddmax_s32(E1[e], -e1, h - oe1);                  // E for the next rows, :1070-1071
            int x2 = __viaddmax_s32(E2[e], -e2, h - oe2);
            if (MODE == 1) {
                const bool inb = (unsigned)(j0 + e - beg) <= (unsigned)(end - beg);
                h = inb ? h : NEG; x1 = inb ? x1 : NEG; x2 = inb ? x2 : NEG;
            } else if (MODE == 2) {
                const bool inb = j0 + e <= end;
                h = inb ? h : NEG; x1 = inb ? x1 : NEG; x2 = inb ? x2 : NEG;
            }
            H[e] = h; E1[e] = x1; E2[e] = x2;
            dd[u] = (h - x1) | ((h - x2) << 16);                           // e <= H - E' <= oe < 65535 (0 outside the band)
            tmax = max(tmax, h);
        }
        st8(tp + CPT + oc * 8, dd[0], dd[1], dd[2], dd[3], dd[4], dd[5], dd[6], dd[7]);
    }
#pragma unroll
    for (int oc = 0; oc < 2; ++oc)
        st8(tp + oc * 8, H[oc * 8], H[oc * 8 + 1], H[oc * 8 + 2], H[oc * 8 + 3], H[oc * 8 + 4], H[oc * 8 + 5], H[oc * 8 + 6], H[oc * 8 + 7]);
}

// left/right-most column of the thread's cells that attain v (only in-band cells count)
template <bool MASKED>
__device__ __forceinline__ void row_argmax(const int (&H)[CPT], int v, int j0, int beg, int end, int &tl, int &tr) {
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < CPT; ++e) m |= H[e] == v ? 1u << e : 0u;
    if (MASKED) {
        const int lo = max(beg - j0, 0), hi = min(end - j0, CPT - 1);
        m &= (2u << hi) - (1u << lo);
    }
    if (m) { tl = j0 + __ffs(m) - 1; tr = j0 + 31 - __clz(m); }
}

// ---------------------------------------------------------------------------------------------------------
// banded convex-gap DP of query q[1..L] against the sorted graph. All threads of the CTA; L + 1 <= 16 * blockDim.x.
// Returns the number of banded cells (sum of dp_end-dp_beg+1), or -1 if the planes outgrew the slot.
// ---------------------------------------------------------------------------------------------------------
__device__ long long dp_sweep(KShared &S, const BatchArgs &A, const uint8_t *__restrict__ qg, int L) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const PoaParams &P = A.P;
    // slot arrays addressed from the kernel parameters (not through the pointers cached in shared memory), so that the
    // compiler knows they are global memory and emits LDG/STG instead of generic accesses
    uint8_t *const sb = A.slots + (int64_t)blockIdx.x * A.lay.slot_bytes;
    const RowRec *const rec_tab = reinterpret_cast<const RowRec *>(sb + A.lay.o_row_rec);
    const int *const pre_row = reinterpret_cast<const int *>(sb + A.lay.o_pre_row);
    RowInfo *const info = reinterpret_cast<RowInfo *>(sb + A.lay.o_row_info);
    int64_t *const row_off = reinterpret_cast<int64_t *>(sb + A.lay.o_row_off);
    int *const planes = A.planes + (int64_t)blockIdx.x * A.lay.plane_cap;
    const int64_t plane_cap = A.lay.plane_cap;
    const int node_n = S.g.node_n, R = node_n - 1;
    const int NEG = P.inf_min, e1 = P.e1, e2 = P.e2, oe1 = P.o1 + P.e1, oe2 = P.o2 + P.e2;
    const int w = P.wb + (int)(P.wf * L);                                    // abpoa_align_simd.c:474
    const int pn_shift = reference_lane_count(P, L, node_n) == 16 ? 4 : 3;
    static_assert(CPT == 16, "shifts below assume 16 columns per thread");
    const int j0 = tid * CPT, je1 = j0 * e1, je2 = j0 * e2;

    // query codes of my 16 columns, 4 bits each (column j scores against q_j = qg[j-1]; column 0 and columns past the
    // query score 0, abpoa_align_simd.c:536)
    uint32_t qc[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < CPT; ++e) {
        const int j = j0 + e;
        const uint32_t c = (j >= 1 && j <= L) ? qg[j - 1] : 5u;
        qc[e >> 3] |= c << ((e & 7) * 4);
    }

    int H[CPT], E1[CPT], E2[CPT];          // the previous row's values of my columns (valid iff prev_active)
    bool prev_active;
    int prev_beg = 0, prev_end, prev_left = 0, prev_right = 0;
    int cur_blk = 0, cells = 0;               // row blocks (of TB ints) written so far; banded cells so far

    // ---- row 0 (simd_abpoa_cg_first_dp, :617-688) ----
    {
        const int dd = L - rec_tab[0].rd;
        prev_end = min(L, max(0, dd) + w);
        const int nT = prev_end / CPT + 1;
        if ((int64_t)nT * TB > plane_cap) return -1;
        prev_active = tid < nT;
        if (prev_active) {
            int *tp = planes + tid * TB;
            int dd[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) {
                const int j = j0 + e;
                if (j == 0) { H[e] = 0; E1[e] = -oe1; E2[e] = -oe2; dd[e] = oe1 | (oe2 << 16); }
                else if (j <= prev_end) { H[e] = max(-P.o1 - e1 * j, -P.o2 - e2 * j); E1[e] = NEG; E2[e] = NEG; dd[e] = E_NEG16 | (E_NEG16 << 16); }
                else { H[e] = E1[e] = E2[e] = NEG; dd[e] = 0; }
            }
#pragma unroll
            for (int oc = 0; oc < 2; ++oc) {
                st8(tp + oc * 8, H[oc * 8], H[oc * 8 + 1], H[oc * 8 + 2], H[oc * 8 + 3], H[oc * 8 + 4], H[oc * 8 + 5], H[oc * 8 + 6], H[oc * 8 + 7]);
                st8(tp + CPT + oc * 8, dd[oc * 8], dd[oc * 8 + 1], dd[oc * 8 + 2], dd[oc * 8 + 3], dd[oc * 8 + 4], dd[oc * 8 + 5], dd[oc * 8 + 6], dd[oc * 8 + 7]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < CPT; ++e) H[e] = E1[e] = E2[e] = NEG;
        }
        if (lane == 31) S.wM[0][3][warp] = prev_active ? H[CPT - 1] : NEG;
        if (tid == 0) { RowInfo ri; ri.beg = 0; ri.end = prev_end; ri.left = 0; ri.right = 0; info[0] = ri; row_off[0] = 0; }
        cur_blk = nT; cells = prev_end + 1;
    }
    __syncthreads();

    RowRec rec = rec_tab[R > 1 ? 1 : 0];
    for (int r = 1; r < R; ++r) {
        const int par = r & 1;
        const RowRec nrec = rec_tab[r + 1 < R ? r + 1 : r];      // next row's record, in flight while this row computes
        // ---- band of the row (GET_AD_DP_BEGIN/END + lane-group snap, :946-960) ----
        const int b = rec.base_npre & 0xff, npre = rec.base_npre >> 8;
        const int dd = L - rec.rd;
        int maxL = node_n, maxR = 0, min_pre_beg = 0x7fffffff;
        bool has_prev = false;
        if (npre == 1 && rec.pre0 == r - 1) {                       // the linear-chain case: everything is in registers
            maxL = min(node_n, prev_left + 1); maxR = prev_right + 1; min_pre_beg = prev_beg; has_prev = true;   // max_pos_left starts at node_n (abpoa_graph.c:347-352)
        } else {
#pragma unroll 1
            for (int k = 0; k < npre; ++k) {
                const int p = k == 0 ? rec.pre0 : pre_row[rec.pre_off + k];
                int pl, pr, pb;
                if (p == r - 1) { pl = prev_left; pr = prev_right; pb = prev_beg; has_prev = true; }
                else { const RowInfo pi = info[p]; pl = pi.left; pr = pi.right; pb = pi.beg; }
                maxL = min(maxL, pl + 1); maxR = max(maxR, pr + 1); min_pre_beg = min(min_pre_beg, pb);
            }
        }
        const bool only_prev = npre == 1 && has_prev;
        int beg = max(0, min(maxL, dd) - w);
        const int end = min(L, max(maxR, dd) + w);
        if ((beg >> pn_shift) < (min_pre_beg >> pn_shift)) beg = min_pre_beg;
        const int t0 = beg >> 4, nT = (end >> 4) - t0 + 1, tt = tid - t0;
        if ((int64_t)(cur_blk + nT) * TB > plane_cap) return -1;        // uniform across the CTA
        const bool active = tt >= 0 && tt < nT;
        // masking is decided per WARP (no divergent double execution): 0 = all active threads inside the band,
        // 1 = some columns left of the band, 2 = some columns right of it only
        const int wmode = __any_sync(FULL, active && j0 < beg) ? 1 : __any_sync(FULL, active && j0 + CPT - 1 > end) ? 2 : 0;
        const int *mrow = S.smat + 8 * b;

        // H of the column left of my first one, previous row
        int hl = __shfl_up_sync(FULL, prev_active ? H[CPT - 1] : NEG, 1);
        if (lane == 0) hl = warp > 0 ? S.wM[par ^ 1][3][warp - 1] : NEG;

        const int id1 = NEG + beg * e1 + P.o1, id2 = NEG + beg * e2 + P.o2;   // identities of the two scans ("A space")
        int agg1 = id1, agg2 = id2;
        if (active) {
            // M candidates: H[e] <- H_pred[e-1]; E candidates stay in E1/E2 (previous row = predecessor case)
            if (has_prev && prev_active) {
#pragma unroll
                for (int e = CPT - 1; e >= 1; --e) H[e] = H[e - 1];
                H[0] = hl;
            } else {
#pragma unroll
                for (int e = 0; e < CPT; ++e) { H[e] = NEG; E1[e] = NEG; E2[e] = NEG; }
                if (has_prev) H[0] = hl;
            }
            // predecessors further back: from the planes in global memory
#pragma unroll 1
            for (int k = only_prev ? npre : 0; k < npre; ++k) {
                const int p = k == 0 ? rec.pre0 : pre_row[rec.pre_off + k];
                if (p == r - 1) continue;
                const RowInfo pi = info[p];
                const int pt0 = pi.beg >> 4, pnT = (pi.end >> 4) - pt0 + 1, ptt = tid - pt0;
                const int *Hp = planes + row_off[p] + (int64_t)ptt * TB;      // my block of row p (if stored)
                if (ptt >= 0 && ptt < pnT) {
#pragma unroll
                    for (int oc = 0; oc < 2; ++oc) {
                        const int8v h = ld8cg(Hp + oc * 8), dv = ld8cg(Hp + CPT + oc * 8);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int e = oc * 8 + u;
                            const int c1 = dv.v[u] & 0xffff, c2 = (int)((unsigned)dv.v[u] >> 16);
                            if (e + 1 < CPT) H[e + 1] = max(H[e + 1], h.v[u]);
                            E1[e] = max(E1[e], c1 == E_NEG16 ? NEG : h.v[u] - c1); E2[e] = max(E2[e], c2 == E_NEG16 ? NEG : h.v[u] - c2);
                        }
                    }
                }
                if (ptt >= 1 && ptt <= pnT) H[0] = max(H[0], __ldcg(Hp - TB + CPT - 1));    // H[16*tid - 1]: last H of the left neighbour's block
            }
            if (wmode != 1) row_pass1<false>(H, E1, E2, mrow, qc, j0, beg, end, NEG, je1, je2, e1, e2, agg1, agg2);
            else row_pass1<true>(H, E1, E2, mrow, qc, j0, beg, end, NEG, je1, je2, e1, e2, agg1, agg2);
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
        if (lane == 31) { S.wF[par][0][warp] = inc1; S.wF[par][1][warp] = inc2; }
        __syncthreads();
        int P1, P2;
        {
            const int wv1 = lane < warp ? S.wF[par][0][lane] : id1, wv2 = lane < warp ? S.wF[par][1][lane] : id2;
            P1 = max(__reduce_max_sync(FULL, wv1), ex1); P2 = max(__reduce_max_sync(FULL, wv2), ex2);
        }
        int tmax = NEG - 1000;
        if (active) {
            int *tp = planes + (int64_t)(cur_blk + tt) * TB;
            if (wmode == 0) row_pass2<0>(H, E1, E2, P1, P2, j0, beg, end, NEG, je1, je2, e1, e2, P.o1, P.o2, tp, tmax);
            else if (wmode == 2) row_pass2<2>(H, E1, E2, P1, P2, j0, beg, end, NEG, je1, je2, e1, e2, P.o1, P.o2, tp, tmax);
            else row_pass2<1>(H, E1, E2, P1, P2, j0, beg, end, NEG, je1, je2, e1, e2, P.o1, P.o2, tp, tmax);
        }
        // ---- left/right-most argmax of H over the band (simd_abpoa_max_in_row, :1107-1119) ----
        {
            const int wmax = __reduce_max_sync(FULL, tmax);
            int tleft = 0x7fffffff, tright = -1;
            if (active && tmax == wmax) {
                if (wmode == 0) row_argmax<false>(H, wmax, j0, beg, end, tleft, tright); else row_argmax<true>(H, wmax, j0, beg, end, tleft, tright);
            }
            const int wl = __reduce_min_sync(FULL, tleft), wr = __reduce_max_sync(FULL, tright);
            if (lane == 0) { S.wM[par][0][warp] = wmax; S.wM[par][1][warp] = wl; S.wM[par][2][warp] = wr; }
            if (lane == 31) S.wM[par][3][warp] = active ? H[CPT - 1] : NEG;
        }
        __syncthreads();
        {
            const int v = lane < nwarps ? S.wM[par][0][lane] : NEG - 1000;
            const int bmax = __reduce_max_sync(FULL, v);
            const int l = (lane < nwarps && v == bmax) ? S.wM[par][1][lane] : 0x7fffffff;
            const int rr = (lane < nwarps && v == bmax) ? S.wM[par][2][lane] : -1;
            prev_left = __reduce_min_sync(FULL, l); prev_right = __reduce_max_sync(FULL, rr);
        }
        prev_beg = beg; prev_end = end; prev_active = active;
        if (tid == 0) { RowInfo ri; ri.beg = beg; ri.end = end; ri.left = prev_left; ri.right = prev_right; info[r] = ri; row_off[r] = (int64_t)cur_blk * TB; }
        cur_blk += nT; cells += end - beg + 1;
        rec = nrec;
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
    rt.rec = (RowRec *)(b + Y.o_row_rec); rt.pre_row = (int *)(b + Y.o_pre_row);
    d.planes = A.planes + (int64_t)slot * Y.plane_cap; d.plane_cap = Y.plane_cap;
    d.row_off = (int64_t *)(b + Y.o_row_off); d.info = (RowInfo *)(b + Y.o_row_info);
    d.cigar = (uint64_t *)(b + Y.o_cigar); d.cigar_cap = Y.cigar_cap; d.n_cigar = 0;
    d.fc = (int *)(b + Y.o_fc); d.fc_cap = Y.fc_cap; d.fc_row = -1; d.fc_hi = -1;
}

#define PHASE_TICK(ph) do { if (A.phase_clk && threadIdx.x == 0) { unsigned long long _n = clock64(); A.phase_clk[(size_t)blockIdx.x * PH_N + (ph)] += _n - t_last; t_last = _n; } } while (0)

__device__ __forceinline__ void poa_msa_body(const BatchArgs &A) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];    // scratch of the topological sort (poa_cta.cuh)
    __shared__ KShared S;
    const int tid = threadIdx.x;
    int *ws = &S.wF[0][0][0];
    if (tid == 0) carve(S, A, blockIdx.x);
    for (int k = tid; k < 40; k += blockDim.x) S.smat[k] = (k & 7) < 5 ? A.P.mat[(k >> 3) * 5 + (k & 7)] : 0;
    unsigned long long t_last = A.phase_clk ? clock64() : 0ULL, t_start = t_last;
    __syncthreads();

    while (true) {
        if (tid == 0) {
            const int k = atomicAdd(A.next_job, 1);
            int j = -1;
            if (k < A.n_jobs) j = A.job_base + k;
            S.job = j;
        }
        __syncthreads();
        const int job = S.job;
        if (job < 0) break;
        const JobDesc jd = A.jobs[job];
        const int K = jd.n_seq;
        const int *lens = A.lens + jd.len_off;
        const int64_t *soff = A.soff + jd.len_off;
        const uint8_t *seqs = A.seqs + jd.seq_off;
        if (tid == 0) { graph_reset(S.g, K); S.abort_s = 0; }
        __syncthreads();
        // read order: the job's guide tree (abpoa_seed.c:705-722), computed by guide_tree_kernel before this launch (guide_tree.cu)
        const int *const order = A.order + jd.len_off;
        if (tid == 0 && A.gt_status[job]) S.g.err = JOB_ERR_GT_CAP;
        __syncthreads();
        long long cells = 0;
        for (int a = 0; a < K && !S.g.err; ++a) {
            const int read = order[a], L = lens[read];
            const uint8_t *q = seqs + soff[read];
            if (a == 0) {
                if (A.serial_phases) { if (tid == 0) graph_add_first_sequence(S.g, q, L, read); }
                else cta_add_first_sequence(S.g, q, L, read);
                PHASE_TICK(PH_FUSE);
            } else {
                long long c = -2;
                if (L + 1 <= CPT * (int)blockDim.x) c = dp_sweep(S, A, q, L);
                if (c < 0) { if (tid == 0) S.g.err = c == -2 ? JOB_ERR_QUERY_LEN : JOB_ERR_PLANE_CAP; c = 0; }
                cells += c;
                __syncthreads();
                PHASE_TICK(PH_DP);
                if (A.serial_phases) { if (tid == 0 && !S.g.err) { dp_best_cell(S.g, S.rt, S.d, A.P, L); dp_backtrack(S.g, S.rt, S.d, A.P, q, L); } }
                else if (tid < 32 && !S.g.err) warp_backtrack(S.g, S.rt, S.d, A.P, S.smat, q, L);
                PHASE_TICK(PH_BACKTRACK);
                __syncthreads();
                if (A.serial_phases) { if (tid == 0 && !S.g.err) graph_fuse_alignment(S.g, q, S.d.cigar, S.d.n_cigar, read); }
                else if (!S.g.err) cta_fuse_alignment(S.g, q, L, S.d.cigar, S.d.n_cigar, read, ws);
                PHASE_TICK(PH_FUSE);
            }
            __syncthreads();
            if (S.g.err) break;
            if (A.serial_phases) { if (tid == 0) graph_topo_sort_serial(S.g, S.rt); __syncthreads(); }
            else cta_topo_sort(S.g, S.rt, dyn_smem, A.scratch_bytes, ws, A.bfs_order == 0);
            PHASE_TICK(PH_TOPO);
            if (S.g.err) break;
        }
        __syncthreads();
        // ---- MSA (abpoa_generate_rc_msa, abpoa_output.c:149-176) ----
        if (!S.g.err) {
            if (A.serial_phases) { if (tid == 0) S.msa_len_s = graph_msa_rank(S.g); __syncthreads(); }
            else cta_msa_rank(S.g, dyn_smem, A.scratch_bytes, ws, &S.msa_len_s);
            if (tid == 0 && !S.g.err && S.msa_len_s > jd.msa_stride) S.g.err = JOB_ERR_MSA_CAP;
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

// One entry point per CTA-size class (the register budget per thread follows from the launch bounds):
//   queries up to 511 bases -> one warp, 16 CTAs per SM;  up to 1023 -> 64 threads;
//   up to 2047 -> 128 threads, 4 CTAs per SM;  up to 4095 -> 256 threads, 2 per SM;
//   up to 10239 (covers Cactus' 10 kbp window) -> 640 threads;  up to 16383 -> 1024 threads.
extern "C" __global__ void __launch_bounds__(32, 16) poa_msa_kernel_t32(const BatchArgs A) { poa_msa_body(A); }
extern "C" __global__ void __launch_bounds__(64, 8) poa_msa_kernel_t64(const BatchArgs A) { poa_msa_body(A); }
#ifndef BARB200_T128_MINB
#define BARB200_T128_MINB 4
#endif
extern "C" __global__ void __launch_bounds__(128, BARB200_T128_MINB) poa_msa_kernel_t128(const BatchArgs A) { poa_msa_body(A); }
extern "C" __global__ void __launch_bounds__(256, 2) poa_msa_kernel_t256(const BatchArgs A) { poa_msa_body(A); }
extern "C" __global__ void __launch_bounds__(640, 1) poa_msa_kernel_t640(const BatchArgs A) { poa_msa_body(A); }
extern "C" __global__ void __launch_bounds__(1024, 1) poa_msa_kernel_t1024(const BatchArgs A) { poa_msa_body(A); }

}  // namespace barb200
