// pecan_cta.cuh -- the banded 5-state pair-HMM forward / backward / posterior program of ONE alignment job, written for
// one thread block. It replaces, for a batch of jobs, what the reference does serially in
//   getPosteriorProbsWithBanding ............ submodules/cPecan/impl/pairwiseAligner.c:766-887
//   diagonalCalculation / cellCalculate ..... pairwiseAligner.c:619-634, stateMachine.c:450-480
//   diagonalCalculationTotalProbability ..... pairwiseAligner.c:646-663 (recomputed every 10th diagonal, :840-848)
//   diagonalCalculationPosteriorMatchProbs .. pairwiseAligner.c:676-699
//   logAdd / lookup ......................... pairwiseAligner.c:297-317
//
// Formulation (not a translation of the reference's per-cell object code):
//  * a block of T threads (32, 128 or 256 by the widest diagonal of the job) walks the x+y diagonals; thread t owns the
//    cells t, t+T, ... of a diagonal (cell index k = (xmy - xmyL) / 2);
//  * the two previous diagonals every cell needs live in a TWO-slot ring in shared memory, one slot per diagonal parity,
//    indexed by an absolute coordinate a = (xmy + parity) / 2 (mod the ring width): the cell at the same xmy two diagonals
//    back -- the "middle" neighbour -- then sits at the very position the new cell is written to, so a diagonal is updated
//    IN PLACE over the one two steps back (each thread only overwrites what it alone has read) and the other slot is
//    read-only during the step; one barrier per diagonal. The same ring holds the backward diagonals during a traceback
//    -- forward and backward are never live together, the two forward diagonals the sweep resumes from are re-loaded;
//  * the ring is SPLIT: positions below RWs are in shared memory, the rest in an HBM/L2 overflow block of the same job.
//    A per-job shift of the absolute coordinate centres the band's cells of mass on the shared part, so that the
//    narrow diagonals (most of them) never leave shared memory and only the flanks of the widest diagonals spill.
//    Every job of a launch therefore runs with the same block shape and the same shared-memory footprint, whatever
//    its widest diagonal, and one work queue balances them;
//  * of the forward matrix only what a traceback reads goes to HBM: the MATCH plane of every cell (the posterior needs
//    f_M only) plus all five states of the few diagonals the total probability is recomputed on (every 10th, and its
//    predecessor) and of the two diagonals a sweep resumes from. Which diagonals those are depends only on the band
//    geometry, so the host marks them (pecan_plan.cpp, foff[]). That is 8 + ~0.2 * 40 bytes per cell instead of the 40
//    the reference keeps, both in power-of-two rings that only span the diagonals between two tracebacks;
//  * the backward pass is a GATHER: B[t] is computed from the final B[t+1], B[t+2] (the reference scatters from t+2 and
//    t+1 into t; the accumulation order into each target state is reproduced exactly, see bwd_diag);
//  * missing neighbours (outside the band / before the first diagonal) are read as LOG_ZERO cells, which is exact
//    because logAdd(x, LOG_ZERO) == x bit for bit, and the first transition into a state is an assignment because
//    logAdd(LOG_ZERO, v) == v;
//  * the total probability is a SERIAL logAdd chain over the cells of a diagonal in the reference (and its value depends
//    on that order); one warp walks it 32 cells at a time and only executes the steps that can change the running
//    total (a cell more than 7.5 nats below it returns the total unchanged by definition of logAdd) -- exact, and
//    typically a handful of steps instead of hundreds;
//  * all arithmetic is IEEE double without contraction (__dadd_rn / __dmul_rn on the device) in the reference's
//    operation order, so forward, backward and total probabilities are BIT-IDENTICAL to the CPU's; the kernel emits
//    the log posterior (f_M + b_M - total) of every candidate pair and the host applies exp / threshold / floor
//    with the same libm the reference uses (pecan.cu), which makes the integer triples identical as well.
//
// The same source compiles for the host (tests/hosttest) where a block is emulated by running the T threads of every
// phase one after the other: threads only communicate through memory between phases, or through the PC_* helpers.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PC_HD __host__ __device__ __forceinline__
#else
#define PC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PC_THREADS(tid, T) for (int tid = (int)threadIdx.x, _pc_once = 1; _pc_once; _pc_once = 0)
#define PC_SYNC() __syncthreads()
#else
#define PC_THREADS(tid, T) for (int tid = 0; tid < (T); ++tid)
#define PC_SYNC() ((void)0)
#endif

namespace barb200 {
namespace pecan {

enum { S_M = 0, S_SX = 1, S_SY = 2, S_LX = 3, S_LY = 4, NSTATE = 5 };

// Constant table (shared memory on the device), filled by fill_constants() on the host in double arithmetic:
//   [0..15]   lookup() cubic coefficients of the four intervals, highest power first
//   [16..23]  gap emission + transition: [16 + 4*isN + j], j = SHORT_OPEN, SHORT_EXTEND, LONG_OPEN, LONG_EXTEND
//   [24..35]  match emission + transition: [24 + 3*cls + j], cls = match, transition, transversion, N; j = CONTINUE, FROM_SHORT, FROM_LONG
//   [36..55]  state vectors: start[5], raggedStart[5], end[5], raggedEnd[5]
enum { K_LOOKUP = 0, K_GAP = 16, K_MATCH = 24, K_START = 36, K_RSTART = 41, K_END = 46, K_REND = 51, K_TOTAL = 56 };

struct Consts { double v[K_TOTAL]; };

// stateMachine.c:395-448 (transitions), :269-292, 351-366 (emissions); pairwiseAligner.c:300-311 (lookup coefficients are
// float literals promoted to double)
inline void fill_constants(Consts &c) {
    const float lk[16] = {-0.009350833524763f, 0.130659527668286f, 0.498799810682272f, 0.693203116424741f,
                          -0.014532321752540f, 0.139942324101744f, 0.495635523139337f, 0.692140569840976f,
                          -0.004605031767994f, 0.063427417320019f, 0.695956496475118f, 0.514272634594009f,
                          -0.000458661602210f, 0.009695946122598f, 0.930734667215156f, 0.168037164329057f};
    for (int i = 0; i < 16; ++i) c.v[K_LOOKUP + i] = (double)lk[i];
    const double T_MATCH_CONTINUE = -0.030064059121770816, T_MATCH_FROM_SHORT = -1.272871422049609,
                 T_MATCH_FROM_LONG = -5.673280173170473, T_SHORT_OPEN = -4.34381910900448,
                 T_SHORT_EXTEND = -0.3388262689231553, T_LONG_OPEN = -6.30810595366929, T_LONG_EXTEND = -0.003442492794189331;
    const double E_MATCH = -2.1149196655034745, E_TRANSVERSION = -4.5691014376830479, E_TRANSITION = -3.9833860032220842,
                 E_GAP = -1.6094379124341003, E_GAP_N = -1.386294361, E_MATCH_N = -2.772588722;
    const volatile double eg[2] = {E_GAP, E_GAP_N};
    const volatile double tg[4] = {T_SHORT_OPEN, T_SHORT_EXTEND, T_LONG_OPEN, T_LONG_EXTEND};
    for (int n = 0; n < 2; ++n) for (int j = 0; j < 4; ++j) c.v[K_GAP + 4 * n + j] = eg[n] + tg[j];           // eP + tP, pairwiseAligner.c:394
    const volatile double em[4] = {E_MATCH, E_TRANSITION, E_TRANSVERSION, E_MATCH_N};
    const volatile double tm[3] = {T_MATCH_CONTINUE, T_MATCH_FROM_SHORT, T_MATCH_FROM_LONG};
    for (int n = 0; n < 4; ++n) for (int j = 0; j < 3; ++j) c.v[K_MATCH + 3 * n + j] = em[n] + tm[j];
    const double LZ = -INFINITY;
    const double st[5] = {0, LZ, LZ, LZ, LZ}, rst[5] = {LZ, LZ, LZ, 0, 0};                                       // stateMachine.c:395-448
    const double en[5] = {T_MATCH_CONTINUE, T_MATCH_FROM_SHORT, T_MATCH_FROM_SHORT, T_MATCH_FROM_LONG, T_MATCH_FROM_LONG};
    const double ren[5] = {T_LONG_OPEN, T_LONG_OPEN, T_LONG_OPEN, T_LONG_EXTEND, T_LONG_EXTEND};
    for (int s = 0; s < 5; ++s) { c.v[K_START + s] = st[s]; c.v[K_RSTART + s] = rst[s]; c.v[K_END + s] = en[s]; c.v[K_REND + s] = ren[s]; }
}

struct Params {
    double log_thr_lo;      // emit candidates with log posterior >= this (slightly below log(threshold); exact test on the host)
    int min_diags;          // minDiagsBetweenTraceBack
    int tb_diags;           // traceBackDiagonals
    int expansion;          // diagonalExpansion
};

// One job = one getPosteriorProbsWithBanding call (one split sub-matrix of one sequence pair).
struct Job {
    long long sx_off, sy_off;   // symbols 0..4 of X / Y in the packed symbol buffer
    long long band_off;         // first of D+2 entries in bandL / coff / foff
    long long out_off;          // first output record of the job
    int lx, ly;
    int ragged;                 // bit 0: ragged left end, bit 1: ragged right end
    int out_cap;                // output records available
    int ring_shift;             // subtracted from the absolute ring coordinate (centres the band on the shared part)
    int pad_;
};

struct Pair { int x, y; double lp; };   // 0-based sequence coordinates, log posterior

struct DiagMeta { int L, co, fo, pad; };   // per diagonal: xmyL, cells before it, cells of MARKED diagonals before it (D+2 entries;
                                            // diagonal d is marked iff fo[d+1] > fo[d])

struct CtaMem {
    double *ring;         // shared part of the ring: 2 parity slots x RWs positions x 5 states (doubles)
    double *ring_o;       // overflow part in HBM / L2: 2 x (RW - RWs) x 5 doubles
    double *tbuf;         // RWs doubles: per-cell terms of a reduction (cells >= RWs: tbuf_o)
    double *tbuf_o;
    double *total;        // 1 double: the running total probability, published by warp 0
    int *n_out;           // candidates appended so far
    int RW;               // ring width (the modulus): >= widest diagonal of every job of the launch
    int RWs;              // positions [0, RWs) live in shared memory
    double *FM;           // HBM ring of forward MATCH values, maskM + 1 doubles
    unsigned maskM;
    double *FF;           // HBM ring of complete forward cells (marked diagonals only), maskF + 1 doubles
    unsigned maskF;
    int T;                // threads in the block
};

PC_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PC_HD double d_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
PC_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PC_HD unsigned long long d_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b; memcpy(&b, &v, 8); return b;
#endif
}
PC_HD double bits_d(unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v; memcpy(&v, &b, 8); return v;
#endif
}
PC_HD double log_zero() { return bits_d(0xfff0000000000000ULL); }

// pairwiseAligner.c:313-317 without branches and with one FP64 operation for the ordering: d = x - y gives the order by its
// sign bit and the (exactly antisymmetric) difference big - small as |d|. The interval of lookup() and the "return the
// larger one" condition come from the HIGH WORD of |d| with 32-bit integer arithmetic, which keeps them off the FP64 pipe:
// the thresholds 1.0, 2.5, 4.5, 7.5 have zero low words, so |d| > T  <=>  hi + (lo != 0) > hi(T)  and  |d| >= T  <=>
// hi >= hi(T). |d| is +inf when small is LOG_ZERO and NaN (either sign, the sign bit is cleared) when both are: both have
// a high word above hi(7.5), so the larger operand is returned exactly as the reference does.
PC_HD double log_add(double x, double y, const double *K) {
    const unsigned long long db_signed = d_bits(d_sub(x, y));
    const bool lt = (db_signed >> 63) != 0;                           // x < y (or a NaN / -0 difference, where big == small in value)
    const double big = lt ? y : x, small = lt ? x : y;
    const unsigned long long db = db_signed & 0x7fffffffffffffffULL;
    const double diff = bits_d(db);
    const int hi = (int)(unsigned)(db >> 32);                          // 0 .. 0x7fffffff
    const int h2 = hi + (int)((unsigned)db != 0u);
    // (T - h2) >>> 31 is 1 exactly when h2 > T (no overflow: both are below 2^31)
    const int idx = (int)(((unsigned)(0x3FF00000 - h2) >> 31) + ((unsigned)(0x40040000 - h2) >> 31) + ((unsigned)(0x40120000 - h2) >> 31));   // > 1.0, 2.5, 4.5
    const bool keep_big = hi >= 0x401E0000;                            // diff >= 7.5, small == LOG_ZERO, or both LOG_ZERO
    const double *c = K + K_LOOKUP + 4 * idx;
    double r = d_add(d_mul(c[0], diff), c[1]);
    r = d_add(d_mul(r, diff), c[2]);
    r = d_add(d_mul(r, diff), c[3]);
    r = d_add(r, small);
    return keep_big ? big : r;
}

// would logAdd(tot, t) return tot unchanged?  (pairwiseAligner.c:313-317: the else branch returning x)
PC_HD bool chain_inactive(double tot, double t) { return !(tot < t) && (t == log_zero() || d_sub(tot, t) >= 7.5); }

PC_HD int match_class(int cx, int cy) { return (cx == 4 || cy == 4) ? 3 : (cx == cy ? 0 : (((cx ^ cy) == 2) ? 1 : 2)); }

// The ring holds complete cells: the five states of ring position i of the slot of parity par are consecutive doubles
// (a thread fetches a neighbour with one address and immediate offsets; 40-byte strides are conflict-free for 64-bit
// shared-memory accesses). Positions [0, RWs) are in shared memory, the rest in the job's overflow block.
PC_HD double &ring_at(const CtaMem &cm, int par, int s, int i) {
    return i < cm.RWs ? cm.ring[(par * cm.RWs + i) * 5 + s] : cm.ring_o[((size_t)par * (size_t)(cm.RW - cm.RWs) + (size_t)(i - cm.RWs)) * 5 + (size_t)s];
}
// hot-path forms: one shared / overflow decision per cell (two plain branches instead of a generic pointer per state)
PC_HD void ring_load3(const CtaMem &cm, int par, int i, int s0, int s1, int s2, double &a, double &b, double &c) {
    if (i < cm.RWs) { const double *q = cm.ring + (par * cm.RWs + i) * 5; a = q[s0]; b = q[s1]; c = q[s2]; }
    else { const double *q = cm.ring_o + ((size_t)par * (size_t)(cm.RW - cm.RWs) + (size_t)(i - cm.RWs)) * 5; a = q[s0]; b = q[s1]; c = q[s2]; }
}
PC_HD void ring_load2(const CtaMem &cm, int par, int i, int s0, int s1, double &a, double &b) {
    if (i < cm.RWs) { const double *q = cm.ring + (par * cm.RWs + i) * 5; a = q[s0]; b = q[s1]; }
    else { const double *q = cm.ring_o + ((size_t)par * (size_t)(cm.RW - cm.RWs) + (size_t)(i - cm.RWs)) * 5; a = q[s0]; b = q[s1]; }
}
PC_HD double &tbuf_at(const CtaMem &cm, int k) { return k < cm.RWs ? cm.tbuf[k] : cm.tbuf_o[k - cm.RWs]; }
// ring index of cell 0 of a diagonal: a = (xmyL + parity) / 2 + ly - shift, modulo the ring width
PC_HD int ring_i0(int L, int d, const Job &J, int RW) { const int a = (((L + (d & 1)) >> 1) + J.ly - J.ring_shift) % RW; return a < 0 ? a + RW : a; }
PC_HD int wrap(int i, int RW) { return i >= RW ? i - RW : (i < 0 ? i + RW : i); }
PC_HD double &fm_at(const CtaMem &cm, int cell) { return cm.FM[(unsigned)cell & cm.maskM]; }
PC_HD double &ff_at(const CtaMem &cm, int base, int w, int s, int k) { return cm.FF[(unsigned)(base + s * w + k) & cm.maskF]; }

PC_HD void emit_one(const Job &J, const CtaMem &cm, Pair *out, int x, int y, double lp) {
#if defined(__CUDA_ARCH__)
    const int pos = atomicAdd(cm.n_out, 1);
#else
    const int pos = (*cm.n_out)++;
#endif
    if (pos < J.out_cap) { Pair p; p.x = x - 1; p.y = y - 1; p.lp = lp; out[pos] = p; }
}

struct Cell5 { double m, sx, sy, lx, ly; };

// ---- forward: diagonal d from d-1 (lower: xmy-1, upper: xmy+1) and d-2 (middle: xmy, updated in place) -------------------
struct FwdDiag {            // per-diagonal uniforms
    const uint8_t *sx, *sy; const double *K;
    int d, p, Ld, w, w1, w2, sl, sm, i0, RW, x0, y0;
};
PC_HD Cell5 fwd_cell(const CtaMem &cm, const FwdDiag &f, int k) {
    const double LZ = log_zero();
    const int RW = f.RW, p = f.p, q = p ^ 1;
    const int x = f.x0 + k, y = f.y0 - k;                                // x = (d + xmy) / 2, y = (d - xmy) / 2 with xmy = Ld + 2k
    const int cx = x > 0 ? f.sx[x - 1] : 4, cy = y > 0 ? f.sy[y - 1] : 4;
    const int kl = k + f.sl, km = k + f.sm;
    const int i = wrap(f.i0 + k, RW), il = wrap(i - p, RW), iu = wrap(i + 1 - p, RW);
    double lM = LZ, lSX = LZ, lLX = LZ, uM = LZ, uSY = LZ, uLY = LZ, mM = LZ, mSX = LZ, mSY = LZ, mLX = LZ, mLY = LZ;
    if ((unsigned)kl < (unsigned)f.w1) ring_load3(cm, q, il, S_M, S_SX, S_LX, lM, lSX, lLX);
    if ((unsigned)(kl + 1) < (unsigned)f.w1) ring_load3(cm, q, iu, S_M, S_SY, S_LY, uM, uSY, uLY);
    if ((unsigned)km < (unsigned)f.w2) { ring_load3(cm, p, i, S_M, S_SX, S_SY, mM, mSX, mSY); ring_load2(cm, p, i, S_LX, S_LY, mLX, mLY); }
    const double *K = f.K;
    const double *gx = K + K_GAP + 4 * (cx == 4), *gy = K + K_GAP + 4 * (cy == 4), *mt = K + K_MATCH + 3 * match_class(cx, cy);
    // stateMachine.c:450-480 in its order of transitions; the first transition into a state is an assignment
    Cell5 v;
    v.sx = d_add(lM, gx[0]); v.sx = log_add(v.sx, d_add(lSX, gx[1]), K);
    v.lx = d_add(lM, gx[2]); v.lx = log_add(v.lx, d_add(lLX, gx[3]), K);
    v.m = d_add(mM, mt[0]);
    v.m = log_add(v.m, d_add(mSX, mt[1]), K); v.m = log_add(v.m, d_add(mSY, mt[1]), K);
    v.m = log_add(v.m, d_add(mLX, mt[2]), K); v.m = log_add(v.m, d_add(mLY, mt[2]), K);
    v.sy = d_add(uM, gy[0]); v.sy = log_add(v.sy, d_add(uSY, gy[1]), K);
    v.ly = d_add(uM, gy[2]); v.ly = log_add(v.ly, d_add(uLY, gy[3]), K);
    return v;
}
PC_HD void ring_store(const CtaMem &cm, int par, int i, const Cell5 &v) {
    double *q = i < cm.RWs ? nullptr : cm.ring_o + ((size_t)par * (size_t)(cm.RW - cm.RWs) + (size_t)(i - cm.RWs)) * 5;
    if (i < cm.RWs) { double *r = cm.ring + (par * cm.RWs + i) * 5; r[S_M] = v.m; r[S_SX] = v.sx; r[S_SY] = v.sy; r[S_LX] = v.lx; r[S_LY] = v.ly; }
    else { q[S_M] = v.m; q[S_SX] = v.sx; q[S_SY] = v.sy; q[S_LX] = v.lx; q[S_LY] = v.ly; }
}
// md = meta of d, mn = of d+1, m1 = of d-1, m2 = of d-2. (Keeping two cells per thread in flight to interleave their
// dependent logAdd chains was measured SLOWER, 5.6 vs 8.6 Gcell/s: the duplicated work on narrow diagonals and the
// doubled register footprint cost more than the latency it hides.)
PC_HD void fwd_diag(const Job &J, const uint8_t *sx, const uint8_t *sy, const CtaMem &cm, const double *K, int d,
                    const DiagMeta &md, const DiagMeta &mn, const DiagMeta &m1, const DiagMeta &m2) {
    FwdDiag f;
    f.sx = sx; f.sy = sy; f.K = K; f.RW = cm.RW; f.d = d; f.p = d & 1;
    f.Ld = md.L; f.w = mn.co - md.co; f.w1 = md.co - m1.co; f.w2 = d >= 2 ? m1.co - m2.co : 0;
    f.sl = (f.Ld - m1.L - 1) >> 1; f.sm = (f.Ld - m2.L) >> 1;      // both differences are even
    f.i0 = ring_i0(f.Ld, d, J, cm.RW);
    f.x0 = (d + f.Ld) >> 1; f.y0 = (d - f.Ld) >> 1;
    const int w = f.w, cbase = md.co, fbase = 5 * md.fo, RW = cm.RW;
    const bool full = mn.fo > md.fo;
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const Cell5 a = fwd_cell(cm, f, k);
            ring_store(cm, f.p, wrap(f.i0 + k, RW), a);                  // in place over the diagonal two back
            fm_at(cm, cbase + k) = a.m;
            if (full) {
                ff_at(cm, fbase, w, S_M, k) = a.m; ff_at(cm, fbase, w, S_SX, k) = a.sx; ff_at(cm, fbase, w, S_SY, k) = a.sy;
                ff_at(cm, fbase, w, S_LX, k) = a.lx; ff_at(cm, fbase, w, S_LY, k) = a.ly;
            }
        }
    }
    PC_SYNC();
}

// ---- backward: B[t] gathered from B[t+1] (cells xmy-1 and xmy+1) and B[t+2] (cell xmy, updated in place); top = the
// diagonal walked from. Order of accumulation into the target cell c in the reference's scatter (pairwiseAligner.c:
// 619-634 walking xmy upwards, stateMachine.c:450-480): while diagonal t+2 is processed c is the MIDDLE of the cell at the
// same xmy (all five states receive from its match state); while t+1 is processed c is first the UPPER of the cell at
// xmy-1 (M += SY, SY += SY, M += LY, LY += LY) and then the LOWER of the cell at xmy+1 (M += SX, SX += SX, M += LX, LX += LX).
// With fuse_emit the posterior candidates of the diagonal (diagonalCalculationPosteriorMatchProbs, :676-699) are produced
// in the same pass from the total probability already published.
struct BwdDiag {            // per-diagonal uniforms
    const uint8_t *sx, *sy; const double *K;
    int t, p, Lt, w1, w2, s1, s2, i0, RW, x0, y0;
};
PC_HD Cell5 bwd_cell(const CtaMem &cm, const BwdDiag &g, int k) {
    const double LZ = log_zero();
    const int RW = g.RW, p = g.p, q = p ^ 1;
    const int x = g.x0 + k, y = g.y0 - k;
    const int ku = k + g.s1, km = k + g.s2;                    // ku: cell (t+1, xmy-1) whose upper is c; ku+1: cell (t+1, xmy+1) whose lower is c
    const int i = wrap(g.i0 + k, RW), iu = wrap(i - p, RW), il = wrap(i + 1 - p, RW);
    double mid = LZ, upSY = LZ, upLY = LZ, loSX = LZ, loLX = LZ;
    const double *K = g.K;
    const double *gx = K + K_GAP, *gy = K + K_GAP, *mt = K + K_MATCH;
    if ((unsigned)km < (unsigned)g.w2) { mid = ring_at(cm, p, S_M, i); mt = K + K_MATCH + 3 * match_class(g.sx[x], g.sy[y]); }                 // cell (x+1, y+1)
    if ((unsigned)ku < (unsigned)g.w1) { ring_load2(cm, q, iu, S_SY, S_LY, upSY, upLY); gy = K + K_GAP + 4 * (g.sy[y] == 4); }               // cell (x, y+1)
    if ((unsigned)(ku + 1) < (unsigned)g.w1) { ring_load2(cm, q, il, S_SX, S_LX, loSX, loLX); gx = K + K_GAP + 4 * (g.sx[x] == 4); }         // cell (x+1, y)
    Cell5 v;
    v.m = d_add(mid, mt[0]);
    v.m = log_add(v.m, d_add(upSY, gy[0]), K); v.m = log_add(v.m, d_add(upLY, gy[2]), K);
    v.m = log_add(v.m, d_add(loSX, gx[0]), K); v.m = log_add(v.m, d_add(loLX, gx[2]), K);
    v.sx = log_add(d_add(mid, mt[1]), d_add(loSX, gx[1]), K);
    v.sy = log_add(d_add(mid, mt[1]), d_add(upSY, gy[1]), K);
    v.lx = log_add(d_add(mid, mt[2]), d_add(loLX, gx[3]), K);
    v.ly = log_add(d_add(mid, mt[2]), d_add(upLY, gy[3]), K);
    return v;
}
// mt = meta of t, mt1 = of t+1, mt2 = of t+2, mt3 = of t+3
PC_HD void bwd_diag(const Job &J, const uint8_t *sx, const uint8_t *sy, const CtaMem &cm, const Params &P, const double *K, int t, int top,
                    const DiagMeta &mt_, const DiagMeta &mt1, const DiagMeta &mt2, const DiagMeta &mt3, bool fuse_emit, Pair *out) {
    BwdDiag g;
    g.sx = sx; g.sy = sy; g.K = K; g.RW = cm.RW; g.t = t; g.p = t & 1;
    g.Lt = mt_.L; g.w1 = mt2.co - mt1.co; g.w2 = (t + 2 <= top) ? mt3.co - mt2.co : 0;
    g.s1 = (g.Lt - 1 - mt1.L) >> 1; g.s2 = (g.Lt - mt2.L) >> 1;
    g.i0 = ring_i0(g.Lt, t, J, cm.RW);
    g.x0 = (t + g.Lt) >> 1; g.y0 = (t - g.Lt) >> 1;
    const int w = mt1.co - mt_.co, cbase = mt_.co, RW = cm.RW, Lt = g.Lt;
    const double total = fuse_emit ? *cm.total : 0.0;
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const int xa = (t + Lt + 2 * k) >> 1, ya = (t - Lt - 2 * k) >> 1;
            const bool cand = fuse_emit && xa > 0 && ya > 0;
            const double fm = cand ? fm_at(cm, cbase + k) : 0.0;   // issued early: the only HBM read of the step
            const Cell5 a = bwd_cell(cm, g, k);
            ring_store(cm, g.p, wrap(g.i0 + k, RW), a);                  // in place over B[t+2]
            if (cand) { const double lp = d_sub(d_add(fm, a.m), total); if (lp >= P.log_thr_lo) emit_one(J, cm, out, xa, ya, lp); }
        }
    }
    PC_SYNC();
}

// Serial logAdd chain over tbuf[0..w) in cell order (dpDiagonal_dotProduct, pairwiseAligner.c:523-534), run by warp 0,
// 32 cells per probe: cells that cannot change the running total are skipped together (exact, see the file header).
// The result is published in *cm.total; callers synchronise before reading it.
PC_HD void chain_warp0(const CtaMem &cm, int w, const double *K, bool accumulate_into_total) {
#if defined(__CUDA_ARCH__)
    if (threadIdx.x < 32) {
        const int lane = (int)threadIdx.x;
        double tot = log_zero();
        for (int k0 = 0; k0 < w; k0 += 32) {
            const double t = (k0 + lane < w) ? tbuf_at(cm, k0 + lane) : log_zero();
            unsigned rem = 0xffffffffu;
            for (;;) {
                const unsigned m = __ballot_sync(0xffffffffu, !chain_inactive(tot, t)) & rem;
                if (!m) break;
                const int j = __ffs((int)m) - 1;
                const double tj = __shfl_sync(0xffffffffu, t, j);
                tot = log_add(tot, tj, K);
                rem = (j == 31) ? 0u : ~((2u << j) - 1u);
            }
        }
        if (lane == 0) *cm.total = accumulate_into_total ? log_add(*cm.total, tot, K) : tot;
    }
#else
    double tot = log_zero();
    for (int k0 = 0; k0 < w; k0 += 32) {            // the same probe / skip structure, lane by lane
        int from = 0;
        for (;;) {
            int j = -1;
            for (int lane = from; lane < 32 && j < 0; ++lane) {
                const double t = (k0 + lane < w) ? tbuf_at(cm, k0 + lane) : log_zero();
                if (!chain_inactive(tot, t)) j = lane;
            }
            if (j < 0) break;
            tot = log_add(tot, tbuf_at(cm, k0 + j), K);
            from = j + 1;
        }
    }
    *cm.total = accumulate_into_total ? log_add(*cm.total, tot, K) : tot;
#endif
}

// diagonalCalculationTotalProbability, pairwiseAligner.c:646-663. Needs the complete forward cells of diagonals t and
// t-1 (marked by the host). Result in *cm.total (after the final synchronisation).
PC_HD void total_probability(const Job &J, const uint8_t *sx, const uint8_t *sy, const DiagMeta *M, const CtaMem &cm, const double *K, int t, int top) {
    const double LZ = log_zero();
    const int RW = cm.RW;
    const DiagMeta mt_ = M[t], mt1 = M[t + 1];
    {
        const int fbase = 5 * mt_.fo, w = mt1.co - mt_.co;
        const int i0 = ring_i0(mt_.L, t, J, RW), pt = t & 1;
        PC_THREADS(tid, cm.T) {
            for (int k = tid; k < w; k += cm.T) {
                const int i = wrap(i0 + k, RW);
                double tt = d_add(ff_at(cm, fbase, w, 0, k), ring_at(cm, pt, 0, i));       // cell_dotProduct, pairwiseAligner.c:412-418
                for (int s = 1; s < NSTATE; ++s) tt = log_add(tt, d_add(ff_at(cm, fbase, w, s, k), ring_at(cm, pt, s, i)), K);
                tbuf_at(cm, k) = tt;
            }
        }
        PC_SYNC();
        chain_warp0(cm, w, K, false);
        PC_SYNC();
    }
    if (t + 1 <= top) {                                     // matches through t: forward t-1 -> match -> backward t+1
        const DiagMeta mf = M[t - 1], mt2 = M[t + 2];
        const int Lq = mt1.L, wq = mt2.co - mt1.co;
        const int fbase = 5 * mf.fo, wf = mt_.co - mf.co;
        const int sm = (Lq - mf.L) >> 1;
        const int i0 = ring_i0(Lq, t + 1, J, RW), pq = (t + 1) & 1;
        PC_THREADS(tid, cm.T) {
            for (int k = tid; k < wq; k += cm.T) {
                const int xmy = Lq + 2 * k, x = (t + 1 + xmy) >> 1, y = (t + 1 - xmy) >> 1, km = k + sm;
                const int cx = x > 0 ? sx[x - 1] : 4, cy = y > 0 ? sy[y - 1] : 4;
                double mM = LZ, mSX = LZ, mSY = LZ, mLX = LZ, mLY = LZ;
                if (km >= 0 && km < wf) {
                    mM = ff_at(cm, fbase, wf, S_M, km); mSX = ff_at(cm, fbase, wf, S_SX, km); mSY = ff_at(cm, fbase, wf, S_SY, km);
                    mLX = ff_at(cm, fbase, wf, S_LX, km); mLY = ff_at(cm, fbase, wf, S_LY, km);
                }
                const double *mt = K + K_MATCH + 3 * match_class(cx, cy);
                double vM = d_add(mM, mt[0]);
                vM = log_add(vM, d_add(mSX, mt[1]), K); vM = log_add(vM, d_add(mSY, mt[1]), K);
                vM = log_add(vM, d_add(mLX, mt[2]), K); vM = log_add(vM, d_add(mLY, mt[2]), K);
                tbuf_at(cm, k) = d_add(vM, ring_at(cm, pq, S_M, wrap(i0 + k, RW)));   // the other four states of the match-only diagonal are LOG_ZERO
            }
        }
        PC_SYNC();
        chain_warp0(cm, wq, K, true);                        // total = logAdd(total, second dot product), :659
        PC_SYNC();
    }
}

// diagonalCalculationPosteriorMatchProbs, pairwiseAligner.c:676-699, as a pass of its own (diagonals the total probability
// was just recomputed on)
PC_HD void emit_diag(const Job &J, const DiagMeta *M, const CtaMem &cm, const Params &P, int t, Pair *out) {
    const DiagMeta mt_ = M[t], mt1 = M[t + 1];
    const int Lt = mt_.L, cbase = mt_.co, w = mt1.co - mt_.co;
    const double total = *cm.total;
    const int i0 = ring_i0(Lt, t, J, cm.RW), pt = t & 1;
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const int xmy = Lt + 2 * k, x = (t + xmy) >> 1, y = (t - xmy) >> 1;
            if (x > 0 && y > 0) {
                const double lp = d_sub(d_add(fm_at(cm, cbase + k), ring_at(cm, pt, S_M, wrap(i0 + k, cm.RW))), total);
                if (lp >= P.log_thr_lo) emit_one(J, cm, out, x, y, lp);
            }
        }
    }
}

// getPosteriorProbsWithBanding, pairwiseAligner.c:766-887. Returns the number of candidate pairs (may exceed out_cap:
// then only out_cap were stored and the job must be re-run with more room).
PC_HD int run_job(const Job &J, const uint8_t *sym, const DiagMeta *meta, const CtaMem &cm, const Params &P, const double *K, Pair *out_all) {
    const int D = J.lx + J.ly;
    if (D == 0) return 0;
    const uint8_t *sx = sym + J.sx_off, *sy = sym + J.sy_off;
    const DiagMeta *M = meta + J.band_off;
    Pair *out = out_all + J.out_off;
    const int RW = cm.RW;
    DiagMeta m2 = M[0], m1 = M[0], md = M[1], mn = M[2];
    {   // diagonal 0: the single cell (0, 0) holds the start state vector (dpDiagonal_initialiseValues, :785-786)
        const double *st = K + ((J.ragged & 1) ? K_RSTART : K_START);
        const int w0 = md.co - m1.co;
        const bool full = md.fo > m1.fo;
        const int i0 = ring_i0(m1.L, 0, J, RW);
        PC_THREADS(tid, cm.T) {
            if (tid == 0) *cm.n_out = 0;
            for (int k = tid; k < w0; k += cm.T) {
                for (int s = 0; s < NSTATE; ++s) { ring_at(cm, 0, s, wrap(i0 + k, RW)) = st[s]; if (full) ff_at(cm, 5 * m1.fo, w0, s, k) = st[s]; }
                fm_at(cm, m1.co + k) = st[S_M];
            }
        }
        PC_SYNC();
    }
    int tb_to = 0;
    for (int d = 1; d <= D; ++d) {
        const DiagMeta mnn = d + 2 <= D + 1 ? M[d + 2] : mn;          // fetched one step ahead
        fwd_diag(J, sx, sy, cm, K, d, md, mn, m1, m2);
        const int w = mn.co - md.co;
        const bool at_end = d == D;
        const bool tb_point = d >= tb_to + P.min_diags && w <= P.expansion * 2 + 1;
        if (at_end || tb_point) {
            {   // the diagonal walked back from holds the end state vector (:806-808)
                const double *en = K + ((at_end && (J.ragged & 2)) ? K_REND : K_END);
                const int i0 = ring_i0(md.L, d, J, RW);
                PC_THREADS(tid, cm.T) { for (int k = tid; k < w; k += cm.T) for (int s = 0; s < NSTATE; ++s) ring_at(cm, d & 1, s, wrap(i0 + k, RW)) = en[s]; }
                PC_SYNC();
            }
            const int tb_from = d - (at_end ? 0 : P.tb_diags + 1);
            int ncalc = 0;
            DiagMeta b0 = md, b1 = mn, b2 = mn, b3 = mn;               // metas of t, t+1, t+2, t+3
            for (int t = d; t > tb_to; --t) {
                const DiagMeta bp = M[t - 1];                           // fetched one step ahead
                const bool post = t <= tb_from;
                const bool recompute = post && (ncalc % 10 == 0);
                if (post) ++ncalc;
                if (t < d) bwd_diag(J, sx, sy, cm, P, K, t, d, b0, b1, b2, b3, post && !recompute, out);
                if (recompute) {
                    total_probability(J, sx, sy, M, cm, K, t, d);
                    emit_diag(J, M, cm, P, t, out);
                }
                b3 = b2; b2 = b1; b1 = b0; b0 = bp;
            }
            tb_to = tb_from;
            if (!at_end) {          // the sweep resumes from the complete forward diagonals d and d-1 (marked by the host)
                PC_SYNC();
                for (int q = 0; q < 2; ++q) {
                    const int dd = d - q;
                    const DiagMeta a = q ? m1 : md, b = q ? md : mn;
                    const int wd = b.co - a.co, fbase = 5 * a.fo, i0 = ring_i0(a.L, dd, J, RW);
                    PC_THREADS(tid, cm.T) { for (int k = tid; k < wd; k += cm.T) for (int s = 0; s < NSTATE; ++s) ring_at(cm, dd & 1, s, wrap(i0 + k, RW)) = ff_at(cm, fbase, wd, s, k); }
                }
                PC_SYNC();
            }
        }
        m2 = m1; m1 = md; md = mn; mn = mnn;
    }
    PC_SYNC();
    return *cm.n_out;
}

}  // namespace pecan
}  // namespace barb200
