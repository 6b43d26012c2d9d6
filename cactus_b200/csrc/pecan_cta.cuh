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
//  * the two previous diagonals every cell needs live in a three-slot RING in shared memory (state-major, so a warp reads
//    consecutive doubles); the same ring holds the backward diagonals during a traceback -- forward and backward are never
//    live together, the two forward diagonals the sweep resumes from are re-loaded afterwards;
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
};

struct Pair { int x, y; double lp; };   // 0-based sequence coordinates, log posterior

struct CtaMem {
    double *ring;         // 3 diagonals x 5 states x RW doubles (shared memory; global scratch for very wide jobs)
    double *tbuf;         // RW doubles: per-cell terms of a reduction / candidate log posteriors
    double *total;        // 1 double: the running total probability, published by warp 0
    int RW;               // >= widest diagonal of the job
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
PC_HD double log_zero() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)0xfff0000000000000ULL);
#else
    return -INFINITY;
#endif
}
// "not a candidate" marker in the candidate buffer (a computed NaN never gets there: it fails the threshold test)
PC_HD double not_a_candidate() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}
PC_HD bool is_candidate(double lp) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(lp) != 0x7ff8000000000000LL;
#else
    return lp == lp;
#endif
}

// pairwiseAligner.c:313-317 without branches: big/small by one comparison; the interval of lookup() and the two "return
// the larger one" conditions are taken from the bit pattern of the (non-negative) difference with integer compares, which
// keeps them off the FP64 pipe
PC_HD double log_add(double x, double y, const double *K) {
    const bool lt = x < y;
    const double big = lt ? y : x, small = lt ? x : y;
    const double diff = d_sub(big, small);                    // >= +0, +inf (small == LOG_ZERO) or NaN (both LOG_ZERO)
#if defined(__CUDA_ARCH__)
    const long long db = __double_as_longlong(diff);
    const int idx = (int)(db > 0x3FF0000000000000LL) + (int)(db > 0x4004000000000000LL) + (int)(db > 0x4012000000000000LL);   // > 1.0, 2.5, 4.5
    const bool keep_big = (__double_as_longlong(small) == (long long)0xFFF0000000000000ULL) | (db >= 0x401E000000000000LL);   // small == LOG_ZERO || diff >= 7.5
#else
    const int idx = (int)(diff > 1.0) + (int)(diff > 2.5) + (int)(diff > 4.5);
    const bool keep_big = small == log_zero() || diff >= 7.5;
#endif
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

PC_HD double *slot(const CtaMem &cm, int d) { return cm.ring + (size_t)(d % 3) * (5 * (size_t)cm.RW); }
PC_HD double &fm_at(const CtaMem &cm, int cell) { return cm.FM[(unsigned)cell & cm.maskM]; }
PC_HD double &ff_at(const CtaMem &cm, int base, int w, int s, int k) { return cm.FF[(unsigned)(base + s * w + k) & cm.maskF]; }

struct Band {            // per-diagonal tables of one job
    const int *L;        // xmyL
    const int *co;       // cells before diagonal d (D+2 entries)
    const int *fo;       // cells of MARKED diagonals before diagonal d (D+2 entries); d is marked iff fo[d+1] > fo[d]
};

// ---- forward: diagonal d from d-1 (lower: xmy-1, upper: xmy+1) and d-2 (middle: xmy) -------------------------------------
PC_HD void fwd_diag(const uint8_t *sx, const uint8_t *sy, const Band &bd, const CtaMem &cm, const double *K, int d) {
    const double LZ = log_zero();
    const int Ld = bd.L[d], w = bd.co[d + 1] - bd.co[d], cbase = bd.co[d], RW = cm.RW;
    const int L1 = bd.L[d - 1], w1 = bd.co[d] - bd.co[d - 1];
    int L2 = 0, w2 = 0;
    if (d >= 2) { L2 = bd.L[d - 2]; w2 = bd.co[d - 1] - bd.co[d - 2]; }
    const int sl = (Ld - L1 - 1) >> 1, sm = (Ld - L2) >> 1;      // both differences are even
    const bool full = bd.fo[d + 1] > bd.fo[d];
    const int fbase = 5 * bd.fo[d];
    double *cur = slot(cm, d);
    const double *m1 = slot(cm, d - 1), *m2 = slot(cm, d + 1);   // (d - 2) % 3 == (d + 1) % 3
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const int xmy = Ld + 2 * k, x = (d + xmy) >> 1, y = (d - xmy) >> 1;
            const int cx = x > 0 ? sx[x - 1] : 4, cy = y > 0 ? sy[y - 1] : 4;
            const int kl = k + sl, ku = kl + 1, km = k + sm;
            double lM = LZ, lSX = LZ, lLX = LZ, uM = LZ, uSY = LZ, uLY = LZ, mM = LZ, mSX = LZ, mSY = LZ, mLX = LZ, mLY = LZ;
            if (kl >= 0 && kl < w1) { lM = m1[S_M * RW + kl]; lSX = m1[S_SX * RW + kl]; lLX = m1[S_LX * RW + kl]; }
            if (ku >= 0 && ku < w1) { uM = m1[S_M * RW + ku]; uSY = m1[S_SY * RW + ku]; uLY = m1[S_LY * RW + ku]; }
            if (km >= 0 && km < w2) {
                mM = m2[S_M * RW + km]; mSX = m2[S_SX * RW + km]; mSY = m2[S_SY * RW + km]; mLX = m2[S_LX * RW + km]; mLY = m2[S_LY * RW + km];
            }
            const double *gx = K + K_GAP + 4 * (cx == 4), *gy = K + K_GAP + 4 * (cy == 4), *mt = K + K_MATCH + 3 * match_class(cx, cy);
            // stateMachine.c:450-480 in its order of transitions; the first transition into a state is an assignment
            double vSX = d_add(lM, gx[0]); vSX = log_add(vSX, d_add(lSX, gx[1]), K);
            double vLX = d_add(lM, gx[2]); vLX = log_add(vLX, d_add(lLX, gx[3]), K);
            double vM = d_add(mM, mt[0]);
            vM = log_add(vM, d_add(mSX, mt[1]), K); vM = log_add(vM, d_add(mSY, mt[1]), K);
            vM = log_add(vM, d_add(mLX, mt[2]), K); vM = log_add(vM, d_add(mLY, mt[2]), K);
            double vSY = d_add(uM, gy[0]); vSY = log_add(vSY, d_add(uSY, gy[1]), K);
            double vLY = d_add(uM, gy[2]); vLY = log_add(vLY, d_add(uLY, gy[3]), K);
            cur[S_M * RW + k] = vM; cur[S_SX * RW + k] = vSX; cur[S_SY * RW + k] = vSY; cur[S_LX * RW + k] = vLX; cur[S_LY * RW + k] = vLY;
            fm_at(cm, cbase + k) = vM;
            if (full) {
                ff_at(cm, fbase, w, S_M, k) = vM; ff_at(cm, fbase, w, S_SX, k) = vSX; ff_at(cm, fbase, w, S_SY, k) = vSY;
                ff_at(cm, fbase, w, S_LX, k) = vLX; ff_at(cm, fbase, w, S_LY, k) = vLY;
            }
        }
    }
    PC_SYNC();
}

// ---- backward: B[t] gathered from B[t+1] (cells xmy-1 and xmy+1) and B[t+2] (cell xmy); top = diagonal walked from ----------
// Order of accumulation into the target cell c in the reference's scatter (pairwiseAligner.c:619-634 walking xmy upwards,
// stateMachine.c:450-480): while diagonal t+2 is processed c is the MIDDLE of the cell at the same xmy (all five states
// receive from its match state); while t+1 is processed c is first the UPPER of the cell at xmy-1 (M += SY, SY += SY,
// M += LY, LY += LY) and then the LOWER of the cell at xmy+1 (M += SX, SX += SX, M += LX, LX += LX).
PC_HD void bwd_diag(const uint8_t *sx, const uint8_t *sy, const Band &bd, const CtaMem &cm, const double *K, int t, int top) {
    const double LZ = log_zero();
    const int Lt = bd.L[t], w = bd.co[t + 1] - bd.co[t], RW = cm.RW;
    const int L1 = bd.L[t + 1], w1 = bd.co[t + 2] - bd.co[t + 1];
    const bool has2 = t + 2 <= top;
    int L2 = 0, w2 = 0;
    if (has2) { L2 = bd.L[t + 2]; w2 = bd.co[t + 3] - bd.co[t + 2]; }
    double *cur = slot(cm, t);
    const double *b1 = slot(cm, t + 1), *b2 = slot(cm, t + 2);
    const int s1 = (Lt - 1 - L1) >> 1, s2 = (Lt - L2) >> 1;
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const int xmy = Lt + 2 * k, x = (t + xmy) >> 1, y = (t - xmy) >> 1;
            const int ku = k + s1, kl = ku + 1, km = k + s2;    // ku: cell (t+1, xmy-1) whose upper is c; kl: cell (t+1, xmy+1) whose lower is c
            double mid = LZ, upSY = LZ, upLY = LZ, loSX = LZ, loLX = LZ;
            const double *gx = K + K_GAP, *gy = K + K_GAP, *mt = K + K_MATCH;
            if (has2 && km >= 0 && km < w2) { mid = b2[S_M * RW + km]; mt = K + K_MATCH + 3 * match_class(sx[x], sy[y]); }   // cell (x+1, y+1)
            if (ku >= 0 && ku < w1) { upSY = b1[S_SY * RW + ku]; upLY = b1[S_LY * RW + ku]; gy = K + K_GAP + 4 * (sy[y] == 4); }  // cell (x, y+1)
            if (kl >= 0 && kl < w1) { loSX = b1[S_SX * RW + kl]; loLX = b1[S_LX * RW + kl]; gx = K + K_GAP + 4 * (sx[x] == 4); }  // cell (x+1, y)
            double vM = d_add(mid, mt[0]);
            vM = log_add(vM, d_add(upSY, gy[0]), K); vM = log_add(vM, d_add(upLY, gy[2]), K);
            vM = log_add(vM, d_add(loSX, gx[0]), K); vM = log_add(vM, d_add(loLX, gx[2]), K);
            const double vSX = log_add(d_add(mid, mt[1]), d_add(loSX, gx[1]), K);
            const double vSY = log_add(d_add(mid, mt[1]), d_add(upSY, gy[1]), K);
            const double vLX = log_add(d_add(mid, mt[2]), d_add(loLX, gx[3]), K);
            const double vLY = log_add(d_add(mid, mt[2]), d_add(upLY, gy[3]), K);
            cur[S_M * RW + k] = vM; cur[S_SX * RW + k] = vSX; cur[S_SY * RW + k] = vSY; cur[S_LX * RW + k] = vLX; cur[S_LY * RW + k] = vLY;
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
            const double t = (k0 + lane < w) ? cm.tbuf[k0 + lane] : log_zero();
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
                const double t = (k0 + lane < w) ? cm.tbuf[k0 + lane] : log_zero();
                if (!chain_inactive(tot, t)) j = lane;
            }
            if (j < 0) break;
            tot = log_add(tot, cm.tbuf[k0 + j], K);
            from = j + 1;
        }
    }
    *cm.total = accumulate_into_total ? log_add(*cm.total, tot, K) : tot;
#endif
}

// diagonalCalculationTotalProbability, pairwiseAligner.c:646-663. Needs the complete forward cells of diagonals t and
// t-1 (marked by the host). Result in *cm.total (after the final synchronisation).
PC_HD void total_probability(const uint8_t *sx, const uint8_t *sy, const Band &bd, const CtaMem &cm, const double *K, int t, int top) {
    const double LZ = log_zero();
    const int RW = cm.RW;
    {
        const int fbase = 5 * bd.fo[t], w = bd.co[t + 1] - bd.co[t];
        const double *bt = slot(cm, t);
        PC_THREADS(tid, cm.T) {
            for (int k = tid; k < w; k += cm.T) {
                double tt = d_add(ff_at(cm, fbase, w, 0, k), bt[k]);                       // cell_dotProduct, pairwiseAligner.c:412-418
                for (int s = 1; s < NSTATE; ++s) tt = log_add(tt, d_add(ff_at(cm, fbase, w, s, k), bt[s * RW + k]), K);
                cm.tbuf[k] = tt;
            }
        }
        PC_SYNC();
        chain_warp0(cm, w, K, false);
        PC_SYNC();
    }
    if (t + 1 <= top) {                                     // matches through t: forward t-1 -> match -> backward t+1
        const int Lq = bd.L[t + 1], wq = bd.co[t + 2] - bd.co[t + 1];
        const int Lf = bd.L[t - 1], fbase = 5 * bd.fo[t - 1], wf = bd.co[t] - bd.co[t - 1];
        const int sm = (Lq - Lf) >> 1;
        const double *bq = slot(cm, t + 1);
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
                cm.tbuf[k] = d_add(vM, bq[k]);               // the other four states of the match-only diagonal are LOG_ZERO
            }
        }
        PC_SYNC();
        chain_warp0(cm, wq, K, true);                        // total = logAdd(total, second dot product), :659
        PC_SYNC();
    }
}

// diagonalCalculationPosteriorMatchProbs, pairwiseAligner.c:676-699: candidates of diagonal t in xmy order.
// Phase 1 (all threads): log posterior of every cell, NaN where it is not a candidate; phase 2 (warp 0): ordered append.
PC_HD void emit_diag(const Job &J, const Band &bd, const CtaMem &cm, const Params &P, int t, Pair *out, int &n_out) {
    const int Lt = bd.L[t], cbase = bd.co[t], w = bd.co[t + 1] - bd.co[t];
    const double *bt = slot(cm, t);
    const double total = *cm.total;
    PC_THREADS(tid, cm.T) {
        for (int k = tid; k < w; k += cm.T) {
            const int xmy = Lt + 2 * k, x = (t + xmy) >> 1, y = (t - xmy) >> 1;
            double lp = not_a_candidate();
            if (x > 0 && y > 0) {
                const double v = d_sub(d_add(fm_at(cm, cbase + k), bt[k]), total);
                if (v >= P.log_thr_lo) lp = v;
            }
            cm.tbuf[k] = lp;
        }
    }
    PC_SYNC();
#if defined(__CUDA_ARCH__)
    if (threadIdx.x < 32) {
        const int lane = (int)threadIdx.x;
        for (int k0 = 0; k0 < w; k0 += 32) {
            const int k = k0 + lane;
            const double lp = k < w ? cm.tbuf[k] : not_a_candidate();
            const bool pred = is_candidate(lp);
            const unsigned m = __ballot_sync(0xffffffffu, pred);
            if (pred) {
                const int pos = n_out + __popc(m & ((1u << lane) - 1u));
                if (pos < J.out_cap) { const int xmy = Lt + 2 * k; Pair p; p.x = ((t + xmy) >> 1) - 1; p.y = ((t - xmy) >> 1) - 1; p.lp = lp; out[pos] = p; }
            }
            n_out += __popc(m);
        }
    }
#else
    for (int k = 0; k < w; ++k) {
        const double lp = cm.tbuf[k];
        if (is_candidate(lp)) {
            if (n_out < J.out_cap) { const int xmy = Lt + 2 * k; Pair p; p.x = ((t + xmy) >> 1) - 1; p.y = ((t - xmy) >> 1) - 1; p.lp = lp; out[n_out] = p; }
            ++n_out;
        }
    }
#endif
}

// getPosteriorProbsWithBanding, pairwiseAligner.c:766-887. Returns the number of candidate pairs (valid in warp 0; may
// exceed out_cap: then only out_cap were stored and the job must be re-run with more room).
PC_HD int run_job(const Job &J, const uint8_t *sym, const int *bandL, const int *coff, const int *foff, const CtaMem &cm, const Params &P,
                  const double *K, Pair *out_all) {
    const int D = J.lx + J.ly;
    if (D == 0) return 0;
    const uint8_t *sx = sym + J.sx_off, *sy = sym + J.sy_off;
    Band bd; bd.L = bandL + J.band_off; bd.co = coff + J.band_off; bd.fo = foff + J.band_off;
    Pair *out = out_all + J.out_off;
    const int RW = cm.RW;
    int n_out = 0;
    {   // diagonal 0: the single cell (0, 0) holds the start state vector (dpDiagonal_initialiseValues, :785-786)
        const double *st = K + ((J.ragged & 1) ? K_RSTART : K_START);
        const int w0 = bd.co[1] - bd.co[0];
        const bool full = bd.fo[1] > bd.fo[0];
        double *cur = slot(cm, 0);
        PC_THREADS(tid, cm.T) {
            for (int k = tid; k < w0; k += cm.T) {
                for (int s = 0; s < NSTATE; ++s) { cur[s * RW + k] = st[s]; if (full) ff_at(cm, 5 * bd.fo[0], w0, s, k) = st[s]; }
                fm_at(cm, bd.co[0] + k) = st[S_M];
            }
        }
        PC_SYNC();
    }
    int tb_to = 0;
    for (int d = 1; d <= D; ++d) {
        fwd_diag(sx, sy, bd, cm, K, d);
        const int w = bd.co[d + 1] - bd.co[d];
        const bool at_end = d == D;
        const bool tb_point = d >= tb_to + P.min_diags && w <= P.expansion * 2 + 1;
        if (!(at_end || tb_point)) continue;
        {   // the diagonal walked back from holds the end state vector (:806-808)
            const double *en = K + ((at_end && (J.ragged & 2)) ? K_REND : K_END);
            double *bt = slot(cm, d);
            PC_THREADS(tid, cm.T) { for (int k = tid; k < w; k += cm.T) for (int s = 0; s < NSTATE; ++s) bt[s * RW + k] = en[s]; }
            PC_SYNC();
        }
        const int tb_from = d - (at_end ? 0 : P.tb_diags + 1);
        int ncalc = 0;
        for (int t = d; t > tb_to; --t) {
            if (t < d) bwd_diag(sx, sy, bd, cm, K, t, d);
            if (t <= tb_from) {
                if (ncalc++ % 10 == 0) total_probability(sx, sy, bd, cm, K, t, d);
                emit_diag(J, bd, cm, P, t, out, n_out);
            }
        }
        tb_to = tb_from;
        if (!at_end) {          // the sweep resumes from the complete forward diagonals d and d-1 (marked by the host)
            PC_SYNC();
            for (int q = 0; q < 2; ++q) {
                const int dd = d - q, wd = bd.co[dd + 1] - bd.co[dd], fbase = 5 * bd.fo[dd];
                double *sl = slot(cm, dd);
                PC_THREADS(tid, cm.T) { for (int k = tid; k < wd; k += cm.T) for (int s = 0; s < NSTATE; ++s) sl[s * RW + k] = ff_at(cm, fbase, wd, s, k); }
            }
            PC_SYNC();
        }
    }
    return n_out;
}

}  // namespace pecan
}  // namespace barb200
