// pecan_warp.cuh -- the banded 5-state pair-HMM forward / backward / posterior program of ONE alignment job, written
// for one warp. It replaces, for a batch of jobs, what the reference does serially in
//   getPosteriorProbsWithBanding ............ submodules/cPecan/impl/pairwiseAligner.c:766-887
//   diagonalCalculation / cellCalculate ..... pairwiseAligner.c:619-634, stateMachine.c:450-480
//   diagonalCalculationTotalProbability ..... pairwiseAligner.c:646-663 (recomputed every 10th diagonal, :840-848)
//   diagonalCalculationPosteriorMatchProbs .. pairwiseAligner.c:676-699
//   logAdd / lookup ......................... pairwiseAligner.c:297-317
//
// Formulation (not a translation of the reference's per-cell object code):
//  * one warp owns one job and walks its x+y diagonals; lanes own cells k, k+32, ... of a diagonal (k = (xmy-xmyL)/2);
//  * the forward matrix lives in a per-warp power-of-two RING in HBM/L2 (state-major per diagonal, so a warp's
//    accesses are contiguous 256-byte runs); only the span between two tracebacks is ever live;
//  * the backward pass is a GATHER: B[t] is computed from the final B[t+1], B[t+2] (the reference scatters from t+2 and
//    t+1 into t; the accumulation order into each target state is reproduced exactly, see bwd_cell) and only three
//    backward diagonals exist at any time;
//  * missing neighbours (outside the band / before the first diagonal) are read as LOG_ZERO cells, which is exact
//    because logAdd(x, LOG_ZERO) == x bit for bit, and the first transition into a state is an assignment because
//    logAdd(LOG_ZERO, v) == v;
//  * all arithmetic is IEEE double without contraction (__dadd_rn / __dmul_rn on the device) in the reference's
//    operation order, so forward, backward and total probabilities are BIT-IDENTICAL to the CPU's; the kernel emits
//    the log posterior (f_M + b_M - total) of every candidate pair and the host applies exp / threshold / floor
//    with the same libm the reference uses (pecan.cu), which makes the integer triples identical as well.
//
// The same source compiles for the host (tests/hosttest) where a warp is emulated by running the 32 lanes of every
// phase one after the other: lanes only communicate through memory between phases, or through PW_* helpers.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PW_LANES(lane) for (int lane = (int)(threadIdx.x & 31u), _pw_once = 1; _pw_once; _pw_once = 0)
#define PW_SYNC() __syncwarp()
#else
#define PW_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define PW_SYNC() ((void)0)
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
    long long band_off;         // first of D+2 entries in bandL / coff
    long long out_off;          // first output record of the job
    int lx, ly;
    int ragged;                 // bit 0: ragged left end, bit 1: ragged right end
    int out_cap;                // output records available
};

struct Pair { int x, y; double lp; };   // 0-based sequence coordinates, log posterior

struct WarpMem {
    double *F;            // forward ring, fmask + 1 doubles
    unsigned fmask;
    double *B;            // 3 backward diagonals, each 5 * ringW doubles (state-major)
    double *tbuf;         // ringW doubles
    int ringW;            // >= widest diagonal of the job
};

PW_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PW_HD double d_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
PW_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PW_HD double log_zero() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)0xfff0000000000000ULL);
#else
    return -INFINITY;
#endif
}

// pairwiseAligner.c:313-317 without branches: big/small by one comparison, interval index by three
PW_HD double log_add(double x, double y, const double *K) {
    const bool lt = x < y;
    const double big = lt ? y : x, small = lt ? x : y;
    const double diff = d_sub(big, small);
    const int idx = (int)(diff > 1.0) + (int)(diff > 2.5) + (int)(diff > 4.5);
    const double *c = K + K_LOOKUP + 4 * idx;
    double r = d_add(d_mul(c[0], diff), c[1]);
    r = d_add(d_mul(r, diff), c[2]);
    r = d_add(d_mul(r, diff), c[3]);
    r = d_add(r, small);
    return (small == log_zero() || diff >= 7.5) ? big : r;
}

PW_HD int match_class(int cx, int cy) { return (cx == 4 || cy == 4) ? 3 : (cx == cy ? 0 : (((cx ^ cy) == 2) ? 1 : 2)); }

// forward ring: diagonal d starts at 5*coff[d]; state s of cell k at (+ s*w + k), all modulo the ring size
PW_HD double &f_at(const WarpMem &wm, int base, int w, int s, int k) { return wm.F[(unsigned)(base + s * w + k) & wm.fmask]; }
PW_HD double *b_diag(const WarpMem &wm, int t) { return wm.B + (size_t)(t % 3) * (5 * (size_t)wm.ringW); }

// ---- forward: diagonal d from d-1 (lower: xmy-1, upper: xmy+1) and d-2 (middle: xmy) -------------------------------------
PW_HD void fwd_diag(const Job &J, const uint8_t *sx, const uint8_t *sy, const int *bL, const int *co, const WarpMem &wm,
                    const double *K, int d) {
    const double LZ = log_zero();
    const int Ld = bL[d], base = 5 * co[d], w = co[d + 1] - co[d];
    const int L1 = bL[d - 1], base1 = 5 * co[d - 1], w1 = co[d] - co[d - 1];
    int L2 = 0, base2 = 0, w2 = 0;
    if (d >= 2) { L2 = bL[d - 2]; base2 = 5 * co[d - 2]; w2 = co[d - 1] - co[d - 2]; }
    const int sl = (Ld - L1 - 1) >> 1, sm = (Ld - L2) >> 1;      // both differences are even
    PW_LANES(lane) {
        for (int k = lane; k < w; k += 32) {
            const int xmy = Ld + 2 * k, x = (d + xmy) >> 1, y = (d - xmy) >> 1;
            const int cx = x > 0 ? sx[x - 1] : 4, cy = y > 0 ? sy[y - 1] : 4;
            const int kl = k + sl, ku = kl + 1, km = k + sm;
            double lM = LZ, lSX = LZ, lLX = LZ, uM = LZ, uSY = LZ, uLY = LZ, mM = LZ, mSX = LZ, mSY = LZ, mLX = LZ, mLY = LZ;
            if (kl >= 0 && kl < w1) { lM = f_at(wm, base1, w1, S_M, kl); lSX = f_at(wm, base1, w1, S_SX, kl); lLX = f_at(wm, base1, w1, S_LX, kl); }
            if (ku >= 0 && ku < w1) { uM = f_at(wm, base1, w1, S_M, ku); uSY = f_at(wm, base1, w1, S_SY, ku); uLY = f_at(wm, base1, w1, S_LY, ku); }
            if (km >= 0 && km < w2) {
                mM = f_at(wm, base2, w2, S_M, km); mSX = f_at(wm, base2, w2, S_SX, km); mSY = f_at(wm, base2, w2, S_SY, km);
                mLX = f_at(wm, base2, w2, S_LX, km); mLY = f_at(wm, base2, w2, S_LY, km);
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
            f_at(wm, base, w, S_M, k) = vM; f_at(wm, base, w, S_SX, k) = vSX; f_at(wm, base, w, S_SY, k) = vSY;
            f_at(wm, base, w, S_LX, k) = vLX; f_at(wm, base, w, S_LY, k) = vLY;
        }
    }
    PW_SYNC();
}

// ---- backward: B[t] gathered from B[t+1] (cells xmy-1 and xmy+1) and B[t+2] (cell xmy); top = diagonal walked from ----------
// Order of accumulation into the target cell c in the reference's scatter (pairwiseAligner.c:619-634 walking xmy upwards,
// stateMachine.c:450-480): while diagonal t+2 is processed c is the MIDDLE of the cell at the same xmy (all five states
// receive from its match state); while t+1 is processed c is first the UPPER of the cell at xmy-1 (M += SY, SY += SY,
// M += LY, LY += LY) and then the LOWER of the cell at xmy+1 (M += SX, SX += SX, M += LX, LX += LX).
PW_HD void bwd_diag(const Job &J, const uint8_t *sx, const uint8_t *sy, const int *bL, const int *co, const WarpMem &wm,
                    const double *K, int t, int top) {
    const double LZ = log_zero();
    const int Lt = bL[t], w = co[t + 1] - co[t], RW = wm.ringW;
    const int L1 = bL[t + 1], w1 = co[t + 2] - co[t + 1];
    const bool has2 = t + 2 <= top;
    int L2 = 0, w2 = 0;
    if (has2) { L2 = bL[t + 2]; w2 = co[t + 3] - co[t + 2]; }
    double *cur = b_diag(wm, t);
    const double *b1 = b_diag(wm, t + 1), *b2 = b_diag(wm, t + 2);
    const int s1 = (Lt - 1 - L1) >> 1, s2 = (Lt - L2) >> 1;
    PW_LANES(lane) {
        for (int k = lane; k < w; k += 32) {
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
    PW_SYNC();
}

// serial logAdd chain over tbuf[0..w) in cell order (dpDiagonal_dotProduct, pairwiseAligner.c:523-534); every lane
// computes the same value
PW_HD double chain(const double *tbuf, int w, const double *K) {
    double tot = log_zero();
    for (int k = 0; k < w; ++k) tot = log_add(tot, tbuf[k], K);
    return tot;
}

// diagonalCalculationTotalProbability, pairwiseAligner.c:646-663
PW_HD double total_probability(const Job &J, const uint8_t *sx, const uint8_t *sy, const int *bL, const int *co, const WarpMem &wm,
                               const double *K, int t, int top) {
    const double LZ = log_zero();
    const int RW = wm.ringW;
    {
        const int base = 5 * co[t], w = co[t + 1] - co[t];
        const double *bt = b_diag(wm, t);
        PW_LANES(lane) {
            for (int k = lane; k < w; k += 32) {
                double tt = d_add(f_at(wm, base, w, 0, k), bt[k]);                       // cell_dotProduct, pairwiseAligner.c:412-418
                for (int s = 1; s < NSTATE; ++s) tt = log_add(tt, d_add(f_at(wm, base, w, s, k), bt[s * RW + k]), K);
                wm.tbuf[k] = tt;
            }
        }
        PW_SYNC();
    }
    double tot = chain(wm.tbuf, co[t + 1] - co[t], K);
    PW_SYNC();
    if (t + 1 <= top) {                                     // matches through t: forward t-1 -> match -> backward t+1
        const int Lq = bL[t + 1], wq = co[t + 2] - co[t + 1];
        const int Lf = bL[t - 1], basef = 5 * co[t - 1], wf = co[t] - co[t - 1];
        const int sm = (Lq - Lf) >> 1;
        const double *bq = b_diag(wm, t + 1);
        PW_LANES(lane) {
            for (int k = lane; k < wq; k += 32) {
                const int xmy = Lq + 2 * k, x = (t + 1 + xmy) >> 1, y = (t + 1 - xmy) >> 1, km = k + sm;
                const int cx = x > 0 ? sx[x - 1] : 4, cy = y > 0 ? sy[y - 1] : 4;
                double mM = LZ, mSX = LZ, mSY = LZ, mLX = LZ, mLY = LZ;
                if (km >= 0 && km < wf) {
                    mM = f_at(wm, basef, wf, S_M, km); mSX = f_at(wm, basef, wf, S_SX, km); mSY = f_at(wm, basef, wf, S_SY, km);
                    mLX = f_at(wm, basef, wf, S_LX, km); mLY = f_at(wm, basef, wf, S_LY, km);
                }
                const double *mt = K + K_MATCH + 3 * match_class(cx, cy);
                double vM = d_add(mM, mt[0]);
                vM = log_add(vM, d_add(mSX, mt[1]), K); vM = log_add(vM, d_add(mSY, mt[1]), K);
                vM = log_add(vM, d_add(mLX, mt[2]), K); vM = log_add(vM, d_add(mLY, mt[2]), K);
                wm.tbuf[k] = d_add(vM, bq[k]);               // the other four states of the match-only diagonal are LOG_ZERO
            }
        }
        PW_SYNC();
        const double tot2 = chain(wm.tbuf, wq, K);
        PW_SYNC();
        tot = log_add(tot, tot2, K);
    }
    return tot;
}

// diagonalCalculationPosteriorMatchProbs, pairwiseAligner.c:676-699: candidates of diagonal t in xmy order
PW_HD void emit_diag(const Job &J, const int *bL, const int *co, const WarpMem &wm, const Params &P, int t, double total,
                     Pair *out, int &n_out, int &overflow) {
    const int Lt = bL[t], base = 5 * co[t], w = co[t + 1] - co[t];
    const double *bt = b_diag(wm, t);
    for (int k0 = 0; k0 < w; k0 += 32) {
#if defined(__CUDA_ARCH__)
        const int lane = (int)(threadIdx.x & 31u), k = k0 + lane;
        bool pred = false; int x = 0, y = 0; double lp = 0;
        if (k < w) {
            const int xmy = Lt + 2 * k; x = (t + xmy) >> 1; y = (t - xmy) >> 1;
            if (x > 0 && y > 0) { lp = d_sub(d_add(f_at(wm, base, w, 0, k), bt[k]), total); pred = lp >= P.log_thr_lo; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (pred) {
            const int pos = n_out + __popc(m & ((1u << lane) - 1u));
            if (pos < J.out_cap) { Pair p; p.x = x - 1; p.y = y - 1; p.lp = lp; out[pos] = p; }
        }
        n_out += __popc(m);
#else
        for (int lane = 0; lane < 32; ++lane) {
            const int k = k0 + lane;
            if (k >= w) break;
            const int xmy = Lt + 2 * k, x = (t + xmy) >> 1, y = (t - xmy) >> 1;
            if (x > 0 && y > 0) {
                const double lp = d_sub(d_add(f_at(wm, base, w, 0, k), bt[k]), total);
                if (lp >= P.log_thr_lo) {
                    if (n_out < J.out_cap) { Pair p; p.x = x - 1; p.y = y - 1; p.lp = lp; out[n_out] = p; }
                    ++n_out;
                }
            }
        }
#endif
    }
    if (n_out > J.out_cap) overflow = 1;
}

// getPosteriorProbsWithBanding, pairwiseAligner.c:766-887. Returns the number of candidate pairs (may exceed out_cap:
// then only out_cap were stored and the job must be re-run with more room).
PW_HD int run_job(const Job &J, const uint8_t *sym, const int *bandL, const int *coff, const WarpMem &wm, const Params &P,
                  const double *K, Pair *out_all, long long *cells_done) {
    const int D = J.lx + J.ly;
    if (D == 0) return 0;
    const uint8_t *sx = sym + J.sx_off, *sy = sym + J.sy_off;
    const int *bL = bandL + J.band_off, *co = coff + J.band_off;
    Pair *out = out_all + J.out_off;
    int n_out = 0, overflow = 0;
    {   // diagonal 0: the single cell (0, 0) holds the start state vector (dpDiagonal_initialiseValues, :785-786)
        const double *st = K + ((J.ragged & 1) ? K_RSTART : K_START);
        const int w0 = co[1] - co[0];
        PW_LANES(lane) { for (int k = lane; k < w0; k += 32) for (int s = 0; s < NSTATE; ++s) f_at(wm, 5 * co[0], w0, s, k) = st[s]; }
        PW_SYNC();
    }
    int tb_to = 0;
    for (int d = 1; d <= D; ++d) {
        fwd_diag(J, sx, sy, bL, co, wm, K, d);
        const int w = co[d + 1] - co[d];
        const bool at_end = d == D;
        const bool tb_point = d >= tb_to + P.min_diags && w <= P.expansion * 2 + 1;
        if (!(at_end || tb_point)) continue;
        {   // the diagonal walked back from holds the end state vector (:806-808)
            const double *en = K + ((at_end && (J.ragged & 2)) ? K_REND : K_END);
            double *bt = b_diag(wm, d);
            PW_LANES(lane) { for (int k = lane; k < w; k += 32) for (int s = 0; s < NSTATE; ++s) bt[s * wm.ringW + k] = en[s]; }
            PW_SYNC();
        }
        const int tb_from = d - (at_end ? 0 : P.tb_diags + 1);
        double total = log_zero();
        int ncalc = 0;
        for (int t = d; t > tb_to; --t) {
            if (t < d) bwd_diag(J, sx, sy, bL, co, wm, K, t, d);
            if (t <= tb_from) {
                if (ncalc++ % 10 == 0) total = total_probability(J, sx, sy, bL, co, wm, K, t, d);
                emit_diag(J, bL, co, wm, P, t, total, out, n_out, overflow);
            }
        }
        tb_to = tb_from;
    }
    if (cells_done) *cells_done = co[D + 1];
    (void)overflow;
    return n_out;
}

}  // namespace pecan
}  // namespace barb200
