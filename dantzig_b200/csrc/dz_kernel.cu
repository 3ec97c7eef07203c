// dz_kernel.cu -- persistent parametric self-dual simplex kernel for sm_100a.
//
// One TEAM (a CTA, or a single warp) solves one LP at a time, pulling LP ids from
// a work queue; the whole pivot loop of Simplex::solve
// (/root/reference/src/simplex.rs:274-343) runs on the device with no host round
// trip.  Per pivot the team
//   * gathers the basis from the CSC template into a dense working matrix W
//     (replaces basis_matrix/collect_columns/to_dense/t, simplex.rs:270-272,
//     linalg.rs:40-48,131-140,188-192),
//   * eliminates [W | rhs] with partial pivoting in the reference's exact
//     operation order (Matrix::factorize linalg.rs:88-128 fused with the forward
//     half of LU::solve linalg.rs:286-291), once for B and once for B^T,
//   * back-substitutes (linalg.rs:292-297),
//   * prices dz = -N^T v over the CSC columns (neg_t_dot linalg.rs:199-207),
//   * runs the ratio tests / arg-max rules with first-index tie-breaks
//     (find_first_pivot, find_second_pivot simplex.rs:423-461) and the four
//     vector updates (pivot simplex.rs:253-268,410-421).
//
// EXACTNESS CONTRACT.  Results must be bit-identical to the reference's CPU
// arithmetic, so every floating point operation of the reference is performed
// in binary64 with round-to-nearest, unfused (__dmul_rn/__dsub_rn/__ddiv_rn;
// the file is also compiled with --fmad=false) and in the reference's order
// wherever order matters.  The only liberties taken are ones that cannot
// change a bit of any value the algorithm observes:
//   * operations with an exact-zero factor and a finite co-factor are skipped;
//   * elimination steps whose pivot column holds a single nonzero in a row that
//     has not been used yet ("virgin unit" steps: every slack column is one
//     until its row is stolen) are pure bookkeeping and are retired by one
//     control thread without touching the matrix;
//   * L is never stored: the right-hand side rides along as column M of W, so
//     the forward substitution happens inside the elimination;
//   * rows are never physically swapped: a logical-position table records the
//     interchanges (positions decide the pivot-search tie-break);
//   * division by an exact 1.0 (every slack pivot) is the identity and is elided.
// When a back-substitution produces a non-finite value the skipping rules are
// no longer no-ops (0*inf = NaN), so that solve is redone without skipping.
//
// LAUNCH SHAPES (template <HOME, WARP>, chosen by plan_launch; all bit-identical):
//   WARP=1          one warp per LP, W in the HBM workspace, no CTA barrier at all
//                   (default for large batches with m_int <= 128: config 2);
//   WARP=0, HOME=0  CTA per LP (G worker warps + a control warp), W and the per-LP
//                   vectors in shared memory (single LPs / small batches);
//   WARP=0, HOME=1  CTA per LP, W in the HBM workspace (config 5, config 1);
//   WARP=0, HOME=2  CTA per LP, W and all per-LP vectors in HBM (config 3).
// The kernel is bound by dependent load/issue latency per LP, so throughput comes
// from LPs in flight per SM and from fetching operands as batches of independent
// loads (DESIGN.md section 4 has the measurements).
//
// THREAD MAP of an elimination step.  Pivot search: every warp scans the whole
// pivot column (threads stand for rows; warps split the rows when m_int > 384)
// and reduces the bit patterns of |v| with REDUX.  Multipliers: worker warp w
// owns rows lane*G + w.  Trailing update: lanes take columns (row-contiguous
// access), rows with a nonzero multiplier go two at a time.  The control thread
// records interchanges and runs ahead over virgin-unit steps; it and the serial
// part of back-substitution live on the control warp (the warp itself if WARP).

#include "dz_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

// Build-time switches.  The round-1 variants that lost their A/B on the GPU (tiled warp step,
// compacted back-substitution, 32-bit offsets, batched pricing loads, no-inline, opaque bases:
// profiles/r02_experiments.md) are gone; what is left:
//   DZ_KERNEL_PER_NR one warp-per-LP kernel per value of ceil(m_int/32) (1,2,3,4,5,6,8) holding only
//                    its own fast paths, instead of one kernel with all of them plus the general
//                    step: a third of the code per kernel (instruction cache); +5 % on config 2,
//                    on by default
//   DZ_STEP_PROFILE  with opt.profile, warp_step_small splits its cycles into the PH_E_* slots
#ifndef DZ_KERNEL_PER_NR
#define DZ_KERNEL_PER_NR 1
#endif
#define DZ_HOT_FN __device__ __forceinline__
#ifndef DZ_STEP_PROFILE
#define DZ_STEP_PROFILE 0
#endif

namespace dz {

namespace {

struct Ctx {
    // problem
    int M, Nn, S, NWK, G, nthreads, nwarps;
    int tid;  // thread index within the team that solves one LP (CTA, or a single warp)
    bool wm;  // warp mode: one warp per LP, team barrier = __syncwarp
    // shared-memory arrays
    double *W;
    double *x, *xb, *dxv, *vv, *z, *zb, *dzv;
    double *red_key;
    int *bas, *nb, *rowAt, *posOf, *cnt, *unitRow, *pend, *pre, *red_idx, *ctl;
    // Large single LPs (HOME == 2): per row of W the half-open column interval
    // [rlo, rhi) outside of which the row is known to be exactly zero (column M,
    // the right-hand side, is tracked separately).  W really holds zeros there, so
    // the intervals only prune loads, zero-fills and loops; they never change values.
    int *rlo, *rhi;
    double *pbuf; // [M + 32 * kMaxWarps] ordered nonzero products of one back-substitution row
    int *plist;   // [M] pending back-substitution rows, descending position
    double *lval; // [nnz] this LP's lowered values, resolved once from theta (HOME == 2)
    int *cand_r;  // [M] rows with a nonzero in the current pivot column (found by the search)
    double *cand_v; // [M] ... and their values
    int parity;
    // optional phase timing (thread 0 accumulates clock64 deltas into shared memory)
    long long *prof;
    long long t_last;
    // counters (per thread)
    unsigned long long n_lu, n_solve, n_price;
};

enum { CTL_K = 0, CTL_FLAG = 1, CTL_LP = 2, CTL_RHS0 = 3, CTL_NC = 4 };

// Phase slots of the optional per-LP cycle profile (BatchDev::prof).
enum {
    PH_STATUS = 0, PH_GATHER = 1, PH_ELIM = 2, PH_BACK_A = 3, PH_BACK_B = 4, PH_PRICE = 5,
    PH_RATIO = 6, PH_UPDATE = 7, PH_NONTRIVIAL = 8, PH_PENDING = 9, PH_SOLVES = 10,
    PH_E_SEARCH = 11, PH_E_B2 = 12, PH_E_UPD = 13, PH_E_B1 = 14, PH_COUNT = 16
};

__device__ __forceinline__ bool in_iv(const Ctx &c, int r, int j) {
    return !c.rlo || (j >= c.rlo[r] && j < c.rhi[r]);
}

// Interval mode: zero exactly the columns a solve may have made nonzero, row by row,
// and reset the intervals; afterwards W is all-zero again (column M included).
__device__ __forceinline__ void zero_intervals(Ctx &c) {
    const int lane = c.tid & 31, warp = c.tid >> 5;
    for (int r = warp; r < c.M; r += c.nwarps) {
        const int lo = c.rlo[r], hi = c.rhi[r];
        double *row = c.W + (size_t)r * c.S;
        if (lo < hi)
            for (int j = lo + lane; j < hi; j += 32) row[j] = 0.0;
        __syncwarp();
        if (lane == 0) {
            row[c.M] = 0.0;
            c.rlo[r] = 0x7fffffff;
            c.rhi[r] = 0;
        }
    }
}

// Team barrier: the CTA in CTA-per-LP mode, the warp in warp-per-LP mode.
__device__ __forceinline__ void csync(const Ctx &c) {
    if (c.wm)
        __syncwarp();
    else
        __syncthreads();
}

__device__ __forceinline__ void tick(Ctx &c, int slot) {
    if (c.prof && c.tid == 0) {
        const long long now = clock64();
        c.prof[slot] += now - c.t_last;
        c.t_last = now;
    }
}

// find_first_pivot (simplex.rs:423-437) for both the dual (z) and primal (x)
// sides with a single barrier.  Returns positions (-1 = None).  The reference's
// sequential reduce keeps its FIRST element when that element's ratio is NaN
// (nothing compares greater than NaN); otherwise NaN ratios never win.
__device__ __forceinline__ void find_first_both(Ctx &c, int &q0, int &p0) {
    Cand<4> cd;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        cd.key[n] = 0.0;
        cd.idx[n] = -1;
    }
    for (int k = c.tid; k < c.Nn; k += c.nthreads) {
        const double yb = c.zb[k];
        if (yb > 0.0) {
            const double ratio = (yb == 1.0) ? -c.z[k] : __ddiv_rn(-c.z[k], yb); // x/1 == x
            if (ratio == ratio && beats(ratio, k, cd.key[0], cd.idx[0])) {
                cd.key[0] = ratio;
                cd.idx[0] = k;
            }
            if (cd.idx[1] < 0) cd.idx[1] = k; // k ascends per thread
        }
    }
    for (int k = c.tid; k < c.M; k += c.nthreads) {
        const double yb = c.xb[k];
        if (yb > 0.0) {
            const double ratio = (yb == 1.0) ? -c.x[k] : __ddiv_rn(-c.x[k], yb);
            if (ratio == ratio && beats(ratio, k, cd.key[2], cd.idx[2])) {
                cd.key[2] = ratio;
                cd.idx[2] = k;
            }
            if (cd.idx[3] < 0) cd.idx[3] = k;
        }
    }
    block_argmax<4>(cd, c.red_key, c.red_idx, c.parity, c.nwarps, c.tid, c.wm);
    q0 = cd.idx[0];
    if (cd.idx[1] >= 0) {
        const double r = __ddiv_rn(-c.z[cd.idx[1]], c.zb[cd.idx[1]]);
        if (r != r) q0 = cd.idx[1];
    }
    p0 = cd.idx[2];
    if (cd.idx[3] >= 0) {
        const double r = __ddiv_rn(-c.x[cd.idx[3]], c.xb[cd.idx[3]]);
        if (r != r) p0 = cd.idx[3];
    }
}

// find_second_pivot (simplex.rs:439-461): arg-max of dy/(y + mu*ybar) over the
// strictly positive ratios, first index wins.  Returns the position or -1.
__device__ __forceinline__ int find_second(Ctx &c, double mu, const double *y, const double *yb,
                                           const double *dy, int len) {
    Cand<1> cd;
    cd.key[0] = 0.0;
    cd.idx[0] = -1;
    for (int k = c.tid; k < len; k += c.nthreads) {
        const double denom = __dadd_rn(y[k], __dmul_rn(mu, yb[k]));
        const double ratio = __ddiv_rn(dy[k], denom);
        if (ratio > 0.0 && beats(ratio, k, cd.key[0], cd.idx[0])) {
            cd.key[0] = ratio;
            cd.idx[0] = k;
        }
    }
    block_argmax<1>(cd, c.red_key, c.red_idx, c.parity, c.nwarps, c.tid, c.wm);
    return cd.idx[0];
}

// Back substitution (linalg.rs:292-297) over the eliminated [W | rhs].
// literal == false: rows whose strict upper part is all zero are solved by
// their own threads up front; the others run serially on the control warp with
// exact-zero products skipped.  literal == true: every term, strictly serial.
// Sets ctl[CTL_FLAG] when a non-finite value is produced.
__device__ __forceinline__ void back_substitute(Ctx &c, double *y, bool literal) {
    const int M = c.M, S = c.S, tid = c.tid;
    double *W = c.W;
    if (tid < c.NWK) {
        for (int r = tid; r < M; r += c.NWK) {
            const int i = c.posOf[r];
            bool has_nz = literal;
            if (!literal) {
                const double *row = W + (size_t)r * S;
                int jb = i + 1, je = M;
                if (c.rlo) {
                    jb = max(jb, c.rlo[r]);
                    je = min(je, c.rhi[r]);
                }
                for (int j = jb; j < je && !has_nz; j += 4) { // four independent loads per trip
                    const double e0 = row[j];
                    const double e1 = (j + 1 < je) ? row[j + 1] : 0.0;
                    const double e2 = (j + 2 < je) ? row[j + 2] : 0.0;
                    const double e3 = (j + 3 < je) ? row[j + 3] : 0.0;
                    has_nz = e0 != 0.0 || e1 != 0.0 || e2 != 0.0 || e3 != 0.0;
                }
            }
            c.pend[i] = has_nz ? 1 : 0;
            if (!has_nz) {
                const double di = in_iv(c, r, i) ? W[(size_t)r * S + i] : 0.0;
                const double bi = W[(size_t)r * S + M];
                const double yi = (di == 1.0) ? bi : __ddiv_rn(bi, di); // x/1 is the identity
                y[i] = yi;
                if (!isfinite(yi)) c.ctl[CTL_FLAG] = 1;
                c.n_solve += 1;
            }
        }
    }
    csync(c);
    tick(c, PH_BACK_A);
    if (c.pbuf) {
        // Large single LP: the products u_ij * y_j of one row are formed by ALL warps
        // (each warp a contiguous column segment, compacted in order), then the control
        // warp runs the strictly sequential subtraction chain over the nonzero products
        // (linalg.rs:294: j ascending).  Rows still go one after the other, M-1..0.
        const int lane = tid & 31, warp = tid >> 5, nw = c.nwarps;
        int *wcnt = c.red_idx; // [kMaxWarps] per-warp product counts
        if (tid >= c.NWK) { // control warp: pending rows, descending, into plist
            int n = 0;
            const int cl = tid - c.NWK;
            for (int base = (M - 1) & ~31; base >= 0; base -= 32) {
                const int il = base + 31 - cl; // lane 0 takes the highest position
                const unsigned pm = __ballot_sync(kFull, il < M && c.pend[il] != 0);
                if (il < M && c.pend[il] != 0) c.plist[n + __popc(pm & ((1u << cl) - 1u))] = il;
                n += __popc(pm);
            }
            if (cl == 0) c.ctl[CTL_K] = n;
        }
        csync(c);
        const int np = c.ctl[CTL_K];
        for (int e = 0; e < np; ++e) {
            const int i = c.plist[e];
            const int rrow = c.rowAt[i];
            const double *row = W + (size_t)rrow * S;
            int jb = i + 1, je = M;
            if (c.rlo && !literal) { // literal mode must see every 0 * y_j (NaN if y_j = inf)
                jb = max(jb, c.rlo[rrow]);
                je = min(je, c.rhi[rrow]);
            }
            const int width = max(je - jb, 0);
            const int seg = (((width + nw - 1) / nw) + 31) & ~31; // columns per warp, multiple of 32
            {
                const int s0 = jb + warp * seg, s1 = min(s0 + seg, je);
                int n = 0;
                for (int j0 = s0; j0 < s1; j0 += 128) { // four chunks of loads in flight
                    double u4[4], y4[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int j = j0 + lane + 32 * cc;
                        u4[cc] = (j < s1) ? row[j] : 0.0;
                        y4[cc] = (j < s1) ? y[j] : 0.0;
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int j = j0 + lane + 32 * cc;
                        const bool on = literal ? (j < s1) : (u4[cc] != 0.0 && y4[cc] != 0.0);
                        const unsigned mk = __ballot_sync(kFull, on);
                        if (on)
                            c.pbuf[warp * seg + n + __popc(mk & ((1u << lane) - 1u))] =
                                __dmul_rn(u4[cc], y4[cc]);
                        n += __popc(mk);
                    }
                }
                if (lane == 0) wcnt[warp] = n;
            }
            csync(c);
            if (tid >= c.NWK) { // control warp: the sequential chain
                const int cl = tid - c.NWK;
                double s = row[M];
                const double di = in_iv(c, rrow, i) ? row[i] : 0.0;
                unsigned long long ops = 0;
                for (int w = 0; w < nw; ++w) {
                    const int n = wcnt[w];
                    // every lane walks the same compacted products (uniform, cached loads:
                    // one line holds 16 of them); the loads run ahead of the DSUB chain
                    const double *pb = c.pbuf + (size_t)w * seg;
                    int q = 0;
                    for (; q + 4 <= n; q += 4) {
                        const double t0 = pb[q], t1 = pb[q + 1], t2 = pb[q + 2], t3 = pb[q + 3];
                        s = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(s, t0), t1), t2), t3);
                    }
                    for (; q < n; ++q) s = __dsub_rn(s, pb[q]);
                    ops += 2ull * n;
                }
                const double yi = (di == 1.0) ? s : __ddiv_rn(s, di);
                if (cl == 0) {
                    y[i] = yi;
                    if (!isfinite(yi)) c.ctl[CTL_FLAG] = 1;
                    c.n_solve += ops + 1;
                    if (c.prof) c.prof[PH_PENDING] += 1;
                }
            }
            csync(c); // y[i] visible, pbuf / wcnt free
        }
        tick(c, PH_BACK_B);
        return;
    }
    if (c.wm || tid >= c.NWK) { // control warp (the warp itself in warp mode)
        const int lane = c.wm ? tid : tid - c.NWK;
        for (int base = (M - 1) & ~31; base >= 0; base -= 32) {
            const int il = base + lane;
            unsigned pm = __ballot_sync(kFull, il < M && c.pend[il] != 0);
            while (pm) {
                const int bit = 31 - __clz(pm);
                pm &= ~(1u << bit);
                const int i = base + bit;
                if (c.prof && lane == 0) c.prof[PH_PENDING] += 1;
                const int rrow = c.rowAt[i];
                const double *row = W + (size_t)rrow * S;
                double s = row[M];
                const double di = in_iv(c, rrow, i) ? row[i] : 0.0;
                int jb = i + 1, je = M;
                if (c.rlo && !literal) { // literal mode must see every 0 * y_j (NaN if y_j = inf)
                    jb = max(jb, c.rlo[rrow]);
                    je = min(je, c.rhi[rrow]);
                }
                for (int j0 = jb; j0 < je; j0 += 128) { // four chunks of loads in flight
                    double u4[4], y4[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int j = j0 + lane + 32 * cc;
                        u4[cc] = (j < je) ? row[j] : 0.0;
                        y4[cc] = (j < je) ? y[j] : 0.0;
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int j = j0 + lane + 32 * cc;
                        const double t = __dmul_rn(u4[cc], y4[cc]);
                        unsigned mk = __ballot_sync(
                            kFull, literal ? (j < je) : (u4[cc] != 0.0 && y4[cc] != 0.0));
                        if (lane == 0) c.n_solve += 2ull * __popc(mk);
                        while (mk) {
                            const int b = __ffs(mk) - 1;
                            mk &= mk - 1;
                            s = __dsub_rn(s, __shfl_sync(kFull, t, b));
                        }
                    }
                }
                const double yi = (di == 1.0) ? s : __ddiv_rn(s, di);
                if (lane == 0) {
                    y[i] = yi;
                    if (!isfinite(yi)) c.ctl[CTL_FLAG] = 1;
                    c.n_solve += 1;
                }
                __syncwarp();
            }
        }
    }
    csync(c);
    tick(c, PH_BACK_B);
}


// Back substitution of the warp-per-LP mode for M <= 128: rows strictly in order
// M-1..0 (linalg.rs:292-297), each row's strict upper part fetched as one batch of
// independent loads and the NEXT row prefetched while the current chain runs.
// Products with an exact-zero factor are skipped unless `literal`; division by an
// exact 1.0 (every slack pivot) is the identity and is elided.
template <int NR>
DZ_HOT_FN void warp_back_substitute_small(Ctx &c, double *y, const bool literal) {
    const int M = c.M, S = c.S, lane = c.tid;
    const double *__restrict__ W = c.W;
    double uu[NR], nu[NR], d = 0.0, rhs = 0.0, nd = 0.0, nrhs = 0.0;
    {
        const double *rowp = W + (size_t)c.rowAt[M - 1] * S;
#pragma unroll
        for (int cc = 0; cc < NR; ++cc) nu[cc] = 0.0;
        nd = rowp[M - 1];
        nrhs = rowp[M];
    }
    unsigned long long ops = 0;
    for (int i = M - 1; i >= 0; --i) {
#pragma unroll
        for (int cc = 0; cc < NR; ++cc) uu[cc] = nu[cc];
        d = nd;
        rhs = nrhs;
        if (i > 0) { // prefetch row i-1 (its U entries do not depend on y)
            const double *rowp = W + (size_t)c.rowAt[i - 1] * S;
#pragma unroll
            for (int cc = 0; cc < NR; ++cc) {
                const int j = i + lane + 32 * cc;
                nu[cc] = (j < M) ? rowp[j] : 0.0;
            }
            nd = rowp[i - 1];
            nrhs = rowp[M];
        }
        double s = rhs;
#pragma unroll
        for (int cc = 0; cc < NR; ++cc) {
            const int j = i + 1 + lane + 32 * cc;
            if (i + 1 + 32 * cc >= M) break;
            const double yj = (j < M) ? y[j] : 0.0;
            const double t = __dmul_rn(uu[cc], yj);
            unsigned mk = __ballot_sync(kFull, literal ? (j < M) : (uu[cc] != 0.0 && yj != 0.0));
            ops += 2ull * __popc(mk);
            while (mk) {
                const int b = __ffs(mk) - 1;
                mk &= mk - 1;
                s = __dsub_rn(s, __shfl_sync(kFull, t, b));
            }
        }
        const double yi = (d == 1.0) ? s : __ddiv_rn(s, d);
        if (lane == 0) {
            y[i] = yi;
            if (!isfinite(yi)) c.ctl[CTL_FLAG] = 1;
        }
        ops += 1;
        __syncwarp();
    }
    if (lane == 0) c.n_solve += ops;
    __syncwarp();
    tick(c, PH_BACK_B);
}

// One elimination step of the warp-per-LP mode for M <= 32*NR <= 128 (at most NR rows
// and NR pivot-row chunks per lane).  Same arithmetic as the general step below; the
// point is memory-level parallelism: the column, the pivot row and the rows to
// update are fetched with batches of independent loads (the basis lives in the
// HBM/L2 workspace, ~0.3-0.8 us away) instead of one dependent load at a time,
// and the values read by the pivot search are reused for the multipliers.
template <int NR>
DZ_HOT_FN void warp_step_small(Ctx &c, double *__restrict__ W, const int k, const bool is_ctl,
                               double *__restrict__ scratch) {
    const int M = c.M, S = c.S, lane = c.tid;
#if DZ_STEP_PROFILE
    long long tp = (c.prof && lane == 0) ? clock64() : 0;
#define DZ_STEP_TICK(slot)                                   \
    if (c.prof && lane == 0) {                               \
        const long long tn = clock64();                      \
        c.prof[slot] += tn - tp;                             \
        tp = tn;                                             \
    }
#else
#define DZ_STEP_TICK(slot)
#endif
    int pos[NR];
    double v[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int r = lane + 32 * i;
        pos[i] = (r < M) ? c.posOf[r] : -1;
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int r = lane + 32 * i;
        v[i] = (pos[i] >= k) ? W[(size_t)r * S + k] : 0.0;
    }
    unsigned bhi = 0u, blo = 0u;
    int bidx = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int r = lane + 32 * i;
        bool cand = pos[i] >= k;
        unsigned hi = (unsigned)__double2hiint(v[i]) & 0x7fffffffu, lo = (unsigned)__double2loint(v[i]);
        if (v[i] != v[i]) { // NaN: only as the incumbent at position k (linalg.rs:99-105)
            cand = cand && pos[i] == k;
            hi = 0xffffffffu;
            lo = 0xffffffffu;
        }
        const int packed = (pos[i] << 16) | r;
        const bool better = cand && (bidx == 0x7fffffff || hi > bhi ||
                                     (hi == bhi && (lo > blo || (lo == blo && packed < bidx))));
        bhi = better ? hi : bhi;
        blo = better ? lo : blo;
        bidx = better ? packed : bidx;
    }
    const unsigned mh = __reduce_max_sync(kFull, bhi);
    const unsigned ml = __reduce_max_sync(kFull, bhi == mh ? blo : 0u);
    const int gi = __reduce_min_sync(kFull, (bhi == mh && blo == ml) ? bidx : 0x7fffffff);
    const int pr = gi & 0xffff, ppos = gi >> 16;
    DZ_STEP_TICK(PH_E_SEARCH) // column loads + arg-max
    const double *__restrict__ prow = W + (size_t)pr * S;
    const double pv = prow[k];
    // pivot row, up to four chunks of 32 columns (column M is the right-hand side)
    double u[NR];
#pragma unroll
    for (int cc = 0; cc < NR; ++cc) {
        const int j = k + 1 + lane + 32 * cc;
        u[cc] = (j <= M) ? prow[j] : 0.0;
    }
    __syncwarp();
    if (is_ctl) { // record the interchange k <-> ppos (linalg.rs:107-114)
        if (c.prof) c.prof[PH_NONTRIVIAL] += 1;
        const int rk = c.rowAt[k];
        c.rowAt[k] = pr;
        c.rowAt[ppos] = rk;
        c.posOf[pr] = k;
        c.posOf[rk] = ppos;
    }
    if (pv == 0.0) return; // linalg.rs:117
    unsigned nzu = 0;
#pragma unroll
    for (int cc = 0; cc < NR; ++cc) nzu += (u[cc] != 0.0) ? 1u : 0u;
    DZ_STEP_TICK(PH_E_B2) // pivot row loads + interchange bookkeeping
    unsigned long long upd = 0;
    unsigned nrows = 0;
    bool nz[NR];
#pragma unroll
    for (int cc = 0; cc < NR; ++cc) nz[cc] = u[cc] != 0.0;
    double *__restrict__ Wb = W + k + 1 + lane; // row r, chunk cc: Wb[r * S + 32 * cc] (int offsets: M <= 256)
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const int r = lane + 32 * i;
        const bool need = pos[i] >= k && r != pr && v[i] != 0.0;
        unsigned rows = __ballot_sync(kFull, need);
        if (!rows) continue;
        const double l = need ? __ddiv_rn(v[i], pv) : 0.0;
        upd += need ? 1 : 0;
        nrows += __popc(rows);
        while (rows) {
            const int b0 = __ffs(rows) - 1;
            rows &= rows - 1;
            const bool two = rows != 0;
            const int b1 = two ? __ffs(rows) - 1 : b0;
            rows &= rows - 1; // no-op when already empty
            const double l0 = __shfl_sync(kFull, l, b0);
            const double l1 = __shfl_sync(kFull, l, b1);
            double *__restrict__ w0 = Wb + (unsigned)(b0 + 32 * i) * (unsigned)S;
            double *__restrict__ w1 = Wb + (unsigned)(b1 + 32 * i) * (unsigned)S;
            double a0[NR], a1[NR];
#pragma unroll
            for (int cc = 0; cc < NR; ++cc) {
                a0[cc] = nz[cc] ? w0[32 * cc] : 0.0;
                a1[cc] = (two && nz[cc]) ? w1[32 * cc] : 0.0;
            }
#pragma unroll
            for (int cc = 0; cc < NR; ++cc) {
                if (nz[cc]) {
                    w0[32 * cc] = __dsub_rn(a0[cc], __dmul_rn(l0, u[cc]));
                    if (two) w1[32 * cc] = __dsub_rn(a1[cc], __dmul_rn(l1, u[cc]));
                }
            }
        }
    }
    upd += 2ull * nzu * nrows;
    c.n_lu += upd;
    DZ_STEP_TICK(PH_E_UPD) // multipliers + row-pair updates
#undef DZ_STEP_TICK
}

// lu_solve (linalg.rs:8-10) of B (transposed == false, rhs = column `arg` of A)
// or of B^T (transposed == true, rhs = e_arg).  Result in y[0..M).
template <int WARP_NR_MAX, int NRX = 0> // NRX > 0: instantiated for exactly ceil(m_int/32) == NRX (DZ_KERNEL_PER_NR)
__device__ __forceinline__ void basis_solve(Ctx &c, const TemplateDev &T,
                                            const double *__restrict__ theta, bool transposed,
                                            int arg, double *y) {
    const int M = c.M, S = c.S, tid = c.tid;
    const int lane = tid & 31, warp = tid >> 5;
    double *W = c.W;

    // ---- gather: W = dense(B) or dense(B^T), rhs in column M ----------------
    // pre[p] = first flat entry of basis position p; position M is the rhs column.
    {
        if (c.rlo) {
            zero_intervals(c); // only what the previous solve may have made nonzero
        } else {
            const size_t total = (size_t)M * S; // W is 16-byte aligned in both homes
            double2 *W2 = reinterpret_cast<double2 *>(W);
            for (size_t e = tid; e < (total >> 1); e += c.nthreads) W2[e] = make_double2(0.0, 0.0);
            if ((total & 1) && tid == 0) W[total - 1] = 0.0;
        }
        for (int i = tid; i < M; i += c.nthreads) {
            c.cnt[i] = 0;
            c.unitRow[i] = -1;
            c.rowAt[i] = i;
            c.posOf[i] = i;
        }
        if (tid == 0) c.ctl[CTL_FLAG] = 0;
        if (warp == c.nwarps - 1) { // exclusive scan of the column lengths (control warp)
            // pend[p] doubles as the first CSC entry of basis position p until the
            // back-substitution needs it; CTL_RHS0 is the same for the rhs column
            const int per = (M + 1 + 31) >> 5;
            const int b0 = lane * per;
            int sum = 0;
            for (int i = 0; i < per; ++i) {
                const int p = b0 + i;
                int len = 0;
                if (p < M) {
                    const int col = c.bas[p];
                    const int e0 = T.col_ptr[col];
                    len = T.col_ptr[col + 1] - e0;
                    c.pend[p] = e0;
                } else if (p == M && !transposed) {
                    const int e0 = T.col_ptr[arg];
                    len = T.col_ptr[arg + 1] - e0;
                    c.ctl[CTL_RHS0] = e0;
                }
                c.pre[p <= M ? p : M + 1] = len; // lengths first, prefix below
                sum += len;
            }
            __syncwarp();
            int incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, off);
                if (lane >= off) incl += t;
            }
            int run = incl - sum;
            for (int i = 0; i < per; ++i) {
                const int p = b0 + i;
                if (p <= M) {
                    const int len = c.pre[p];
                    c.pre[p] = run;
                    run += len;
                }
            }
            __syncwarp();
            if (lane == 31) c.pre[M + 1] = incl;
        }
    }
    csync(c);
    {
        // flat enumeration of every entry of every basis column (and of the rhs
        // column): four independent entries per thread in flight
        const int total = c.pre[M + 1];
        for (int base = tid; base < total; base += 4 * c.nthreads) {
            int pp[4], ee[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = base + q * c.nthreads;
                int lo = 0, hi = M;
                if (idx < total) {
                    while (lo < hi) { // largest p in [0, M] with pre[p] <= idx
                        const int mid = (lo + hi + 1) >> 1;
                        if (c.pre[mid] <= idx)
                            lo = mid;
                        else
                            hi = mid - 1;
                    }
                    pp[q] = lo;
                    ee[q] = (lo < M ? c.pend[lo] : c.ctl[CTL_RHS0]) + (idx - c.pre[lo]);
                } else {
                    pp[q] = -1;
                    ee[q] = 0;
                }
            }
            int ref[4], row[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ref[q] = pp[q] >= 0 ? T.val_ref[ee[q]] : -1;
                row[q] = pp[q] >= 0 ? T.row_idx[ee[q]] : 0;
            }
            double val[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                val[q] = (c.lval && pp[q] >= 0) ? c.lval[ee[q]] : load_ref(theta, ref[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (pp[q] < 0 || val[q] == 0.0) continue;
                if (pp[q] == M) {
                    W[(size_t)row[q] * S + M] = val[q];
                } else {
                    const int wr = transposed ? pp[q] : row[q], wc = transposed ? row[q] : pp[q];
                    W[(size_t)wr * S + wc] = val[q];
                    atomicAdd(&c.cnt[wc], 1);
                    c.unitRow[wc] = wr; // meaningful only where cnt ends at 1
                    if (c.rlo) {
                        atomicMin(&c.rlo[wr], wc);
                        atomicMax(&c.rhi[wr], wc + 1);
                    }
                }
            }
        }
        if (transposed && tid == 0) W[(size_t)arg * S + M] = 1.0;
    }
    csync(c);
    tick(c, PH_GATHER);

    // ---- elimination ---------------------------------------------------------
    int k = 0;
    const bool is_ctl = c.wm ? (tid == 0) : (tid == c.NWK);
    long long tb1 = (c.prof && tid == 0) ? clock64() : 0;
    for (;;) {
        if (c.wm || tid >= c.NWK) { // control warp: retire virgin-unit / empty columns
            // Pure bookkeeping, strictly sequential in general -- but a step whose unit
            // row already sits at position k (or whose column is empty) changes nothing,
            // so a run of such steps is checked 32 at a time against the current tables.
            const int cl = c.wm ? tid : tid - c.NWK;
            __syncwarp(); // the control lane's interchange of the previous step is visible to every lane
            for (;;) {
                const int kk = k + cl;
                bool noop = false;
                if (kk < M - 1) {
                    const int cn = c.cnt[kk];
                    noop = cn == 0 || (cn == 1 && c.unitRow[kk] == c.rowAt[kk]);
                }
                const unsigned m = __ballot_sync(kFull, noop);
                const int run = (m == kFull) ? 32 : __ffs(~m) - 1;
                k += run;
                if (run == 32) continue;
                if (k >= M - 1) break;
                // step k is not a no-op: a real virgin-unit interchange, or a non-trivial step
                int adv = 0;
                if (cl == 0 && c.cnt[k] == 1) {
                    const int ur = c.unitRow[k];
                    const int mu = c.posOf[ur];
                    if (mu >= k) { // its row is still active: no fill has reached this column
                        const int rk = c.rowAt[k];
                        c.rowAt[k] = ur;
                        c.rowAt[mu] = rk;
                        c.posOf[ur] = k;
                        c.posOf[rk] = mu;
                        adv = 1;
                    }
                }
                adv = __shfl_sync(kFull, adv, 0);
                if (!adv) break;
                ++k;
                __syncwarp();
            }
            if (cl == 0) {
                c.ctl[CTL_K] = k;
                c.ctl[CTL_NC] = 0;
            }
        }
        csync(c); // B1: updates of the previous step and the position tables are visible
        k = c.ctl[CTL_K];
        long long tq = 0;
        if (c.prof && tid == 0) {
            tq = clock64();
            c.prof[PH_E_B1] += tq - tb1;
        }
        if (k >= M - 1) break;
#if DZ_KERNEL_PER_NR
        if constexpr (NRX > 0) {
            warp_step_small<NRX>(c, W, k, is_ctl, y);
            ++k;
            continue;
        }
#endif
        if (c.wm && M <= 32 * WARP_NR_MAX) { // warp-per-LP fast path, NR = ceil(M/32)
            if (WARP_NR_MAX <= 4) {
                if (M <= 32)
                    warp_step_small<1>(c, W, k, is_ctl, y);
                else if (M <= 64)
                    warp_step_small<2>(c, W, k, is_ctl, y);
                else if (M <= 96)
                    warp_step_small<3>(c, W, k, is_ctl, y);
                else
                    warp_step_small<4>(c, W, k, is_ctl, y);
            } else {
                if (M <= 160)
                    warp_step_small<5>(c, W, k, is_ctl, y);
                else if (M <= 192)
                    warp_step_small<6>(c, W, k, is_ctl, y);
                else
                    warp_step_small<8>(c, W, k, is_ctl, y);
            }
            ++k;
            continue;
        }

        // Pivot search over the active rows of column k (linalg.rs:98-105): largest
        // |a_ik|, ties to the smallest logical position.  EVERY warp scans the whole
        // column and reduces it with REDUX ops on the bit pattern of |v| (monotone for
        // non-negative doubles), so no partial results cross warps.
        unsigned bhi = 0u, blo = 0u;
        int bidx = 0x7fffffff, brow = 0;
        // Small M: every warp scans the whole column (no exchange, one barrier less).
        // Large M: warps split the rows and exchange one partial each.
        const bool split = M > 32 * 12;
        const bool listed = split && c.cand_r != nullptr;
        {
            // four rows per thread in flight: positions first, then the predicated
            // column loads as one batch of independent loads
            const int rstart = split ? tid : lane, rstep = split ? c.nthreads : 32;
            for (int rb = rstart; rb < M; rb += 4 * rstep) {
                int pos4[4];
                double v4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = rb + q * rstep;
                    pos4[q] = (r < M) ? c.posOf[r] : -1;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = rb + q * rstep;
                    v4[q] = (pos4[q] >= k && in_iv(c, r, k)) ? W[(size_t)r * S + k] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = rb + q * rstep;
                    bool cand = pos4[q] >= k;
                    unsigned hi = (unsigned)__double2hiint(v4[q]) & 0x7fffffffu;
                    unsigned lo = (unsigned)__double2loint(v4[q]);
                    if (v4[q] != v4[q]) { // NaN never beats the incumbent, nor is it beaten as one
                        cand = cand && pos4[q] == k;
                        hi = 0xffffffffu;
                        lo = 0xffffffffu;
                    }
                    const bool better =
                        cand && (bidx == 0x7fffffff || hi > bhi ||
                                 (hi == bhi && (lo > blo || (lo == blo && pos4[q] < bidx))));
                    bhi = better ? hi : bhi;
                    blo = better ? lo : blo;
                    bidx = better ? pos4[q] : bidx;   // logical position (unique per row)
                    brow = better ? r : brow;
                    if (listed && pos4[q] >= k && v4[q] != 0.0) { // rows the update has to visit
                        const int e = atomicAdd(&c.ctl[CTL_NC], 1);
                        c.cand_r[e] = r;
                        c.cand_v[e] = v4[q];
                    }
                }
            }
        }
        unsigned mh = __reduce_max_sync(kFull, bhi);
        unsigned ml = __reduce_max_sync(kFull, bhi == mh ? blo : 0u);
        int gi = __reduce_min_sync(kFull, (bhi == mh && blo == ml) ? bidx : 0x7fffffff);
        // the lane that holds the winning position also holds its row
        int gr = __shfl_sync(kFull, brow, __ffs(__ballot_sync(kFull, bidx == gi)) - 1);
        if (split) {
            int *rp = c.red_idx + c.parity * 4 * kMaxWarps;
            if (lane == 0) {
                rp[warp] = (int)mh;
                rp[kMaxWarps + warp] = (int)ml;
                rp[2 * kMaxWarps + warp] = gi;
                rp[3 * kMaxWarps + warp] = gr;
            }
            csync(c);
            const bool have = lane < c.nwarps;
            bhi = have ? (unsigned)rp[lane] : 0u;
            blo = have ? (unsigned)rp[kMaxWarps + lane] : 0u;
            bidx = have ? rp[2 * kMaxWarps + lane] : 0x7fffffff;
            brow = have ? rp[3 * kMaxWarps + lane] : 0;
            if (bidx == 0x7fffffff) bhi = blo = 0u;
            mh = __reduce_max_sync(kFull, bhi);
            ml = __reduce_max_sync(kFull, bhi == mh ? blo : 0u);
            gi = __reduce_min_sync(kFull, (bhi == mh && blo == ml) ? bidx : 0x7fffffff);
            gr = __shfl_sync(kFull, brow, __ffs(__ballot_sync(kFull, bidx == gi)) - 1);
            c.parity ^= 1;
        }
        const int pr = gr, ppos = gi;
        const double pv = in_iv(c, pr, k) ? W[(size_t)pr * S + k] : 0.0;
        if (c.prof && tid == 0) {
            const long long t = clock64();
            c.prof[PH_E_SEARCH] += t - tq;
            tq = t;
        }
        csync(c); // B2: every warp has read the positions it needs for this search
        if (c.prof && tid == 0) {
            const long long t = clock64();
            c.prof[PH_E_B2] += t - tq;
            tq = t;
        }
        if (is_ctl) { // record the interchange k <-> ppos (linalg.rs:107-114)
            if (c.prof) c.prof[PH_NONTRIVIAL] += 1;
            const int rk = c.rowAt[k];
            c.rowAt[k] = pr;
            c.rowAt[ppos] = rk;
            c.posOf[pr] = k;
            c.posOf[rk] = ppos;
        }
        if (pv != 0.0 && warp < c.G) { // linalg.rs:117
            // Worker warp w owns rows r = lane*G + w.  Multipliers l_i = a_ik / pivot live
            // in registers; rows with an exact-zero entry are untouched by this step (the
            // pivot row is finite).  Then lanes switch to columns: for every owned row
            // with a multiplier, a_ij -= l_i * a_kj over the nonzero a_kj (linalg.rs:118-124),
            // four rows in flight.
            const int G = c.G;
            const double *__restrict__ prow = W + (size_t)pr * S;
            // columns of the pivot row that can be nonzero: (k, M) clipped to its
            // interval when intervals are tracked, plus the right-hand side column M
            int jlo = k + 1, jhi = M;
            if (c.rlo) {
                jlo = max(jlo, c.rlo[pr]);
                jhi = min(jhi, c.rhi[pr]);
            }
            const double urhs = prow[M];
            const int nslots = listed ? c.ctl[CTL_NC] : M;
            for (int rbase = 0; listed ? (rbase < nslots) : (rbase * G < M); rbase += listed ? 32 * G : 32) {
                // the rows this warp updates: from the search's candidate list when there
                // is one (no second scan of the column), else the rows it owns
                int r_own;
                bool need = false;
                double l = 0.0;
                if (listed) {
                    const int e = rbase + warp * 32 + lane;
                    r_own = (e < nslots) ? c.cand_r[e] : pr;
                    if (e < nslots && r_own != pr) {
                        need = true;
                        l = __ddiv_rn(c.cand_v[e], pv);
                    }
                } else {
                    r_own = (rbase + lane) * G + warp;
                    if (r_own < M && r_own != pr && c.posOf[r_own] >= k && in_iv(c, r_own, k)) {
                        const double v = W[(size_t)r_own * S + k];
                        if (v != 0.0) {
                            need = true;
                            l = __ddiv_rn(v, pv);
                        }
                    }
                }
                const unsigned rows = __ballot_sync(kFull, need);
                if (!rows) continue;
                unsigned long long upd = need ? 1 : 0;
                if (need) {
                    if (urhs != 0.0) { // the right-hand side rides along (forward substitution)
                        double *rp = W + (size_t)r_own * S + M;
                        *rp = __dsub_rn(*rp, __dmul_rn(l, urhs));
                        upd += 2;
                    }
                    if (c.rlo && jlo < jhi) { // the row may now be nonzero wherever the pivot row is
                        c.rlo[r_own] = min(c.rlo[r_own], jlo);
                        c.rhi[r_own] = max(c.rhi[r_own], jhi);
                    }
                }
                // the pivot row four chunks (128 columns) at a time, each batch of loads
                // independent; two rows of this warp in flight per batch
                for (int c0 = jlo; c0 < jhi; c0 += 128) {
                    double u[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int j = c0 + lane + 32 * cc;
                        u[cc] = (j < jhi) ? prow[j] : 0.0;
                    }
                    unsigned nzu = 0;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) nzu += (u[cc] != 0.0) ? 1u : 0u;
                    if (!__ballot_sync(kFull, nzu != 0)) continue;
                    unsigned rr = rows;
                    while (rr) {
                        const int b0 = __ffs(rr) - 1;
                        rr &= rr - 1;
                        const bool two = rr != 0;
                        const int b1 = two ? __ffs(rr) - 1 : b0;
                        if (two) rr &= rr - 1;
                        const double l0 = __shfl_sync(kFull, l, b0);
                        const double l1 = __shfl_sync(kFull, l, b1);
                        const int ra = __shfl_sync(kFull, r_own, b0), rb2 = __shfl_sync(kFull, r_own, b1);
                        double *__restrict__ w0 = W + (size_t)ra * S + c0 + lane;
                        double *__restrict__ w1 = W + (size_t)rb2 * S + c0 + lane;
                        double a0[4], a1[4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            a0[cc] = (u[cc] != 0.0) ? w0[32 * cc] : 0.0;
                            a1[cc] = (two && u[cc] != 0.0) ? w1[32 * cc] : 0.0;
                        }
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            if (u[cc] != 0.0) {
                                w0[32 * cc] = __dsub_rn(a0[cc], __dmul_rn(l0, u[cc]));
                                if (two) w1[32 * cc] = __dsub_rn(a1[cc], __dmul_rn(l1, u[cc]));
                            }
                        }
                        upd += 2ull * nzu * (two ? 2 : 1);
                    }
                }
                c.n_lu += upd;
            }
        }
        if (c.prof && tid == 0) {
            tb1 = clock64();
            c.prof[PH_E_UPD] += tb1 - tq;
        }
        ++k;
    }
    tick(c, PH_ELIM);
    if (c.prof && tid == 0) c.prof[PH_SOLVES] += 1;
    // ---- back substitution ------------------------------------------------------
    // (second trip only when a non-finite value appeared: redo without skipping)
    for (int literal = 0; literal < 2; ++literal) {
#if DZ_KERNEL_PER_NR
        if constexpr (NRX > 0)
            warp_back_substitute_small<NRX>(c, y, literal != 0);
        else
#endif
        if (c.wm && WARP_NR_MAX <= 4 && M <= 32)
            warp_back_substitute_small<1>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX <= 4 && M <= 64)
            warp_back_substitute_small<2>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX <= 4 && M <= 96)
            warp_back_substitute_small<3>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX <= 4 && M <= 128)
            warp_back_substitute_small<4>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX > 4 && M <= 160)
            warp_back_substitute_small<5>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX > 4 && M <= 192)
            warp_back_substitute_small<6>(c, y, literal != 0);
        else if (c.wm && WARP_NR_MAX > 4 && M <= 256)
            warp_back_substitute_small<8>(c, y, literal != 0);
        else
            back_substitute(c, y, literal != 0);
        if (literal || !c.ctl[CTL_FLAG]) break;
        csync(c);
        if (tid == 0) c.ctl[CTL_FLAG] = 0;
    }
}

// HOME: 0 = everything in shared memory, 1 = the working basis W in the HBM
// workspace, 2 = W and all per-LP vectors in the HBM workspace (large single LPs).
// WARP: one warp per LP (the CTA is just a bundle of independent warps, each with
// its own shared-memory slab, workspace slab and work-queue pulls; no CTA barrier
// is ever executed).  Otherwise one CTA per LP.
// NRMAX: largest ceil(m_int/32) the warp fast paths are instantiated for: 4 (64
// registers, 32 warps per SM) or 8 (128 registers, 16 warps per SM: config 5).
template <int HOME, bool WARP, int NRMAX, int NRX = 0>
__global__ void __launch_bounds__(WARP ? 128 : 1024, WARP ? (NRMAX <= 4 ? 8 : 4) : 1)
dz_batch_kernel(const TemplateDev T, const BatchDev Bt, const int smem_per_team) {
#ifdef DZ_EMU
    unsigned char *smem_raw = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    Ctx c;
    c.M = T.M;
    c.Nn = T.Nn;
    c.S = T.S; // (M + 1) | 1, read from the parameter bank wherever it is needed
    c.wm = WARP;
    c.tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    c.nthreads = WARP ? 32 : (int)blockDim.x;
    c.nwarps = WARP ? 1 : (int)(blockDim.x >> 5);
    c.G = WARP ? 1 : c.nwarps - 1;
    c.NWK = c.G * 32;
    c.parity = 0;
    const int M = c.M, Nn = c.Nn, tid = c.tid;
    const int team_in_cta = WARP ? (int)(threadIdx.x >> 5) : 0;
    const size_t team = WARP ? (size_t)blockIdx.x * (blockDim.x >> 5) + team_in_cta : (size_t)blockIdx.x;
    {
        double *sp = reinterpret_cast<double *>(smem_raw + (size_t)team_in_cta * smem_per_team);
        double *gp = Bt.gws + team * Bt.gws_stride;
        const size_t wsz = ((size_t)M * c.S + 1) & ~(size_t)1;
        if (HOME == 0) {
            c.W = sp;
            sp += wsz;
        } else {
            c.W = gp;
            gp += wsz;
        }
        // small fixed-size scratch always lives in shared memory (a single warp
        // needs no reduction scratch)
        if (!WARP) {
            c.red_key = sp, sp += 2 * 4 * kMaxWarps;
        } else {
            c.red_key = nullptr;
        }
        c.prof = Bt.prof ? reinterpret_cast<long long *>(sp) : nullptr;
        sp += PH_COUNT;
        int *isp = reinterpret_cast<int *>(sp);
        if (!WARP) {
            c.red_idx = isp, isp += 2 * 4 * kMaxWarps;
        } else {
            c.red_idx = nullptr;
        }
        c.ctl = isp, isp += 8;
        // vectors indexed by basis position stay in shared memory (HOME <= 1); the
        // nonbasic-side vectors, touched a handful of times per pivot, move to the
        // workspace in warp mode so that more warps fit per SM
        double *dp = (HOME == 2) ? gp : reinterpret_cast<double *>(isp);
        c.x = dp, dp += M;
        c.xb = dp, dp += M;
        c.dxv = dp, dp += M;
        c.vv = dp, dp += M;
        double *zp = WARP ? gp : dp;
        c.z = zp, zp += Nn;
        c.zb = zp, zp += Nn;
        c.dzv = zp, zp += Nn;
        int *zip = reinterpret_cast<int *>(zp);
        c.nb = zip, zip += Nn + (Nn & 1);
        if (WARP)
            gp = reinterpret_cast<double *>(zip);
        else
            dp = reinterpret_cast<double *>(zip);
        int *ip = reinterpret_cast<int *>(dp);
        c.bas = ip, ip += M;
        c.rowAt = ip, ip += M;
        c.posOf = ip, ip += M;
        c.cnt = ip, ip += M;
        c.unitRow = ip, ip += M;
        c.pend = ip, ip += M;
        c.pre = ip, ip += M + 2;
        if (HOME == 2) {
            c.rlo = ip, ip += M;
            c.rhi = ip, ip += M;
            c.plist = ip, ip += M;
            c.cand_r = ip, ip += M + (M & 1); // 11 M + 2 (+1) ints: the doubles below stay aligned
            c.pbuf = reinterpret_cast<double *>(ip);
            c.cand_v = c.pbuf + (M + 32 * kMaxWarps + 64);
            c.lval = c.cand_v + M;
        } else {
            c.rlo = c.rhi = c.plist = c.cand_r = nullptr;
            c.pbuf = c.lval = c.cand_v = nullptr;
        }
    }
    const long long max_pivots = Bt.max_pivots;

    for (;;) {
        csync(c);
        if (tid == 0) c.ctl[CTL_LP] = (int)atomicAdd(Bt.resume ? Bt.next_lp2 : Bt.next_lp, 1u);
        csync(c);
        const long long item = (unsigned)c.ctl[CTL_LP];
        // second launch behind the on-chip core kernel (dz_core.cu): the LPs it handed over,
        // continued from the state they were in
        if (item >= (Bt.resume ? (long long)*Bt.exo_count : Bt.B)) break;
        const long long lp = Bt.resume ? (long long)Bt.exo_list[item] : item;
        const double *__restrict__ theta = Bt.theta + (size_t)lp * Bt.n_theta;
        c.n_lu = c.n_solve = c.n_price = 0;
        unsigned long long n_upd = 0;
        if (c.prof) {
            if (tid < PH_COUNT) c.prof[tid] = 0;
            c.t_last = clock64();
        }

        if (c.lval) { // resolve the lowered values once: pricing and gathers then stream them
            for (int e = tid; e < T.nnz; e += c.nthreads) c.lval[e] = load_ref(theta, T.val_ref[e]);
        }
        if (c.rlo) { // interval mode: W is all-zero here (the host zeroes the workspace before
                     // the launch, every LP cleans up after itself); every interval empty
            for (int r = tid; r < M; r += c.nthreads) {
                c.rlo[r] = 0x7fffffff;
                c.rhi[r] = 0;
            }
        }
        // initial state (simplex.rs:190-205)
        for (int p = tid; p < M; p += c.nthreads) {
            c.bas[p] = T.basis0[p];
            c.x[p] = load_ref(theta, T.b_ref[p]);
            c.xb[p] = 1.0;
        }
        for (int k = tid; k < Nn; k += c.nthreads) {
            const int col = T.nonbasis0[k];
            c.nb[k] = col;
            c.z[k] = -load_ref(theta, T.c_ref[col]);
            c.zb[k] = 1.0;
        }
        csync(c);

        int status = DZ_OPTIMAL;
        long long pivots = 0, n_primal = 0;
        unsigned long long hash = 0xcbf29ce484222325ULL;
        if (Bt.resume) {
            const unsigned char *st = Bt.exo_state + (size_t)item * Bt.exo_stride;
            const double *sd = reinterpret_cast<const double *>(st);
            const long long *sl = reinterpret_cast<const long long *>(sd + 2 * M + 2 * Nn);
            const int *si = reinterpret_cast<const int *>(sl + 3);
            for (int p = tid; p < M; p += c.nthreads) {
                c.x[p] = sd[p];
                c.xb[p] = sd[M + p];
                c.bas[p] = si[p];
            }
            for (int k = tid; k < Nn; k += c.nthreads) {
                c.z[k] = sd[2 * M + k];
                c.zb[k] = sd[2 * M + Nn + k];
                c.nb[k] = si[M + k];
            }
            pivots = sl[0];
            n_primal = sl[1];
            hash = (unsigned long long)sl[2];
            csync(c);
        }
        if (M == 0) status = DZ_BREAKDOWN; // 0x0 basis: the reference panics (linalg.rs:95)

        while (M > 0) {
            // ---- status(), simplex.rs:274-306 ----
            int q0, p0;
            find_first_both(c, q0, p0);
            tick(c, PH_STATUS);
            bool primal_step;
            double mu;
            if (q0 >= 0 && p0 >= 0) {
                const double primal = __ddiv_rn(-c.x[p0], c.xb[p0]);
                const double dual = __ddiv_rn(-c.z[q0], c.zb[q0]);
                if (primal <= 1e-12 && dual <= 1e-12) break;
                if (primal < dual) {
                    primal_step = true;
                    mu = dual;
                } else {
                    primal_step = false;
                    mu = primal;
                }
            } else if (q0 >= 0) {
                primal_step = true;
                mu = __ddiv_rn(-c.z[q0], c.zb[q0]);
            } else if (p0 >= 0) {
                primal_step = false;
                mu = __ddiv_rn(-c.x[p0], c.xb[p0]);
            } else {
                status = DZ_BREAKDOWN; // "unexpected code path", simplex.rs:304
                break;
            }
            if (pivots >= max_pivots) {
                status = DZ_PIVOT_CAP;
                break;
            }
            // primal_step (simplex.rs:308-318): dx = B^-1 a_j, ratio test on x, then dz;
            // dual_step   (simplex.rs:320-330): dz first, ratio test on z, then dx.
            // One solve call site serves both orders.
            int p = p0, q = q0;
            bool failed = false;
            for (int pass = 0; pass < 2; ++pass) {
                const bool transposed = (pass == 0) != primal_step;
                basis_solve<NRMAX, NRX>(c, T, theta, transposed, transposed ? p : c.nb[q],
                                   transposed ? c.vv : c.dxv);
                if (transposed) {
                    // pricing: dz = -N^T v (simplex.rs:235, linalg.rs:199-207); each
                    // column is summed sequentially in ascending row order
                    for (int k = tid; k < Nn; k += c.nthreads) {
                        const int col = c.nb[k];
                        const int e0 = T.col_ptr[col], e1 = T.col_ptr[col + 1];
                        double s = 0.0;
                        unsigned cntp = 0;
#pragma unroll 4
                        for (int e = e0; e < e1; ++e) {
                            const double a = c.lval ? c.lval[e] : load_ref(theta, T.val_ref[e]);
                            const double t = __dadd_rn(s, __dmul_rn(a, -c.vv[T.row_idx[e]]));
                            const bool nz = (a != 0.0);
                            s = nz ? t : s;
                            cntp += nz ? 2u : 0u;
                        }
                        c.n_price += cntp;
                        c.dzv[k] = s;
                    }
                    csync(c);
                    tick(c, PH_PRICE);
                }
                if (pass == 0) {
                    if (primal_step) {
                        p = find_second(c, mu, c.x, c.xb, c.dxv, M);
                        if (p < 0) {
                            status = DZ_UNBOUNDED;
                            failed = true;
                        }
                    } else {
                        q = find_second(c, mu, c.z, c.zb, c.dzv, Nn);
                        if (q < 0) {
                            status = DZ_INFEASIBLE;
                            failed = true;
                        }
                    }
                    tick(c, PH_RATIO);
                    if (failed) break;
                }
            }
            if (failed) break;
            // ---- Simplex::pivot, simplex.rs:253-268 ----
            const int leaving = c.bas[p], entering = c.nb[q];
            double t, s, t_bar, s_bar;
            bool ok = true;
            {
                const double xp = c.x[p], dxp = c.dxv[p], zq = c.z[q], dzq = c.dzv[q];
                const double xbp = c.xb[p], zbq = c.zb[q];
                t = (xp == 0.0 && dxp == 0.0) ? 0.0 : __ddiv_rn(xp, dxp);
                s = (zq == 0.0 && dzq == 0.0) ? 0.0 : __ddiv_rn(zq, dzq);
                t_bar = (xbp == 0.0 && dxp == 0.0) ? 0.0 : __ddiv_rn(xbp, dxp);
                s_bar = (zbq == 0.0 && dzq == 0.0) ? 0.0 : __ddiv_rn(zbq, dzq);
                ok = isfinite(t) && isfinite(s) && isfinite(t_bar) && isfinite(s_bar);
            }
            if (!ok) {
                status = DZ_BREAKDOWN; // safe_divide assert, simplex.rs:466
                break;
            }
            csync(c); // everyone has read x[p], z[q], ... before they change
            for (int k = tid; k < M; k += c.nthreads) { // fn pivot, simplex.rs:410-421
                const double d = c.dxv[k];
                if (k == p) {
                    c.x[k] = t;
                    c.xb[k] = t_bar;
                } else {
                    c.x[k] = __dsub_rn(c.x[k], __dmul_rn(t, d));
                    c.xb[k] = __dsub_rn(c.xb[k], __dmul_rn(t_bar, d));
                }
            }
            for (int k = tid; k < Nn; k += c.nthreads) {
                const double d = c.dzv[k];
                if (k == q) {
                    c.z[k] = s;
                    c.zb[k] = s_bar;
                } else {
                    c.z[k] = __dsub_rn(c.z[k], __dmul_rn(s, d));
                    c.zb[k] = __dsub_rn(c.zb[k], __dmul_rn(s_bar, d));
                }
            }
            n_upd += 4ull * (M + Nn);
            if (tid == 0) { // swap, simplex.rs:239-251
                c.bas[p] = entering;
                c.nb[q] = leaving;
                if (Bt.trace && pivots < Bt.trace_cap) {
                    int *tr = Bt.trace + ((size_t)lp * Bt.trace_cap + pivots) * 3;
                    tr[0] = primal_step ? 0 : 1;
                    tr[1] = leaving;
                    tr[2] = entering;
                }
            }
            {
                const unsigned long long w = (unsigned long long)(primal_step ? 0u : 1u) |
                                             ((unsigned long long)(unsigned)leaving << 1) |
                                             ((unsigned long long)(unsigned)entering << 32);
                hash = (hash ^ w) * 0x100000001b3ULL;
            }
            ++pivots;
            if (primal_step) ++n_primal;
            csync(c);
            tick(c, PH_UPDATE);
        }

        // ---- results: objective_value / solution, simplex.rs:345-371 ----
        csync(c);
        if (c.rlo) zero_intervals(c); // leave W all-zero for the next LP of this team
        if (tid == 0) {
            double obj = 0.0;
            for (int p = 0; p < M; ++p)
                obj = __dadd_rn(obj, __dmul_rn(load_ref(theta, T.c_ref[c.bas[p]]), c.x[p]));
            obj = __dadd_rn(load_ref(theta, T.c0_ref), obj);
            Bt.status[lp] = status;
            Bt.pivots[lp] = (int)pivots;
            Bt.n_primal[lp] = (int)n_primal;
            Bt.trace_hash[lp] = hash;
            Bt.objective[lp] = obj;
        }
        if (Bt.x_basic)
            for (int p = tid; p < M; p += c.nthreads) Bt.x_basic[(size_t)lp * M + p] = c.x[p];
        if (Bt.basis)
            for (int p = tid; p < M; p += c.nthreads) Bt.basis[(size_t)lp * M + p] = c.bas[p];
        if (Bt.values) {
            for (int v = tid; v < T.n_orig; v += c.nthreads) {
                const int cp = T.pos_index[v], cn = T.neg_index[v];
                double pos = 0.0, neg = 0.0;
                for (int p = 0; p < M; ++p) {
                    const int col = c.bas[p];
                    if (col == cp) pos = c.x[p];
                    if (col == cn) neg = c.x[p];
                }
                Bt.values[(size_t)lp * T.n_orig + v] = __dsub_rn(pos, neg);
            }
        }
        if (c.prof) {
            csync(c);
            if (tid < PH_COUNT) Bt.prof[(size_t)lp * PH_COUNT + tid] = c.prof[tid];
        }
        if (Bt.work) {
            // per-LP executed flop counts (block sum via atomics on the output row)
            double *w = Bt.work + (size_t)lp * 8;
            if (c.n_lu) atomicAdd(&w[0], (double)c.n_lu);
            if (c.n_solve) atomicAdd(&w[1], (double)c.n_solve);
            if (c.n_price) atomicAdd(&w[2], (double)c.n_price);
            if (tid == 0) atomicAdd(&w[3], (double)n_upd);
        }
    }
}

size_t zvec_bytes_for(int Nn) { return 3 * (size_t)Nn * 8 + ((size_t)Nn + (Nn & 1)) * 4; }
size_t pvec_bytes_for(int M) { return 4 * (size_t)M * 8 + (7 * (size_t)M + 2) * 4 + 16; }
size_t iv_bytes_for(int M) { // intervals, pending list, product buffer (HOME == 2)
    return 2 * (size_t)M * 4 + 2 * ((size_t)M + 1) * 4 + (2 * (size_t)M + 32 * kMaxWarps + 64) * 8 + 16;
}
size_t vec_bytes_for(int M, int Nn) { return zvec_bytes_for(Nn) + pvec_bytes_for(M); }
size_t fixed_smem_bytes(bool warp) {
    return warp ? PH_COUNT * 8 + 8 * 4 + 16
                : (2 * 4 * kMaxWarps + PH_COUNT) * 8 + (2 * 4 * kMaxWarps + 8) * 4 + 16;
}
size_t w_bytes_for(int M) { return ((((size_t)M * (size_t)((M + 1) | 1)) + 1) & ~(size_t)1) * 8; }
size_t smem_bytes_for(int M, int Nn, int home) {
    return fixed_smem_bytes(false) + (home == 0 ? w_bytes_for(M) : 0) +
           (home <= 1 ? vec_bytes_for(M, Nn) : 0);
}

template <int HOME, bool WARP, int NRMAX, int NRX = 0>
cudaError_t launch_one(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan,
                       cudaStream_t st) {
    auto kern = dz_batch_kernel<HOME, WARP, NRMAX, NRX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         plan.smem_bytes);
    if (e != cudaSuccess) return e;
#ifdef DZ_EMU
    (void)st;
    return emu::launch(kern, plan.grid, plan.block, (size_t)plan.smem_bytes, T, Bt, plan.smem_per_team);
#else
    kern<<<plan.grid, plan.block, plan.smem_bytes, st>>>(T, Bt, plan.smem_per_team);
    return cudaGetLastError();
#endif
}

} // namespace

int plan_launch(int device, int32_t M, int32_t Nn, int64_t nnz, int64_t B, int32_t warps_hint,
                int32_t cps_hint, int32_t basis_home, LaunchPlan *plan, std::string *err) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        *err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    if (M >= (1 << 30)) {
        *err = "kernel supports m_int < 2^30";
        return DZ_ERR_LIMIT;
    }
    const size_t max_smem = prop.sharedMemPerBlockOptin;
    const size_t per_sm = prop.sharedMemPerMultiprocessor;
    const int sms = prop.multiProcessorCount;
    const size_t with_w = smem_bytes_for(M, Nn, 0);
    const size_t without = smem_bytes_for(M, Nn, 1);
    // Where the dense working basis lives.  The kernel is bound by dependent-issue
    // latency, so LPs in flight per SM is what buys throughput: with the basis in
    // shared memory only floor(227 KB / (8 M^2)) CTAs fit per SM; with it in an
    // HBM workspace up to 12 do.  Measured on config 2 (profiles/): 3.3 kLP/s
    // shared vs 5.1 kLP/s workspace.  Small batches (fewer LPs than the
    // shared-memory grid) keep the lower-latency shared home.  When even the
    // per-LP vectors do not fit (M in the thousands) everything moves to HBM.
    plan->warp_mode = false;
    plan->smem_per_team = 0;
    plan->core_mode = false;
    plan->grid_mode = false;
    // Whole-GPU single-LP kernel (dz_grid.cu): one cooperative grid, one CTA per SM, for the
    // few-and-large case (configs 3 and 4); basis_home == 5 forces it (tests use small grids
    // through ctas_per_sm = number of CTAs, worker_warps = warps per CTA).
    if (basis_home == 5 || (basis_home == 0 && warps_hint == 0 && M > 1024 && B <= 4)) {
        plan->grid_mode = true;
        plan->home = 5;
        plan->block = basis_home == 5 && warps_hint > 0 ? 32 * std::min(warps_hint, 16) : 512;
        plan->grid = basis_home == 5 && cps_hint > 0 ? cps_hint : sms;
        plan->worker_warps = plan->block / 32;
        plan->w_in_smem = false;
        plan->smem_bytes = (int32_t)grid_smem_bytes();
        plan->ctas_per_sm = 1;
        plan->teams = 1;
        plan->gws_doubles_per_cta = 0;
        return DZ_OK;
    }
    // On-chip coupled-core kernel (dz_core.cu): basis_home == 4.  Bit-identical, but measured
    // slower than the warp-per-LP shape on configs 2 and 5 (profiles/r02_experiments.md), so
    // it is not an automatic choice.
    {
        const int nq = core_nq(M);
        const size_t fixed = nq ? core_fixed_smem_bytes(M, Nn, nq) : 0;
        // ... except for small batches of the wide classes (129 <= m_int <= 256, config 5's
        // shape): with at most ~2.5 LPs per SM its low per-LP latency wins (measured: 296
        // config-5 LPs 146 LP/s against 121 LP/s; 592 LPs 167 against 211)
        const bool small_wide = basis_home == 0 && warps_hint == 0 && cps_hint == 0 && M > 128 && B > 1 &&
                                B * 2 <= (int64_t)sms * 5;
        const bool want = basis_home == 4 || small_wide;
        if (want && nq && M >= 1 && fixed + 1024 <= max_smem) {
            const size_t full = (size_t)M * (size_t)((M + 1) | 1) * 8;
            int cps = 1;
            if (cps_hint > 0) {
                cps = cps_hint;
            } else {
                // as many CTAs per SM as still keep >= 3/4 of a full-size core on chip; never
                // more CTAs than LPs
                for (int t = 8; t >= 1; --t) {
                    const size_t per_cta = per_sm / t - 1024;
                    if (per_cta > fixed && per_cta - fixed >= full * 3 / 4) {
                        cps = t;
                        break;
                    }
                }
                const int64_t need = (B + sms - 1) / sms;
                if (need < cps) cps = (int)std::max<int64_t>(need, 1);
                // wide classes: a second CTA per SM (256 threads each, part of the core's rows in
                // the workspace) beats a second wave: 296 config-5 LPs 1.64 s against 2.03 s
                if (nq > 4 && B > sms) cps = 2;
            }
            cps = std::max(1, std::min(cps, 16));
            size_t per_cta = std::min<size_t>(per_sm / cps - 1024, max_smem);
            if (per_cta < fixed + 64) per_cta = fixed + 64;
            size_t cap_bytes = std::min(per_cta - fixed, (full + 15) & ~(size_t)15);
            cap_bytes &= ~(size_t)15;
            plan->core_mode = true;
            plan->core_cap_w = (int32_t)(cap_bytes / 8);
            plan->home = 4;
            plan->worker_warps = nq <= 4 ? 3 : (cps >= 2 ? 7 : 15);
            plan->w_in_smem = cap_bytes >= full;
            plan->block = nq <= 4 ? 128 : (cps >= 2 ? 256 : 512);
            plan->smem_bytes = (int32_t)(fixed + cap_bytes);
            plan->ctas_per_sm = cps;
            plan->gws_doubles_per_cta = (int64_t)((full + 15) / 8);
            int64_t grid = (int64_t)sms * cps;
            if (grid > B) grid = B;
            plan->grid = (int32_t)std::max<int64_t>(grid, 1);
            plan->teams = plan->grid;
            return DZ_OK;
        }
    }
    // Auto: one warp per LP when the batch is large enough to fill the machine with
    // independent warps and the fast small-M step applies; measured on config 2:
    // 8.4 kLP/s (32 warps/SM) vs 5.3 kLP/s CTA-per-LP (profiles/).
    // ... and for the wide classes (config 5) once the batch fills its 16 warps per SM: measured
    // 4736 LPs 465 LP/s against 414 (CTA per LP); 1184 LPs 317 against 368, so smaller batches
    // stay with the CTA shape (and the smallest go to the core kernel above).
    const bool auto_warp = warps_hint == 0 && basis_home == 0 &&
                           ((M <= 128 && B >= (int64_t)sms * 8) || (M > 128 && M <= 256 && B >= (int64_t)sms * 12));
    if ((warps_hint < 0 || auto_warp) && without <= max_smem / 2 && M <= 1024) {
        // warp-per-LP: WPC independent warps per CTA, basis in the HBM workspace
        const size_t per_team =
            (fixed_smem_bytes(true) + pvec_bytes_for(M) + 15) & ~(size_t)15;
        int wpc = (int)std::min<size_t>(4, max_smem / per_team);
        wpc = std::max(1, wpc);
        plan->warp_mode = true;
        plan->home = 1;
        plan->worker_warps = 1;
        plan->w_in_smem = false;
        plan->block = 32 * wpc;
        plan->smem_per_team = (int32_t)per_team;
        plan->smem_bytes = (int32_t)(per_team * wpc);
        plan->gws_doubles_per_cta = (int64_t)(((w_bytes_for(M) + zvec_bytes_for(Nn) + 15) & ~(size_t)15) / 8);
        const int cps_max = std::max(1, std::min((int)(per_sm / ((size_t)plan->smem_bytes + 1024)),
                                                 2048 / plan->block));
        int cps = cps_max;
        if (cps_hint > 0) cps = std::min(cps_max, cps_hint);
        plan->ctas_per_sm = cps;
        int64_t grid = (int64_t)sms * cps;
        const int64_t need = (B + wpc - 1) / wpc;
        if (grid > need) grid = need;
        plan->grid = (int32_t)std::max<int64_t>(grid, 1);
        plan->teams = (int64_t)plan->grid * wpc;
        return DZ_OK;
    }
    const int cps_smem = with_w <= max_smem ? (int)(per_sm / (with_w + 1024)) : 0;
    int home;
    if (without > max_smem || basis_home == 3)
        home = 2;
    else if (basis_home == 1)
        home = cps_smem > 0 ? 0 : 1;
    else if (basis_home == 2)
        home = 1;
    else
        home = (cps_smem > 0 && (cps_smem >= 6 || B <= (int64_t)sms * cps_smem)) ? 0 : 1;
    int g = warps_hint > 0 ? warps_hint
                           : (home == 0 ? std::min(8, std::max(2, (M + 31) / 32 + 1))
                                        : (home == 2 ? std::min(31, std::max(2, (M + 31) / 32))
                                                     // measured on config 5 (M=192): 3 worker
                                                     // warps 405 LP/s, 2: 340, 4: 355, 6: 314
                                                     : std::min(8, std::max(3, (M + 63) / 64))));
    if (B <= sms && home != 0 && warps_hint <= 0) g = std::min(31, std::max(g, (M + 31) / 32));
    g = std::max(1, std::min(g, 31));
    plan->worker_warps = g;
    plan->home = home;
    plan->block = (g + 1) * 32;
    plan->w_in_smem = home == 0;
    plan->smem_bytes = (int32_t)smem_bytes_for(M, Nn, home);
    size_t ws = home == 0 ? 0
                          : w_bytes_for(M) + (home == 2 ? vec_bytes_for(M, Nn) + iv_bytes_for(M) +
                                                              (size_t)nnz * 8 + 16
                                                        : 0);
    plan->gws_doubles_per_cta = (int64_t)(((ws + 15) & ~(size_t)15) / 8);
    const int cps_max = std::max(1, std::min((int)(per_sm / ((size_t)plan->smem_bytes + 1024)),
                                             2048 / plan->block));
    int cps = std::min(cps_max, home == 0 ? 32 : 12);
    if (cps_hint > 0) cps = std::min(cps_max, cps_hint);
    plan->ctas_per_sm = cps;
    int64_t grid = (int64_t)sms * cps;
    if (grid > B) grid = B;
    if (grid < 1) grid = 1;
    plan->grid = (int32_t)grid;
    plan->teams = grid;
    return DZ_OK;
}

int launch_batch(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream,
                 std::string *err) {
    if (plan.core_mode) return launch_core(T, Bt, plan, stream, err);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
#if DZ_KERNEL_PER_NR
    if (plan.warp_mode && T.M <= 256) {
        switch ((T.M + 31) / 32) {
        case 0:
        case 1: e = launch_one<1, true, 4, 1>(T, Bt, plan, st); break;
        case 2: e = launch_one<1, true, 4, 2>(T, Bt, plan, st); break;
        case 3: e = launch_one<1, true, 4, 3>(T, Bt, plan, st); break;
        case 4: e = launch_one<1, true, 4, 4>(T, Bt, plan, st); break;
        case 5: e = launch_one<1, true, 8, 5>(T, Bt, plan, st); break;
        case 6: e = launch_one<1, true, 8, 6>(T, Bt, plan, st); break;
        default: e = launch_one<1, true, 8, 8>(T, Bt, plan, st); break;
        }
    } else
#endif
    if (plan.warp_mode)
        e = T.M <= 128 ? launch_one<1, true, 4>(T, Bt, plan, st) : launch_one<1, true, 8>(T, Bt, plan, st);
    else
        e = plan.home == 0 ? launch_one<0, false, 4>(T, Bt, plan, st)
            : plan.home == 1 ? launch_one<1, false, 4>(T, Bt, plan, st)
                             : launch_one<2, false, 4>(T, Bt, plan, st);
    if (e != cudaSuccess) {
        *err = std::string("dz_batch_kernel launch: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    return DZ_OK;
}

#ifdef DZ_EMU
int measure_fp64_peak(int, double *, double *, std::string *err) {
    *err = "not available on the SIMT emulator";
    return DZ_ERR_CUDA;
}
#else
// ---------------------------------------------------------------------------
// FP64 pipe micro-benchmark: the roofline denominator for the exact path.
// ---------------------------------------------------------------------------
namespace {
template <bool FUSED> __global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
    double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 1.0 - 1e-12, s = 1e-13;
    for (int i = 0; i < iters; ++i) {
        if (FUSED) {
            a0 = __fma_rn(a0, m, -s), a1 = __fma_rn(a1, m, -s), a2 = __fma_rn(a2, m, -s);
            a3 = __fma_rn(a3, m, -s), a4 = __fma_rn(a4, m, -s), a5 = __fma_rn(a5, m, -s);
            a6 = __fma_rn(a6, m, -s), a7 = __fma_rn(a7, m, -s);
        } else {
            a0 = __dsub_rn(__dmul_rn(a0, m), s), a1 = __dsub_rn(__dmul_rn(a1, m), s);
            a2 = __dsub_rn(__dmul_rn(a2, m), s), a3 = __dsub_rn(__dmul_rn(a3, m), s);
            a4 = __dsub_rn(__dmul_rn(a4, m), s), a5 = __dsub_rn(__dmul_rn(a5, m), s);
            a6 = __dsub_rn(__dmul_rn(a6, m), s), a7 = __dsub_rn(__dmul_rn(a7, m), s);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
} // namespace

int measure_fp64_peak(int device, double *mul_sub_gflops, double *fma_gflops, std::string *err) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        *err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int block = 512, grid = prop.multiProcessorCount * 4, iters = 20000;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)grid * block) != cudaSuccess) {
        *err = "cudaMalloc failed";
        return DZ_ERR_ALLOC;
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double res[2] = {0, 0};
    for (int fused = 0; fused < 2; ++fused) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            if (fused)
                fp64_peak_kernel<true><<<grid, block>>>(out, iters);
            else
                fp64_peak_kernel<false><<<grid, block>>>(out, iters);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < best) best = ms;
        }
        const double flop = 2.0 * 8.0 * (double)iters * (double)grid * block;
        res[fused] = flop / (best * 1e-3) / 1e9;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        *err = std::string("fp64 peak kernel: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    if (mul_sub_gflops) *mul_sub_gflops = res[0];
    if (fma_gflops) *fma_gflops = res[1];
    return DZ_OK;
}

#endif // DZ_EMU

} // namespace dz
