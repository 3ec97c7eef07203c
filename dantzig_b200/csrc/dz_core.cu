// dz_core.cu -- the on-chip "coupled core" kernel for batches of small LPs (sm_100a).
//
// One CTA solves one LP at a time (persistent, work queue) and runs the whole pivot
// loop of Simplex::solve (/root/reference/src/simplex.rs:274-343) like the general
// kernel of dz_kernel.cu does -- same arithmetic, same order, same bits -- but the
// working matrix of each lu_solve (linalg.rs:8-10) lives in SHARED MEMORY and holds
// only the part of the basis that is not trivially decoupled:
//
//   A row of the basis that no structural basis column touches has one nonzero, its
//   own slack; that slack column has one nonzero, in that row.  Such a (row, slack)
//   pair is a 1x1 block of B and of B^T: its elimination step is pure bookkeeping
//   (the interchange still moves positions, which decide later tie-breaks) and its
//   solution component is the right-hand side entry.  What is left -- the rows that
//   structural basis columns touch, and the columns that are structural or slacks of
//   such rows -- is a square "core" of nr <= m_int rows (config 2: 66 of 96 on
//   average), eliminated as a dense nr x (nr+1) array [W | rhs] with the reference's
//   partial pivoting (Matrix::factorize linalg.rs:88-128, forward half of LU::solve
//   :286-291 riding along as the last column) and back-substituted (:292-297).
//   Rows that do not fit the CTA's shared-memory budget overflow, row by row, into
//   an HBM workspace behind the same row pointer.
//
// The liberties taken are the ones of dz_kernel.cu (skipping operations with an
// exact-zero factor and a finite co-factor, retiring virgin unit columns as
// bookkeeping, tracking interchanges in a position table, x/1 == x), plus the 1x1
// blocks above.  They are no-ops only while every value involved is finite, so:
//   * a back-substitution that produces a non-finite value (division by a zero pivot
//     that factorize skipped, linalg.rs:117,296) is completed by the rule the
//     reference's literal arithmetic implies (0 * inf = NaN poisons every earlier row
//     that does not multiply that inf by a nonzero);
//   * anything rarer -- a non-finite pivot, pivot row or multiplier, a structurally
//     singular basis, a 1x1 block whose row was used up by a zero-pivot step -- makes
//     the CTA hand the LP, WITH its current state, to the general kernel, which
//     continues it from that pivot (BatchDev::exo_*).
//
// THREAD MAP.  Warp 0 is the control warp: it walks the bookkeeping steps, searches
// the pivot column (lanes on core rows, REDUX arg-max with the position tie-break)
// and records the interchange; then every warp computes the multipliers of the rows
// it owns (row = lane * NW + warp) and updates them, lanes on columns.  Two CTA
// barriers per non-trivial step.  Back-substitution rows are solved in order by the
// control warp from per-row nonzero bit masks kept during the elimination.

#include "dz_device.cuh"

#include <algorithm>
#include <cstdio>
#include <string>

namespace dz {

namespace {

enum { CC_K = 0, CC_C, CC_PR, CC_DONE, CC_EXOTIC, CC_LP, CC_FLAG, CC_TOTAL, CC_NR, CC_CURSOR, CC_WORDS = 16 };
constexpr int kCoreWarpBuf = 64; // staged products per warp in the back-substitution

// Slots of the optional per-LP cycle profile (BatchDev::prof, 16 per LP).
enum { CP_STATUS = 0, CP_LISTS, CP_GATHER, CP_ELIM, CP_BACK, CP_PRICE, CP_RATIO, CP_UPDATE, CP_REAL_STEPS,
       CP_SOLVES, CP_BOOK, CP_SEARCH, CP_STEP_UPD, CP_HANDED, CP_NR_SUM, CP_OVERFLOW_ROWS };

struct Core {
    int M, Nn, NT, NW, tid, lane, warp;
    long long *prof; // shared-memory accumulators, or null
    long long t_last;
    // per-LP state, shared memory
    double *x, *xb, *dxv, *vv, *z, *zb, *dzv;
    double *ycore; // [M] solution of the core system by core column
    double *pbuf;  // [16 * kCoreWarpBuf (>= M)] per-warp staging of a back-substitution row's ordered products
    double *pvs;   // [2] pivot value of the current step
    int *bas, *nb;
    int *rowAt, *posOf;   // [M] interchanges of the running elimination (rows of B or of B^T)
    int *rowcnt;          // [M] structural basis columns touching each constraint row
    int *srow;            // [M] position -> row of the slack that sits there, -1 if structural
    int *spos;            // [M] row -> position of its slack, -1 if the slack is nonbasic
    int *rmapR, *rlist;   // constraint row -> core index / back
    int *pmap, *plist;    // basis position -> core index / back
    int *cstart;          // [M+1] first flat CSC entry of each core position (scatter)
    int *pivr;            // [M] core column -> core row that is its pivot row
    int *done;            // [M] back-substitution: this core column's component is final
    unsigned *rmask;      // [M][NQ] columns of each core row that may be nonzero
    int *scan;            // [2 * kMaxWarps]
    int *ctl;             // [CC_WORDS]
    double *red_key;
    int *red_idx;
    int parity;
    // working core
    double *Ws;   // shared-memory part
    int capW;     // its capacity in doubles
    double *Wg;   // overflow rows (HBM workspace)
    int nr, S, rs;
    unsigned long long n_lu, n_solve, n_price;
};

__device__ __forceinline__ void ctick(Core &c, int slot) {
    if (c.prof && c.tid == 0) {
        const long long now = clock64();
        c.prof[slot] += now - c.t_last;
        c.t_last = now;
    }
}

__device__ __forceinline__ double *wrow(const Core &c, int i) {
    return i < c.rs ? c.Ws + i * c.S : c.Wg + (size_t)(i - c.rs) * c.S;
}

// find_first_pivot (simplex.rs:423-437) on both sides, one barrier; see dz_kernel.cu.
__device__ __forceinline__ void core_find_first_both(Core &c, int &q0, int &p0) {
    Cand<4> cd;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        cd.key[n] = 0.0;
        cd.idx[n] = -1;
    }
    for (int k = c.tid; k < c.Nn; k += c.NT) {
        const double yb = c.zb[k];
        if (yb > 0.0) {
            const double ratio = (yb == 1.0) ? -c.z[k] : __ddiv_rn(-c.z[k], yb);
            if (ratio == ratio && beats(ratio, k, cd.key[0], cd.idx[0])) {
                cd.key[0] = ratio;
                cd.idx[0] = k;
            }
            if (cd.idx[1] < 0) cd.idx[1] = k;
        }
    }
    for (int k = c.tid; k < c.M; k += c.NT) {
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
    block_argmax<4>(cd, c.red_key, c.red_idx, c.parity, c.NW, c.tid, false);
    q0 = cd.idx[0];
    if (cd.idx[1] >= 0) { // the reference's reduce keeps a NaN first element (simplex.rs:432-435)
        const double r = __ddiv_rn(-c.z[cd.idx[1]], c.zb[cd.idx[1]]);
        if (r != r) q0 = cd.idx[1];
    }
    p0 = cd.idx[2];
    if (cd.idx[3] >= 0) {
        const double r = __ddiv_rn(-c.x[cd.idx[3]], c.xb[cd.idx[3]]);
        if (r != r) p0 = cd.idx[3];
    }
}

// find_second_pivot (simplex.rs:439-461).
__device__ __forceinline__ int core_find_second(Core &c, double mu, const double *y, const double *yb,
                                                const double *dy, int len) {
    Cand<1> cd;
    cd.key[0] = 0.0;
    cd.idx[0] = -1;
    for (int k = c.tid; k < len; k += c.NT) {
        const double denom = __dadd_rn(y[k], __dmul_rn(mu, yb[k]));
        const double ratio = __ddiv_rn(dy[k], denom);
        if (ratio > 0.0 && beats(ratio, k, cd.key[0], cd.idx[0])) {
            cd.key[0] = ratio;
            cd.idx[0] = k;
        }
    }
    block_argmax<1>(cd, c.red_key, c.red_idx, c.parity, c.NW, c.tid, false);
    return cd.idx[0];
}

// The coupled core of the current basis: constraint rows touched by a structural basis
// column, and the basis positions that are structural or hold the slack of such a row.
// Also the flat entry offsets of the core positions for the scatter.  Once per pivot.
// Returns false when the two lists differ in length (a row without any basic entry).
__device__ __forceinline__ bool core_lists(Core &c, const TemplateDev &T) {
    const int M = c.M;
    const unsigned lt = (1u << c.lane) - 1u;
    int nrow = 0, ncol = 0;
    for (int base = 0; base < M; base += c.NT) {
        const int i = base + c.tid;
        bool pr = false, pc = false;
        if (i < M) {
            pr = c.rowcnt[i] > 0;
            const int sr = c.srow[i];
            pc = sr < 0 || c.rowcnt[sr] > 0;
        }
        const unsigned mr = __ballot_sync(kFull, pr), mc = __ballot_sync(kFull, pc);
        if (c.lane == 0) {
            c.scan[c.warp] = __popc(mr);
            c.scan[kMaxWarps + c.warp] = __popc(mc);
        }
        __syncthreads();
        int offr = nrow, offc = ncol;
        for (int w = 0; w < c.NW; ++w) {
            const int a = c.scan[w], b = c.scan[kMaxWarps + w];
            if (w < c.warp) {
                offr += a;
                offc += b;
            }
            nrow += a;
            ncol += b;
        }
        if (i < M) {
            const int ir = pr ? offr + __popc(mr & lt) : -1;
            c.rmapR[i] = ir;
            if (pr) c.rlist[ir] = i;
            const int ic = pc ? offc + __popc(mc & lt) : -1;
            c.pmap[i] = ic;
            if (pc) c.plist[ic] = i;
        }
        __syncthreads();
    }
    c.nr = nrow;
    if (nrow != ncol) return false;
    // exclusive scan of the core positions' column lengths (warp 0)
    if (c.warp == 0) {
        const int per = (nrow + 31) >> 5;
        const int b0 = c.lane * per;
        int sum = 0;
        for (int t = 0; t < per; ++t) {
            const int cc = b0 + t;
            if (cc < nrow) {
                const int col = c.bas[c.plist[cc]];
                sum += T.col_ptr[col + 1] - T.col_ptr[col];
            }
        }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, off);
            if (c.lane >= off) incl += t;
        }
        int run = incl - sum;
        for (int t = 0; t < per; ++t) {
            const int cc = b0 + t;
            if (cc < nrow) {
                const int col = c.bas[c.plist[cc]];
                c.cstart[cc] = run;
                run += T.col_ptr[col + 1] - T.col_ptr[col];
            }
        }
        if (c.lane == 31) c.cstart[nrow] = incl;
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ void swap_pos(Core &c, int k, int mu) {
    const int rk = c.rowAt[k], rm = c.rowAt[mu];
    c.rowAt[k] = rm;
    c.rowAt[mu] = rk;
    c.posOf[rm] = k;
    c.posOf[rk] = mu;
}

// lu_solve (linalg.rs:8-10) of B (transposed == false, rhs = column `arg` of A) or of
// B^T (transposed == true, rhs = e_arg); result in y[0..M).  Returns false when the
// solve has to be continued by the general kernel (see the header).
template <int NQ>
__device__ __forceinline__ bool core_solve(Core &c, const TemplateDev &T, const double *__restrict__ theta,
                                           const bool transposed, const int arg, double *y) {
    const int M = c.M, nr = c.nr, tid = c.tid, lane = c.lane, warp = c.warp, NW = c.NW;
    const int S = (nr + 1) | 1;
    c.S = S;
    c.rs = min(nr, c.capW / S);
    // rows / columns of the system being solved, in the reference's index spaces
    const int *__restrict__ rmap = transposed ? c.pmap : c.rmapR;   // system row -> core row
    const int *__restrict__ cmap = transposed ? c.rmapR : c.pmap;   // system column -> core column
    const int *__restrict__ rlist = transposed ? c.plist : c.rlist; // core row -> system row
    const int *__restrict__ clist = transposed ? c.rlist : c.plist; // core column -> system column

    // ---- gather: core of B or B^T into [W | rhs], 1x1 blocks straight into y ------------
    {
        const int ns = c.rs * S, ng = (nr - c.rs) * S;
        for (int e = tid; e < ns; e += c.NT) c.Ws[e] = 0.0;
        for (int e = tid; e < ng; e += c.NT) c.Wg[e] = 0.0;
        for (int e = tid; e < nr * NQ; e += c.NT) c.rmask[e] = 0u;
        for (int i = tid; i < M; i += c.NT) {
            c.rowAt[i] = i;
            c.posOf[i] = i;
            c.pivr[i] = -1;
            c.done[i] = 0;
            y[i] = 0.0;
        }
        if (tid < CC_WORDS && tid != CC_LP) c.ctl[tid] = 0;
    }
    __syncthreads();
    {
        const int total = c.cstart[nr];
        for (int base = tid; base < total; base += 4 * c.NT) {
            int cc4[4], ee[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = base + q * c.NT;
                cc4[q] = -1;
                ee[q] = 0;
                if (idx < total) {
                    int lo = 0, hi = nr - 1;
                    while (lo < hi) { // largest cc with cstart[cc] <= idx
                        const int mid = (lo + hi + 1) >> 1;
                        if (c.cstart[mid] <= idx)
                            lo = mid;
                        else
                            hi = mid - 1;
                    }
                    cc4[q] = lo;
                    ee[q] = T.col_ptr[c.bas[c.plist[lo]]] + (idx - c.cstart[lo]);
                }
            }
            int ref[4], row[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ref[q] = cc4[q] >= 0 ? T.val_ref[ee[q]] : -1;
                row[q] = cc4[q] >= 0 ? T.row_idx[ee[q]] : 0;
            }
            double val[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) val[q] = load_ref(theta, ref[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (cc4[q] < 0 || val[q] == 0.0) continue; // exact zeros are not stored (linalg.rs:261)
                const int ir = c.rmapR[row[q]];            // a row a core column touches is a core row
                const int wi = transposed ? cc4[q] : ir, wj = transposed ? ir : cc4[q];
                wrow(c, wi)[wj] = val[q];
                atomicOr(&c.rmask[wi * NQ + (wj >> 5)], 1u << (wj & 31));
            }
        }
        if (transposed) {
            if (tid == 0) {
                const int ci = c.pmap[arg];
                if (ci >= 0)
                    wrow(c, ci)[nr] = 1.0;
                else
                    y[c.srow[arg]] = 1.0; // 1x1 block: v_r = 1 / 1
            }
        } else {
            const int e0 = T.col_ptr[arg], e1 = T.col_ptr[arg + 1];
            for (int e = e0 + tid; e < e1; e += c.NT) {
                const double val = load_ref(theta, T.val_ref[e]);
                if (val == 0.0) continue;
                const int r = T.row_idx[e];
                const int ir = c.rmapR[r];
                if (ir >= 0)
                    wrow(c, ir)[nr] = val;
                else
                    y[c.spos[r]] = val; // 1x1 block: dx_p = a_r / 1
            }
        }
    }
    __syncthreads();
    ctick(c, CP_GATHER);
    if (c.prof && tid == 0) {
        c.prof[CP_SOLVES] += 1;
        c.prof[CP_NR_SUM] += nr;
        c.prof[CP_OVERFLOW_ROWS] += nr - c.rs;
    }

    // ---- elimination ------------------------------------------------------------------------
    int k = 0;
    for (;;) {
        long long tw = (c.prof && tid == 0) ? clock64() : 0;
        if (warp == 0) {
            // bookkeeping steps: 1x1 blocks and virgin unit columns of the core.  A step whose
            // unit row already sits at position k changes nothing, so a run of such steps is
            // checked 32 at a time against the current tables.
            int irregular = 0;
            for (;;) {
                const int kk = k + lane;
                bool noop = false;
                if (kk < M - 1) {
                    const int cc = cmap[kk];
                    const int u = transposed ? (cc < 0 ? c.spos[kk] : -1) : c.srow[kk];
                    noop = u >= 0 && c.rowAt[kk] == u;
                    if (noop && cc >= 0) c.pivr[cc] = rmap[u];
                }
                const unsigned m = __ballot_sync(kFull, noop);
                const int run = (m == kFull) ? 32 : __ffs(~m) - 1;
                // lanes past the run may have recorded a pivot row for a step that is not
                // reached yet: harmless, the step itself records it again when it comes
                k += run;
                if (run == 32) continue;
                if (k >= M - 1) break;
                int adv = 0;
                if (lane == 0) {
                    const int cc = cmap[k];
                    const int u = transposed ? (cc < 0 ? c.spos[k] : -1) : c.srow[k];
                    if (cc < 0) {
                        const int pu = u >= 0 ? c.posOf[u] : -1;
                        if (pu < k) {
                            irregular = 1; // the block's row was used up by a zero-pivot step
                        } else {
                            swap_pos(c, k, pu);
                            adv = 1;
                        }
                    } else if (u >= 0 && c.posOf[u] >= k) {
                        swap_pos(c, k, c.posOf[u]);
                        c.pivr[cc] = rmap[u];
                        adv = 1;
                    }
                }
                adv = __shfl_sync(kFull, adv, 0);
                irregular = __shfl_sync(kFull, irregular, 0);
                __syncwarp();
                if (!adv) break;
                ++k;
            }
            if (c.prof && tid == 0) {
                const long long t = clock64();
                c.prof[CP_BOOK] += t - tw;
                tw = t;
            }
            if (irregular) {
                if (lane == 0) c.ctl[CC_EXOTIC] = 1;
            } else if (k >= M - 1) {
                if (lane == 0) {
                    c.ctl[CC_DONE] = 1;
                    // the row left at the last position is the pivot row of the last column
                    const int cc = cmap[M - 1], r = c.rowAt[M - 1];
                    if (cc >= 0) {
                        if (rmap[r] < 0)
                            c.ctl[CC_EXOTIC] = 1;
                        else
                            c.pivr[cc] = rmap[r];
                    } else {
                        const int u = transposed ? c.spos[M - 1] : c.srow[M - 1];
                        if (u != r) c.ctl[CC_EXOTIC] = 1;
                    }
                }
            } else {
                // pivot search in core column cc over the rows at positions >= k
                // (linalg.rs:98-105): largest |a_ik|, ties to the smallest position
                const int cc = cmap[k];
                unsigned bhi = 0u, blo = 0u;
                int bidx = 0x7fffffff;
                bool bad = false;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int i = lane + 32 * q;
                    if (i < nr) {
                        const int pos = c.posOf[rlist[i]];
                        if (pos >= k) {
                            const double v = wrow(c, i)[cc];
                            bad = bad || !isfinite(v);
                            const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu;
                            const unsigned lo = (unsigned)__double2loint(v);
                            const int packed = (pos << 16) | i;
                            const bool better = (hi | lo) != 0u &&
                                                (hi > bhi || (hi == bhi && (lo > blo || (lo == blo && packed < bidx))));
                            bhi = better ? hi : bhi;
                            blo = better ? lo : blo;
                            bidx = better ? packed : bidx;
                        }
                    }
                }
                const unsigned mh = __reduce_max_sync(kFull, bhi);
                const unsigned ml = __reduce_max_sync(kFull, bhi == mh ? blo : 0u);
                const int gi = __reduce_min_sync(kFull, (bhi == mh && blo == ml) ? bidx : 0x7fffffff);
                const bool anybad = __ballot_sync(kFull, bad) != 0u;
                if (lane == 0) {
                    c.ctl[CC_K] = k;
                    c.ctl[CC_C] = cc;
                    if (anybad) {
                        c.ctl[CC_EXOTIC] = 1;
                    } else if ((mh | ml) == 0u) {
                        // no nonzero candidate: the incumbent at position k stays, pivot 0, the
                        // step is skipped (linalg.rs:117)
                        const int ir = rmap[c.rowAt[k]];
                        if (ir < 0) {
                            c.ctl[CC_EXOTIC] = 1;
                        } else {
                            c.pivr[cc] = ir;
                            c.ctl[CC_PR] = ir;
                            c.pvs[0] = 0.0;
                        }
                    } else {
                        const int pr = gi & 0xffff, ppos = gi >> 16;
                        swap_pos(c, k, ppos); // linalg.rs:107-114
                        c.pivr[cc] = pr;
                        c.ctl[CC_PR] = pr;
                        c.pvs[0] = wrow(c, pr)[cc];
                    }
                }
            }
        }
        __syncthreads(); // A: the step (or the end) is published, positions are up to date
        if (c.prof && tid == 0) {
            const long long t = clock64();
            c.prof[CP_SEARCH] += t - tw;
            tw = t;
        }
        if (c.ctl[CC_EXOTIC]) return false;
        if (c.ctl[CC_DONE]) break;
        if (c.prof && tid == 0) c.prof[CP_REAL_STEPS] += 1;
        k = c.ctl[CC_K];
        const double pv = c.pvs[0];
        if (pv != 0.0) {
            const int cc = c.ctl[CC_C], pr = c.ctl[CC_PR];
            const double *__restrict__ prow = wrow(c, pr);
            // the pivot row right of the pivot, lanes on absolute core columns; column nr is the rhs
            const int q0 = (cc + 1) >> 5, qn = nr >> 5;
            double u[NQ + 1];
            unsigned cm = 0u, nzu = 0u;
            bool bad = false;
#pragma unroll
            for (int q = 0; q <= NQ; ++q) {
                const int j = 32 * q + lane;
                u[q] = (q >= q0 && q <= qn && j > cc && j <= nr) ? prow[j] : 0.0;
                bad = bad || !isfinite(u[q]);
                const unsigned bm = __ballot_sync(kFull, u[q] != 0.0);
                cm |= bm ? (1u << q) : 0u;
                nzu += __popc(bm);
            }
            // multipliers of the rows this warp owns (linalg.rs:119-120)
            const int i_own = lane * NW + warp;
            double l = 0.0;
            if (i_own < nr && i_own != pr && c.posOf[rlist[i_own]] > k) {
                const double v = wrow(c, i_own)[cc];
                if (v != 0.0) l = (pv == 1.0) ? v : ((pv == -1.0) ? -v : __ddiv_rn(v, pv));
                bad = bad || !isfinite(l);
            }
            if (__ballot_sync(kFull, bad)) {
                if (lane == 0) c.ctl[CC_EXOTIC] = 1;
            } else {
                unsigned rows = __ballot_sync(kFull, l != 0.0);
                // pattern of the pivot row right of cc, merged into every updated row
                unsigned pm = 0u;
                if (lane < NQ) {
                    pm = c.rmask[pr * NQ + lane];
                    const int lo = cc + 1 - 32 * lane; // bits >= lo
                    pm = lo <= 0 ? pm : (lo >= 32 ? 0u : (pm & ~((1u << lo) - 1u)));
                }
                if (lane == 0) c.n_lu += (unsigned long long)__popc(rows) * (1ull + 2ull * nzu);
                while (rows) {
                    const int b = __ffs(rows) - 1;
                    rows &= rows - 1;
                    const double lb = __shfl_sync(kFull, l, b);
                    const int i = b * NW + warp;
                    double *__restrict__ row = wrow(c, i);
                    if (lane < NQ && pm) c.rmask[i * NQ + lane] |= pm;
                    double a[NQ + 1];
#pragma unroll
                    for (int q = 0; q <= NQ; ++q)
                        if ((cm >> q) & 1u) {
                            const int j = 32 * q + lane;
                            a[q] = (j > cc && j <= nr) ? row[j] : 0.0;
                        }
#pragma unroll
                    for (int q = 0; q <= NQ; ++q)
                        if ((cm >> q) & 1u) {
                            const int j = 32 * q + lane;
                            if (j > cc && j <= nr) row[j] = __dsub_rn(a[q], __dmul_rn(lb, u[q])); // linalg.rs:121-123
                        }
                }
            }
        }
        __syncthreads(); // B: the updates of this step are visible
        if (c.prof && tid == 0) c.prof[CP_STEP_UPD] += clock64() - tw;
        if (c.ctl[CC_EXOTIC]) return false;
        ++k;
    }
    ctick(c, CP_ELIM);

    // ---- back substitution (linalg.rs:292-297), core columns nr-1 .. 0 -----------------------
    // LEVEL-SCHEDULED by data flow: every warp takes core columns in descending order from a
    // shared cursor and waits, 32 pattern columns at a time, for the components its row needs
    // (done[j] in shared memory); rows that do not depend on each other proceed side by side.
    // A row's products u_ij * y_j are formed by the lanes, staged in column order in the
    // warp's slice of pbuf and subtracted one after the other (ascending j, linalg.rs:294).
    {
        const unsigned lt = (1u << lane) - 1u;
        double *buf = c.pbuf + warp * kCoreWarpBuf;
        unsigned long long ops = 0;
        for (;;) {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(&c.ctl[CC_CURSOR], 1);
            idx = __shfl_sync(kFull, idx, 0);
            if (idx >= nr) break;
            const int cc = nr - 1 - idx;
            const int i = c.pivr[cc];
            const double *__restrict__ row = wrow(c, i);
            double s = row[nr];
            const double d = row[cc];
            int cnt = 0;
            const int q0 = (cc + 1) >> 5;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (q < q0) continue;
                unsigned bits = c.rmask[i * NQ + q];
                const int lo = cc + 1 - 32 * q;
                bits = lo <= 0 ? bits : (lo >= 32 ? 0u : (bits & ~((1u << lo) - 1u)));
                if (!bits) continue; // warp-uniform
                const int j = 32 * q + lane;
                const bool has = (bits >> lane) & 1u;
                for (;;) { // wait for the components these columns need
                    const bool ok = !has || *((volatile int *)&c.done[j]) != 0;
                    if (__ballot_sync(kFull, !ok) == 0u) break;
#ifdef DZ_EMU
                    emu::yield();
#endif
                }
                double uu = 0.0, yj = 0.0;
                if (has) {
                    uu = row[j];
                    yj = *((volatile double *)&c.ycore[j]);
                }
                const bool take = uu != 0.0 && yj != 0.0;
                const unsigned mk = __ballot_sync(kFull, take);
                if (cnt + __popc(mk) > kCoreWarpBuf) { // flush the staged products, in order
                    __syncwarp();
                    for (int t = 0; t < cnt; ++t) s = __dsub_rn(s, buf[t]);
                    ops += 2ull * cnt;
                    cnt = 0;
                    __syncwarp();
                }
                if (take) buf[cnt + __popc(mk & lt)] = __dmul_rn(uu, yj);
                cnt += __popc(mk);
            }
            __syncwarp();
            int t = 0;
            for (; t + 4 <= cnt; t += 4) { // ascending column order (linalg.rs:294)
                const double p0 = buf[t], p1 = buf[t + 1], p2 = buf[t + 2], p3 = buf[t + 3];
                s = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(s, p0), p1), p2), p3);
            }
            for (; t < cnt; ++t) s = __dsub_rn(s, buf[t]);
            const double yi = (d == 1.0) ? s : __ddiv_rn(s, d);
            if (lane == 0) {
                c.ycore[cc] = yi;
                __threadfence_block();
                *((volatile int *)&c.done[cc]) = 1;
                if (!isfinite(yi)) c.ctl[CC_FLAG] = 1;
            }
            ops += 2ull * cnt + 1ull;
            __syncwarp();
        }
        if (lane == 0) c.n_solve += ops;
    }
    __syncthreads();
    for (int cc = tid; cc < nr; cc += c.NT) y[clist[cc]] = c.ycore[cc];
    __syncthreads();
    ctick(c, CP_BACK);
    if (c.ctl[CC_FLAG]) {
        // A non-finite component (division by a pivot that factorize skipped).  The literal
        // arithmetic multiplies it into EVERY earlier row: by an exact zero wherever the row
        // has no entry in that column, which gives NaN.  So, walking the system columns
        // downwards: once a NaN has appeared everything below is NaN; a row below an
        // infinity keeps its value (already computed with that infinity) only if its entry
        // in that column is nonzero.
        if (tid == 0) {
            bool any_nan = false;
            int n_inf = 0; // core columns whose component is +-inf, kept in pbuf's space as ints
            int *inf_list = reinterpret_cast<int *>(c.pbuf);
            for (int kcol = M - 1; kcol >= 0; --kcol) {
                const int cc = cmap[kcol];
                double v = y[kcol];
                if (any_nan) {
                    v = __longlong_as_double(0x7ff8000000000000LL);
                } else if (n_inf > 0) {
                    bool all_nz = cc >= 0;
                    if (cc >= 0) {
                        const double *row = wrow(c, c.pivr[cc]);
                        for (int t = 0; t < n_inf && all_nz; ++t) all_nz = row[inf_list[t]] != 0.0;
                    }
                    if (!all_nz) v = __longlong_as_double(0x7ff8000000000000LL);
                }
                y[kcol] = v;
                if (v != v)
                    any_nan = true;
                else if (!isfinite(v) && cc >= 0)
                    inf_list[n_inf++] = cc;
                else if (!isfinite(v))
                    any_nan = true; // cannot happen: 1x1 components are finite inputs
            }
        }
        __syncthreads();
    }
    return true;
}

// State handed to the general kernel when this one gives an LP up (see the header).
__device__ __forceinline__ void core_hand_over(Core &c, const BatchDev &Bt, long long lp, long long pivots,
                                               long long n_primal, unsigned long long hash) {
    __syncthreads();
    if (c.tid == 0) c.ctl[CC_TOTAL] = (int)atomicAdd(Bt.exo_count, 1u);
    __syncthreads();
    const int slot = c.ctl[CC_TOTAL];
    if (c.tid == 0) Bt.exo_list[slot] = (int)lp;
    unsigned char *st = Bt.exo_state + (size_t)slot * Bt.exo_stride;
    double *sd = reinterpret_cast<double *>(st);
    const int M = c.M, Nn = c.Nn;
    for (int i = c.tid; i < M; i += c.NT) {
        sd[i] = c.x[i];
        sd[M + i] = c.xb[i];
    }
    for (int i = c.tid; i < Nn; i += c.NT) {
        sd[2 * M + i] = c.z[i];
        sd[2 * M + Nn + i] = c.zb[i];
    }
    long long *sl = reinterpret_cast<long long *>(sd + 2 * M + 2 * Nn);
    if (c.tid == 0) {
        sl[0] = pivots;
        sl[1] = n_primal;
        sl[2] = (long long)hash;
    }
    int *si = reinterpret_cast<int *>(sl + 3);
    for (int i = c.tid; i < M; i += c.NT) si[i] = c.bas[i];
    for (int i = c.tid; i < Nn; i += c.NT) si[M + i] = c.nb[i];
}

// NTH threads per CTA: 128 for m_int <= 128 (3-4 CTAs per SM); for the wide classes 512 (one CTA
// per SM, the whole core in shared memory) or 256 (two CTAs per SM, part of the core's rows in the
// HBM/L2 workspace).
template <int NQ, int NTH>
__global__ void __launch_bounds__(NTH, NTH == 128 ? 3 : (NTH == 256 ? 2 : 1))
dz_core_kernel(const TemplateDev T, const BatchDev Bt, const int capW) {
#ifdef DZ_EMU
    unsigned char *smem_raw = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    Core c;
    c.M = T.M;
    c.Nn = T.Nn;
    c.NT = (int)blockDim.x;
    c.NW = c.NT >> 5;
    c.tid = (int)threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.parity = 0;
    const int M = c.M, Nn = c.Nn, tid = c.tid;
    {
        double *dp = reinterpret_cast<double *>(smem_raw);
        c.Ws = dp, dp += capW;
        c.capW = capW;
        c.x = dp, dp += M;
        c.xb = dp, dp += M;
        c.dxv = dp, dp += M;
        c.vv = dp, dp += M;
        c.ycore = dp, dp += M;
        c.pbuf = dp, dp += (16 * kCoreWarpBuf > M ? 16 * kCoreWarpBuf : M);
        c.z = dp, dp += Nn;
        c.zb = dp, dp += Nn;
        c.dzv = dp, dp += Nn;
        c.pvs = dp, dp += 2;
        c.prof = Bt.prof ? reinterpret_cast<long long *>(dp) : nullptr;
        dp += 16;
        c.red_key = dp, dp += 2 * 4 * kMaxWarps;
        int *ip = reinterpret_cast<int *>(dp);
        c.red_idx = ip, ip += 2 * 4 * kMaxWarps;
        c.scan = ip, ip += 2 * kMaxWarps;
        c.ctl = ip, ip += CC_WORDS;
        c.bas = ip, ip += M;
        c.rowAt = ip, ip += M;
        c.posOf = ip, ip += M;
        c.rowcnt = ip, ip += M;
        c.srow = ip, ip += M;
        c.spos = ip, ip += M;
        c.rmapR = ip, ip += M;
        c.rlist = ip, ip += M;
        c.pmap = ip, ip += M;
        c.plist = ip, ip += M;
        c.pivr = ip, ip += M;
        c.done = ip, ip += M;
        c.cstart = ip, ip += M + 1;
        c.nb = ip, ip += Nn;
        c.rmask = reinterpret_cast<unsigned *>(ip), ip += M * NQ;
        c.Wg = Bt.gws + (size_t)blockIdx.x * Bt.gws_stride;
    }
    const long long max_pivots = Bt.max_pivots;

    for (;;) {
        __syncthreads();
        if (tid == 0) c.ctl[CC_LP] = (int)atomicAdd(Bt.next_lp, 1u);
        __syncthreads();
        const long long lp = (unsigned)c.ctl[CC_LP];
        if (lp >= Bt.B) break;
        const double *__restrict__ theta = Bt.theta + (size_t)lp * Bt.n_theta;
        c.n_lu = c.n_solve = c.n_price = 0;
        unsigned long long n_upd = 0;
        if (c.prof) {
            if (tid < 16) c.prof[tid] = 0;
            c.t_last = clock64();
        }

        // initial state (simplex.rs:190-205): the slacks are basic, in row order
        for (int p = tid; p < M; p += c.NT) {
            const int col = T.basis0[p];
            c.bas[p] = col;
            c.x[p] = load_ref(theta, T.b_ref[p]);
            c.xb[p] = 1.0;
            c.rowcnt[p] = 0;
            c.spos[p] = -1;
        }
        for (int k = tid; k < Nn; k += c.NT) {
            const int col = T.nonbasis0[k];
            c.nb[k] = col;
            c.z[k] = -load_ref(theta, T.c_ref[col]);
            c.zb[k] = 1.0;
        }
        __syncthreads();
        for (int p = tid; p < M; p += c.NT) {
            const int col = c.bas[p];
            const int sr = T.slack_row[col];
            c.srow[p] = sr;
            if (sr >= 0) {
                c.spos[sr] = p;
            } else {
                for (int e = T.col_ptr[col]; e < T.col_ptr[col + 1]; ++e) atomicAdd(&c.rowcnt[T.row_idx[e]], 1);
            }
        }
        __syncthreads();

        int status = DZ_OPTIMAL;
        long long pivots = 0, n_primal = 0;
        unsigned long long hash = 0xcbf29ce484222325ULL;
        bool handed_over = false;

        while (true) {
            // ---- status(), simplex.rs:274-306 ----
            int q0, p0;
            core_find_first_both(c, q0, p0);
            ctick(c, CP_STATUS);
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
            if (!core_lists(c, T)) {
                handed_over = true;
                break;
            }
            ctick(c, CP_LISTS);
            // primal_step (simplex.rs:308-318): dx = B^-1 a_j, ratio test on x, then dz;
            // dual_step   (simplex.rs:320-330): dz first, ratio test on z, then dx.
            int p = p0, q = q0;
            bool failed = false;
            for (int pass = 0; pass < 2 && !handed_over; ++pass) {
                const bool transposed = (pass == 0) != primal_step;
                if (!core_solve<NQ>(c, T, theta, transposed, transposed ? p : c.nb[q], transposed ? c.vv : c.dxv)) {
                    handed_over = true;
                    break;
                }
                if (transposed) {
                    // pricing: dz = -N^T v (simplex.rs:235, linalg.rs:199-207); each column is
                    // summed sequentially in ascending row order
                    for (int k = tid; k < Nn; k += c.NT) {
                        const int col = c.nb[k];
                        const int e0 = T.col_ptr[col], e1 = T.col_ptr[col + 1];
                        double s = 0.0;
                        unsigned cntp = 0;
#pragma unroll 4
                        for (int e = e0; e < e1; ++e) {
                            const double vr = c.vv[T.row_idx[e]];
                            if (vr == 0.0) continue; // a * -0 adds a zero (a is finite)
                            const double a = load_ref(theta, T.val_ref[e]);
                            if (a != 0.0) {
                                s = __dadd_rn(s, __dmul_rn(a, -vr));
                                cntp += 2u;
                            }
                        }
                        c.n_price += cntp;
                        c.dzv[k] = s;
                    }
                    __syncthreads();
                    ctick(c, CP_PRICE);
                }
                if (pass == 0) {
                    if (primal_step) {
                        p = core_find_second(c, mu, c.x, c.xb, c.dxv, M);
                        if (p < 0) {
                            status = DZ_UNBOUNDED;
                            failed = true;
                        }
                    } else {
                        q = core_find_second(c, mu, c.z, c.zb, c.dzv, Nn);
                        if (q < 0) {
                            status = DZ_INFEASIBLE;
                            failed = true;
                        }
                    }
                    ctick(c, CP_RATIO);
                    if (failed) break;
                }
            }
            if (failed || handed_over) break;
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
            __syncthreads();
            for (int k = tid; k < M; k += c.NT) { // fn pivot, simplex.rs:410-421
                const double d = c.dxv[k];
                if (k == p) {
                    c.x[k] = t;
                    c.xb[k] = t_bar;
                } else {
                    c.x[k] = __dsub_rn(c.x[k], __dmul_rn(t, d));
                    c.xb[k] = __dsub_rn(c.xb[k], __dmul_rn(t_bar, d));
                }
            }
            for (int k = tid; k < Nn; k += c.NT) {
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
            // swap (simplex.rs:239-251) and the structure that follows the basis
            {
                const int srl = T.slack_row[leaving], sre = T.slack_row[entering];
                if (srl < 0)
                    for (int e = T.col_ptr[leaving] + tid; e < T.col_ptr[leaving + 1]; e += c.NT)
                        atomicAdd(&c.rowcnt[T.row_idx[e]], -1);
                if (sre < 0)
                    for (int e = T.col_ptr[entering] + tid; e < T.col_ptr[entering + 1]; e += c.NT)
                        atomicAdd(&c.rowcnt[T.row_idx[e]], 1);
                if (tid == 0) {
                    c.bas[p] = entering;
                    c.nb[q] = leaving;
                    if (srl >= 0) c.spos[srl] = -1;
                    if (sre >= 0) c.spos[sre] = p;
                    c.srow[p] = sre;
                    if (Bt.trace && pivots < Bt.trace_cap) {
                        int *tr = Bt.trace + ((size_t)lp * Bt.trace_cap + pivots) * 3;
                        tr[0] = primal_step ? 0 : 1;
                        tr[1] = leaving;
                        tr[2] = entering;
                    }
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
            __syncthreads();
            ctick(c, CP_UPDATE);
        }

        __syncthreads();
        if (c.prof) {
            if (tid == 0) c.prof[CP_HANDED] = handed_over ? 1 : 0;
            __syncthreads();
            if (tid < 16) Bt.prof[(size_t)lp * 16 + tid] = c.prof[tid];
        }
        if (Bt.work) { // executed flop counts of this kernel's share of the LP
            double *w = Bt.work + (size_t)lp * 8;
            if (c.n_lu) atomicAdd(&w[0], (double)c.n_lu);
            if (c.n_solve) atomicAdd(&w[1], (double)c.n_solve);
            if (c.n_price) atomicAdd(&w[2], (double)c.n_price);
            if (tid == 0) atomicAdd(&w[3], (double)n_upd);
        }
        if (handed_over) {
            core_hand_over(c, Bt, lp, pivots, n_primal, hash);
            continue;
        }
        // ---- results: objective_value / solution, simplex.rs:345-371 ----
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
            for (int p = tid; p < M; p += c.NT) Bt.x_basic[(size_t)lp * M + p] = c.x[p];
        if (Bt.basis)
            for (int p = tid; p < M; p += c.NT) Bt.basis[(size_t)lp * M + p] = c.bas[p];
        if (Bt.values) {
            for (int v = tid; v < T.n_orig; v += c.NT) {
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
    }
}

} // namespace

size_t core_fixed_smem_bytes(int M, int Nn, int NQ) {
    const size_t doubles = 5 * (size_t)M + std::max<size_t>(16 * kCoreWarpBuf, (size_t)M) + 3 * (size_t)Nn + 2 + 16 + 2 * 4 * kMaxWarps;
    const size_t ints = 2 * 4 * kMaxWarps + 2 * kMaxWarps + CC_WORDS + 13 * (size_t)M + 1 + (size_t)Nn +
                        (size_t)M * NQ;
    return doubles * 8 + ints * 4 + 16;
}

int core_nq(int M) {
    const int nq = (M + 31) / 32;
    if (nq <= 4) return nq < 1 ? 1 : nq;
    if (nq <= 6) return 6;
    if (nq <= 8) return 8;
    return 0;
}

template <int NQ, int NTH>
static cudaError_t launch_core_one(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan,
                                   cudaStream_t st) {
    auto kern = dz_core_kernel<NQ, NTH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
    if (e != cudaSuccess) return e;
#ifdef DZ_EMU
    (void)st;
    return emu::launch(kern, plan.grid, plan.block, (size_t)plan.smem_bytes, T, Bt, plan.core_cap_w);
#else
    kern<<<plan.grid, plan.block, plan.smem_bytes, st>>>(T, Bt, plan.core_cap_w);
    return cudaGetLastError();
#endif
}

int launch_core(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream,
                std::string *err) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (core_nq(T.M)) {
    case 1: e = launch_core_one<1, 128>(T, Bt, plan, st); break;
    case 2: e = launch_core_one<2, 128>(T, Bt, plan, st); break;
    case 3: e = launch_core_one<3, 128>(T, Bt, plan, st); break;
    case 4: e = launch_core_one<4, 128>(T, Bt, plan, st); break;
    case 6: e = plan.block == 256 ? launch_core_one<6, 256>(T, Bt, plan, st) : launch_core_one<6, 512>(T, Bt, plan, st); break;
    case 8: e = plan.block == 256 ? launch_core_one<8, 256>(T, Bt, plan, st) : launch_core_one<8, 512>(T, Bt, plan, st); break;
    default:
        *err = "dz_core_kernel: m_int > 256";
        return DZ_ERR_LIMIT;
    }
    if (e != cudaSuccess) {
        *err = std::string("dz_core_kernel launch: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    return DZ_OK;
}

} // namespace dz
