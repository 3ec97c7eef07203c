// dz_grid.cu -- one large LP on the whole GPU (BASELINE configs[2] and configs[3]).
//
// A single LP's pivots are serially dependent (simplex.rs:332-343), so it stays on one
// GPU; but each pivot's two lu_solve calls (linalg.rs:8-10), the pricing (linalg.rs:199-207)
// and the vector updates (simplex.rs:410-421) are wide enough for every SM.  This kernel is
// launched cooperatively, one CTA per SM.  CTA 0 is the MASTER: it runs the pivot loop as
// straight-line code and hands every wide phase to the whole grid as a JOB (a descriptor in
// HBM between two grid barriers); the other CTAs sit in a job loop.  Phases that are narrow
// -- the bookkeeping steps of the elimination, the pivot search, an elimination step that
// touches only a few rows (nearly all of them on the sparse config 4), the ordered
// subtraction chains of the back-substitution -- run on the master alone, without any grid
// barrier.
//
// Arithmetic, operation order and the liberties taken are exactly those of dz_core.cu: the
// working matrix holds only the coupled core of the basis (rows touched by a structural basis
// column x columns that are structural or slacks of such rows), dense, in HBM/L2, with a bit
// mask per row of the columns that may be nonzero, so that updates, back-substitution and
// the search for rows to update follow the nonzeros instead of the dimension.  1x1 blocks are
// bookkeeping.  Anything irregular (non-finite pivot/row/multiplier, structurally singular
// basis) hands the LP with its state to the general kernel (BatchDev::exo_*).
//
// Pricing walks each nonbasic column with one warp, forms the products a * (-v_r) with the
// lanes and adds them in row order (linalg.rs:203); a column whose exact negative (the other
// half of a split variable, model.rs:11-22) is nonbasic too is priced once and negated.

#include "dz_device.cuh"

#include <algorithm>
#include <cstdio>
#include <string>

namespace dz {

namespace {

enum {
    J_EXIT = 0, J_INIT, J_INIT2, J_STATUS, J_LISTS_COUNT, J_LISTS_WRITE, J_ZERO, J_SCATTER, J_UPDATE,
    J_BACK_PREP, J_YSCATTER, J_PRICE, J_RATIO, J_VECUPD, J_OBJ, J_BACKSUB
};
// job descriptor words (ints) and doubles
enum { JI_TYPE = 0, JI_A0, JI_A1, JI_A2, JI_A3, JI_A4, JI_NR, JI_EXOTIC, JI_NLIST, JI_NCOLS, JI_NWORDS, JI_CURSOR, JI_WORDS = 16 };
enum { JD_0 = 0, JD_1, JD_2, JD_3, JD_WORDS = 8 };
constexpr int kBufTerms = 2048; // ordered products of one back-substitution row per round (shared memory)
#ifndef DZ_GRID_HEAP_CAP
#define DZ_GRID_HEAP_CAP 4096 // tests shrink it to exercise the overflow path
#endif
constexpr int kHeapCap = DZ_GRID_HEAP_CAP; // positions disturbed by interchanges, waiting for their step (master, shared memory)
#ifndef DZ_GRID_TILE
#define DZ_GRID_TILE 256
#endif
#ifndef DZ_GRID_WIN
#define DZ_GRID_WIN 64
#endif
constexpr int kTile = DZ_GRID_TILE; // pattern columns per work item of an elimination update (multiple of 32)
constexpr int kWin = DZ_GRID_WIN;      // core columns per look-ahead window of the transposed elimination
constexpr int kWarpBuf = kBufTerms / 16; // ... per warp in the grid-wide back-substitution (16 warps per CTA)

// Slots of the optional cycle profile of the master CTA (BatchDev::prof, 16 per LP).
enum { GP_STATUS = 0, GP_LISTS, GP_ZERO_SCATTER, GP_BOOK_FAST, GP_SEARCH, GP_UPDATE, GP_EPOCH, GP_BACK_PREP, GP_BACK_CHAIN,
       GP_PRICE, GP_RATIO, GP_VECUPD, GP_FAST_COMMITS, GP_SLOW_STEPS, GP_EPOCHS, GP_PENDING_ROWS };

struct G {
    long long *prof; // shared memory of the master, or null
    long long t_last;
    // geometry
    int M, Nn, NT, NW, tid, lane, warp, nblocks, blk;
    long long gtid, GT;
    int gwarp, GW;
    // per solve
    int nr, MW;
    long long S;
    // shared memory of this CTA
    double *red_key;
    int *red_idx;
    int *scan;
    int *sctl;
    double *sbuf;  // [kBufTerms]
    unsigned long long *wmax; // [kWin] window: column maxima as bit patterns
    int *wcnt, *wcand;        // [kWin], [4 kWin]: rows attaining them
    int *wgid, *winert;       // [4 kWin] each: the candidates' system rows; whether a step they win changes nothing
    int *heap;                // [kHeapCap] min-heap of disturbed positions; size in sctl[10], overflow in sctl[11]
    int parity;
    unsigned long long n_lu, n_solve, n_price;
    int nse[2]; // entries of the exceptional-position lists seu / set of this basis
    unsigned long long stat_core, stat_real, stat_grid; // master: working-core doubles, steps, grid-wide steps
};

__device__ __forceinline__ void gsync(G &g, const GridDev &D) {
#ifdef DZ_EMU
    emu::grid_sync();
#else
    __syncthreads();
    if (g.tid == 0) {
        volatile unsigned *gen = D.bar + 1;
        __threadfence();
        const unsigned my = *gen;
        if (atomicAdd(D.bar, 1u) == (unsigned)g.nblocks - 1u) {
            atomicExch(D.bar, 0u);
            __threadfence();
            atomicAdd(D.bar + 1, 1u);
        } else {
            while (*gen == my) {
            }
        }
        __threadfence();
    }
    __syncthreads();
#endif
}

__device__ __forceinline__ void gtick(G &g, int slot) {
    if (g.prof && g.tid == 0) {
        const long long now = clock64();
        g.prof[slot] += now - g.t_last;
        g.t_last = now;
    }
}
__device__ __forceinline__ void gcount(G &g, int slot, long long n) {
    if (g.prof && g.tid == 0) g.prof[slot] += n;
}

__device__ __forceinline__ double fast_div(double v, double pv) {
    return (pv == 1.0) ? v : ((pv == -1.0) ? -v : __ddiv_rn(v, pv));
}

// Elimination update of the candidate rows list[first], list[first + stride], ... (one warp
// per row): l = a_ik / pivot, then a_ij -= l * a_kj over the nonzero pattern of the pivot row
// right of the pivot column, the right-hand side included (linalg.rs:118-124, :288-290).  The
// pivot row arrives compacted by the master: its nonzero-pattern columns and values (pcols,
// pvals: ncols entries, any order -- the element updates of one step are independent) and the
// mask words that hold them (pwq, pwb: nwords entries).
__device__ __forceinline__ void update_rows(G &g, const GridDev &D, int first, int stride, int n_list, int cc,
                                            int pr, double pv, int ncols, int nwords) {
    const int nr = g.nr, MW = g.MW, lane = g.lane;
    const long long S = g.S;
    const double urhs = D.W[(size_t)pr * S + nr];
    bool bad = !isfinite(urhs);
    // work items are (candidate row, tile of kTile pattern columns); tile 0 also carries the
    // right-hand side and the row's pattern bookkeeping
    const int ntile = max(1, (ncols + kTile - 1) / kTile);
    const long long items = (long long)n_list * ntile;
    for (long long it = first; it < items; it += stride) {
        const int e = (int)(it / ntile), tile = (int)(it - (long long)e * ntile);
        const int i = D.list[e];
        if (i == pr) continue;
        double *__restrict__ row = D.W + (size_t)i * S;
        const double l = fast_div(row[cc], pv);
        bad = bad || !isfinite(l);
        const int t0 = tile * kTile, t1 = min(ncols, t0 + kTile);
#pragma unroll 4
        for (int t = t0 + lane; t < t1; t += 32) {
            const int j = D.pcols[t];
            row[j] = __dsub_rn(row[j], __dmul_rn(l, D.pvals[t]));
        }
        if (tile == 0) {
            unsigned *__restrict__ imask = D.rmask + (size_t)i * MW;
            for (int t = lane; t < nwords; t += 32) imask[D.pwq[t]] |= D.pwb[t]; // fill pattern
            if (lane == 0) {
                if (ncols > 0) D.rlast[i] = max(D.rlast[i], D.rlast[pr]);
                unsigned long long cnt = 2ull * ncols + 1;
                if (urhs != 0.0) {
                    row[nr] = __dsub_rn(row[nr], __dmul_rn(l, urhs));
                    cnt += 2;
                }
                g.n_lu += cnt;
            }
        }
    }
    if (__ballot_sync(kFull, bad) && lane == 0) D.job[JI_EXOTIC] = 1;
}

// Every wide phase, executed by all CTAs between two grid barriers.
__device__ __forceinline__ void exec_job(G &g, const GridDev &D, const TemplateDev &T, const BatchDev &Bt) {
    const int type = D.job[JI_TYPE];
    const int a0 = D.job[JI_A0], a1 = D.job[JI_A1], a2 = D.job[JI_A2];
    const int M = g.M, Nn = g.Nn;
    const long long lp = D.job[JI_A4];
    const double *__restrict__ theta = Bt.theta + (size_t)lp * Bt.n_theta;
    switch (type) {
    case J_INIT: {
        for (long long e = g.gtid; e < T.nnz; e += g.GT) D.lval[e] = load_ref(theta, T.val_ref[e]);
        for (long long p = g.gtid; p < M; p += g.GT) {
            const int col = T.basis0[p];
            D.bas[p] = col;
            D.x[p] = load_ref(theta, T.b_ref[p]);
            D.xb[p] = 1.0;
            D.rowcnt[p] = 0;
            D.spos[p] = -1;
            D.where[col] = -1 - (int)p;
        }
        for (long long k = g.gtid; k < Nn; k += g.GT) {
            const int col = T.nonbasis0[k];
            D.nb[k] = col;
            D.z[k] = -load_ref(theta, T.c_ref[col]);
            D.zb[k] = 1.0;
            D.where[col] = (int)k;
        }
        break;
    }
    case J_INIT2: {
        for (long long p = g.gtid; p < M; p += g.GT) {
            const int col = D.bas[p];
            const int sr = T.slack_row[col];
            D.srow[p] = sr;
            if (sr >= 0) {
                D.spos[sr] = (int)p;
            } else {
                for (int e = T.col_ptr[col]; e < T.col_ptr[col + 1]; ++e) atomicAdd(&D.rowcnt[T.row_idx[e]], 1);
            }
        }
        break;
    }
    case J_STATUS: { // find_first_pivot on both sides (simplex.rs:423-437): per-CTA partial results
        Cand<4> cd;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            cd.key[n] = 0.0;
            cd.idx[n] = -1;
        }
        for (long long k = g.gtid; k < Nn; k += g.GT) {
            const double yb = D.zb[k];
            if (yb > 0.0) {
                const double ratio = (yb == 1.0) ? -D.z[k] : __ddiv_rn(-D.z[k], yb);
                if (ratio == ratio && beats(ratio, (int)k, cd.key[0], cd.idx[0])) {
                    cd.key[0] = ratio;
                    cd.idx[0] = (int)k;
                }
                if (cd.idx[1] < 0) cd.idx[1] = (int)k;
            }
        }
        for (long long k = g.gtid; k < M; k += g.GT) {
            const double yb = D.xb[k];
            if (yb > 0.0) {
                const double ratio = (yb == 1.0) ? -D.x[k] : __ddiv_rn(-D.x[k], yb);
                if (ratio == ratio && beats(ratio, (int)k, cd.key[2], cd.idx[2])) {
                    cd.key[2] = ratio;
                    cd.idx[2] = (int)k;
                }
                if (cd.idx[3] < 0) cd.idx[3] = (int)k;
            }
        }
        block_argmax<4>(cd, g.red_key, g.red_idx, g.parity, g.NW, g.tid, false);
        if (g.tid < 4) {
            D.pkey[g.blk * 4 + g.tid] = cd.key[g.tid];
            D.pidx[g.blk * 4 + g.tid] = cd.idx[g.tid];
        }
        break;
    }
    case J_RATIO: { // find_second_pivot (simplex.rs:439-461); a0: 1 = on x (primal step), 0 = on z
        const double mu = D.jobd[JD_0];
        const double *y = a0 ? D.x : D.z, *yb = a0 ? D.xb : D.zb, *dy = a0 ? D.dxv : D.dzv;
        const int len = a0 ? M : Nn;
        Cand<1> cd;
        cd.key[0] = 0.0;
        cd.idx[0] = -1;
        for (long long k = g.gtid; k < len; k += g.GT) {
            const double denom = __dadd_rn(y[k], __dmul_rn(mu, yb[k]));
            const double ratio = __ddiv_rn(dy[k], denom);
            if (ratio > 0.0 && beats(ratio, (int)k, cd.key[0], cd.idx[0])) {
                cd.key[0] = ratio;
                cd.idx[0] = (int)k;
            }
        }
        block_argmax<1>(cd, g.red_key, g.red_idx, g.parity, g.NW, g.tid, false);
        if (g.tid == 0) {
            D.pkey[g.blk * 4] = cd.key[0];
            D.pidx[g.blk * 4] = cd.idx[0];
        }
        break;
    }
    case J_LISTS_COUNT:
    case J_LISTS_WRITE: {
        // Four ordered lists of indices in [0, M), built in one pass (CTA b owns the range
        // [b*chunk, (b+1)*chunk), so the concatenation of the CTAs' pieces is in index order):
        //   0  core rows: constraint rows a structural basis column touches   (rmapR, rlist)
        //   1  core positions: structural, or the slack of a core row          (pmap, plist)
        //   2  positions whose elimination step in B is not a no-op from the start: anything but
        //      "the slack of row p sits at position p"                         (seu)
        //   3  the same for B^T, whose columns are the constraint rows          (set)
        const int chunk = (M + g.nblocks - 1) / g.nblocks;
        const int lo = g.blk * chunk, hi = min(M, lo + chunk);
        const unsigned lt = (1u << g.lane) - 1u;
        int off[4], tot[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 4; ++q) off[q] = type == J_LISTS_WRITE ? D.bcnt[(4 + q) * g.nblocks + g.blk] : 0;
        for (int base = lo; base < hi; base += g.NT) {
            const int i = base + g.tid;
            bool pr[4] = {false, false, false, false};
            if (i < hi) {
                const int rc = D.rowcnt[i], sr = D.srow[i];
                pr[0] = rc > 0;
                pr[1] = sr < 0 || D.rowcnt[sr] > 0;
                pr[2] = sr != i;
                pr[3] = rc > 0 || D.spos[i] != i;
            }
            unsigned mk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mk[q] = __ballot_sync(kFull, pr[q]);
                if (g.lane == 0) g.scan[q * kMaxWarps + g.warp] = __popc(mk[q]);
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int before = 0, all = 0;
                for (int w = 0; w < g.NW; ++w) {
                    const int a = g.scan[q * kMaxWarps + w];
                    if (w < g.warp) before += a;
                    all += a;
                }
                if (type == J_LISTS_WRITE && i < hi) {
                    const int at = pr[q] ? off[q] + tot[q] + before + __popc(mk[q] & lt) : -1;
                    if (q == 0) {
                        D.rmapR[i] = at;
                        if (pr[q]) D.rlist[at] = i;
                    } else if (q == 1) {
                        D.pmap[i] = at;
                        if (pr[q]) D.plist[at] = i;
                    } else if (pr[q]) {
                        (q == 2 ? D.seu : D.set_)[at] = i;
                    }
                }
                tot[q] += all;
            }
            __syncthreads();
        }
        if (type == J_LISTS_COUNT && g.tid < 4) D.bcnt[g.tid * g.nblocks + g.blk] = tot[g.tid];
        break;
    }
    case J_ZERO: { // a0 = rows of W and of rmask to clear; also the tables of a fresh elimination
        // a2 = 1: clear everything any solve of this LP may have touched (count in jobd[JD_1])
        const long long nw = a2 ? (long long)D.jobd[JD_1] : (long long)a0 * g.S;
        const long long nm = a2 ? 0 : (long long)a0 * g.MW;
        for (long long e = g.gtid; e < nw; e += g.GT) D.W[e] = 0.0;
        for (long long e = g.gtid; e < nm; e += g.GT) D.rmask[e] = 0u;
        for (long long e = g.gtid; e < a0; e += g.GT) {
            D.rlast[e] = -1;
            D.done[e] = 0;
        }
        double *y = a1 ? D.vv : D.dxv;
        for (long long i = g.gtid; i < M; i += g.GT) {
            D.rowAt[i] = (int)i;
            D.posOf[i] = (int)i;
            D.pivr[i] = -1;
            y[i] = 0.0;
        }
        break;
    }
    case J_SCATTER: { // a0 = transposed, a1 = arg: core of B or B^T into [W | rhs], 1x1 blocks into y
        const bool transposed = a0 != 0;
        const int nr = g.nr, MW = g.MW;
        const long long S = g.S;
        double *y = transposed ? D.vv : D.dxv;
        for (int cc = g.gwarp; cc < nr; cc += g.GW) {
            const int col = D.bas[D.plist[cc]];
            if (!transposed && g.lane == 0) { // a slack column of the core is its row's pivot column unless
                                              // the row is used up before (then its step says otherwise)
                const int sr = D.srow[D.plist[cc]];
                if (sr >= 0) D.pivr[cc] = D.rmapR[sr];
            }
            for (int e = T.col_ptr[col] + g.lane; e < T.col_ptr[col + 1]; e += 32) {
                const double val = D.lval[e];
                if (val == 0.0) continue; // exact zeros are not stored (linalg.rs:261)
                const int ir = D.rmapR[T.row_idx[e]];
                const int wi = transposed ? cc : ir, wj = transposed ? ir : cc;
                D.W[(size_t)wi * S + wj] = val;
                atomicOr(&D.rmask[(size_t)wi * MW + (wj >> 5)], 1u << (wj & 31));
                atomicMax(&D.rlast[wi], wj);
            }
        }
        if (transposed) {
            if (g.gtid == 0) {
                const int ci = D.pmap[a1];
                if (ci >= 0)
                    D.W[(size_t)ci * S + nr] = 1.0;
                else
                    y[D.srow[a1]] = 1.0;
            }
        } else {
            for (long long e = T.col_ptr[a1] + g.gtid; e < T.col_ptr[a1 + 1]; e += g.GT) {
                const double val = D.lval[e];
                if (val == 0.0) continue;
                const int r = T.row_idx[e];
                const int ir = D.rmapR[r];
                if (ir >= 0)
                    D.W[(size_t)ir * S + nr] = val;
                else
                    y[D.spos[r]] = val;
            }
        }
        break;
    }
    case J_UPDATE:
        update_rows(g, D, g.gwarp, g.GW, D.job[JI_NLIST], a1, a2, D.jobd[JD_0], D.job[JI_NCOLS], D.job[JI_NWORDS]);
        break;
    case J_BACK_PREP: {
        // rows whose strict upper part is empty are solved at once (x/1 == x); the others are
        // left to the master's ordered chains.  pend[cc] = 1 marks them.
        const int nr = g.nr, MW = g.MW;
        const long long S = g.S;
        for (int cc = g.gwarp; cc < nr; cc += g.GW) {
            const int i = D.pivr[cc];
            const unsigned *mask = D.rmask + (size_t)i * MW;
            const int q0 = (cc + 1) >> 5;
            bool any = false;
            for (int q = q0 + g.lane; q < MW; q += 32) {
                unsigned word = mask[q];
                if (q == q0) {
                    const int lo = cc + 1 - 32 * q0;
                    word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
                }
                any = any || word != 0u;
            }
            any = __ballot_sync(kFull, any) != 0u;
            if (g.lane == 0) {
                D.pend[cc] = any ? 1 : 0;
                if (!any) {
                    const double *row = D.W + (size_t)i * S;
                    const double d = row[cc], s = row[nr];
                    const double yi = (d == 1.0) ? s : __ddiv_rn(s, d);
                    D.ycore[cc] = yi;
                    D.done[cc] = 1;
                    if (!isfinite(yi)) D.job[JI_EXOTIC] = 2; // non-finite component: see the master
                    g.n_solve += 1;
                }
            }
        }
        break;
    }
    case J_BACKSUB: {
        // Back substitution (linalg.rs:292-297) of the rows the prep job left, LEVEL-SCHEDULED by
        // data flow: every warp of the grid takes core columns in descending order from a shared
        // cursor and waits, entry by entry of its row's pattern, for the components it needs
        // (done[j]); rows that do not depend on each other proceed side by side.  A row's
        // products u_ij * y_j are formed by the lanes, staged in column order in the warp's slice
        // of shared memory and subtracted one after the other (ascending j, linalg.rs:294).
        const int nr = g.nr, MW = g.MW, lane = g.lane;
        const long long S = g.S;
        double *buf = g.sbuf + (size_t)g.warp * kWarpBuf;
        const unsigned lt = (1u << lane) - 1u;
        for (;;) {
            int idx = 0;
            if (lane == 0) idx = atomicAdd(&D.job[JI_CURSOR], 1);
            idx = __shfl_sync(kFull, idx, 0);
            if (idx >= nr) break;
            const int cc = nr - 1 - idx;
            if (!D.pend[cc]) continue;
            const int i = D.pivr[cc];
            const double *__restrict__ row = D.W + (size_t)i * S;
            const unsigned *__restrict__ mask = D.rmask + (size_t)i * MW;
            double s = row[nr];
            const double d = row[cc];
            const int q0 = (cc + 1) >> 5;
            int nbuf = 0;
            unsigned long long ops = 0;
            for (int wb = q0; wb < MW; wb += 32) {
                const int q = wb + lane;
                unsigned word = (q < MW) ? mask[q] : 0u;
                if (q == q0) {
                    const int lo = cc + 1 - 32 * q0;
                    word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
                }
                unsigned nzw = __ballot_sync(kFull, word != 0u);
                while (nzw) {
                    // four mask words (128 columns) per trip: their flags, then their loads, are
                    // all in flight together
                    bool has[4];
                    int jc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        has[u] = false;
                        jc[u] = 0;
                        if (nzw) {
                            const int wq = __ffs(nzw) - 1;
                            nzw &= nzw - 1;
                            const unsigned bits = __shfl_sync(kFull, word, wq);
                            has[u] = (bits >> lane) & 1u;
                            jc[u] = 32 * (wb + wq) + lane;
                        }
                    }
                    for (;;) { // wait for the components these columns need
                        bool ok = true;
#pragma unroll
                        for (int u = 0; u < 4; ++u) ok = ok && (!has[u] || *((volatile int *)&D.done[jc[u]]) != 0);
                        if (__ballot_sync(kFull, !ok) == 0u) break;
#ifdef DZ_EMU
                        emu::yield();
#endif
                    }
                    __threadfence();
                    double uu[4], yy[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uu[u] = has[u] ? row[jc[u]] : 0.0;
#ifdef DZ_EMU
                        yy[u] = has[u] ? D.ycore[jc[u]] : 0.0;
#else
                        yy[u] = has[u] ? __ldcg(&D.ycore[jc[u]]) : 0.0;
#endif
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double p = __dmul_rn(uu[u], yy[u]);
                        const bool take = has[u] && p != 0.0; // subtracting an exact zero changes nothing
                        const unsigned mk = __ballot_sync(kFull, take);
                        if (nbuf + __popc(mk) > kWarpBuf) { // flush the staged products, in order
                            __syncwarp();
                            for (int t = 0; t < nbuf; ++t) s = __dsub_rn(s, buf[t]);
                            ops += 2ull * nbuf;
                            nbuf = 0;
                            __syncwarp();
                        }
                        if (take) buf[nbuf + __popc(mk & lt)] = p;
                        nbuf += __popc(mk);
                    }
                }
            }
            __syncwarp();
            {
                int t = 0;
                for (; t + 4 <= nbuf; t += 4) {
                    const double p0 = buf[t], p1 = buf[t + 1], p2 = buf[t + 2], p3 = buf[t + 3];
                    s = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(s, p0), p1), p2), p3);
                }
                for (; t < nbuf; ++t) s = __dsub_rn(s, buf[t]);
                ops += 2ull * nbuf + 1;
            }
            const double yi = (d == 1.0) ? s : __ddiv_rn(s, d);
            if (lane == 0) {
                D.ycore[cc] = yi;
                __threadfence();
                atomicExch((unsigned *)&D.done[cc], 1u);
                if (!isfinite(yi)) D.job[JI_EXOTIC] = 2;
                g.n_solve += ops;
            }
            __syncwarp();
        }
        break;
    }
    case J_YSCATTER: { // a0 = transposed
        double *y = a0 ? D.vv : D.dxv;
        const int *clist = a0 ? D.rlist : D.plist;
        for (long long cc = g.gtid; cc < g.nr; cc += g.GT) y[clist[cc]] = D.ycore[cc];
        break;
    }
    case J_PRICE: { // dz = -N^T v (simplex.rs:235, linalg.rs:199-207), one warp per nonbasic column
        for (int k = g.gwarp; k < Nn; k += g.GW) {
            const int col = D.nb[k];
            const int tw = T.twin[col];
            int ktw = -1;
            if (tw >= 0) {
                ktw = D.where[tw];
                if (ktw >= 0 && tw < col) continue; // the twin prices both
            }
            double s = 0.0;
            unsigned long long cnt = 0;
            const int e0 = T.col_ptr[col], e1 = T.col_ptr[col + 1];
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + g.lane;
                double p = 0.0;
                bool take = false;
                if (e < e1) {
                    const double vr = D.vv[T.row_idx[e]];
                    if (vr != 0.0) {
                        const double a = D.lval[e];
                        take = a != 0.0;
                        p = __dmul_rn(a, -vr);
                    }
                }
                unsigned mk = __ballot_sync(kFull, take);
                cnt += 2ull * __popc(mk);
                while (mk) { // ascending row order
                    const int b = __ffs(mk) - 1;
                    mk &= mk - 1;
                    s = __dadd_rn(s, __shfl_sync(kFull, p, b));
                }
            }
            if (g.lane == 0) {
                D.dzv[k] = s;
                if (ktw >= 0) D.dzv[ktw] = -s; // every product and partial sum of the twin is the exact negative
                g.n_price += cnt;
            }
        }
        break;
    }
    case J_VECUPD: { // fn pivot (simplex.rs:410-421) on the four vectors; then the basis change
        const int p = a0, q = a1;
        const double t = D.jobd[JD_0], tb = D.jobd[JD_1], s = D.jobd[JD_2], sb = D.jobd[JD_3];
        for (long long k = g.gtid; k < M; k += g.GT) {
            const double dd = D.dxv[k];
            if (k == p) {
                D.x[k] = t;
                D.xb[k] = tb;
            } else {
                D.x[k] = __dsub_rn(D.x[k], __dmul_rn(t, dd));
                D.xb[k] = __dsub_rn(D.xb[k], __dmul_rn(tb, dd));
            }
        }
        for (long long k = g.gtid; k < Nn; k += g.GT) {
            const double dd = D.dzv[k];
            if (k == q) {
                D.z[k] = s;
                D.zb[k] = sb;
            } else {
                D.z[k] = __dsub_rn(D.z[k], __dmul_rn(s, dd));
                D.zb[k] = __dsub_rn(D.zb[k], __dmul_rn(sb, dd));
            }
        }
        // a2 = leaving column, a3 = entering column: structure that follows the basis
        const int leaving = a2, entering = D.job[JI_A3];
        if (T.slack_row[leaving] < 0)
            for (long long e = T.col_ptr[leaving] + g.gtid; e < T.col_ptr[leaving + 1]; e += g.GT)
                atomicAdd(&D.rowcnt[T.row_idx[e]], -1);
        if (T.slack_row[entering] < 0)
            for (long long e = T.col_ptr[entering] + g.gtid; e < T.col_ptr[entering + 1]; e += g.GT)
                atomicAdd(&D.rowcnt[T.row_idx[e]], 1);
        break;
    }
    case J_OBJ: { // products c_k * x_k by basis position (simplex.rs:345-352), summed in order by the master
        for (long long p = g.gtid; p < M; p += g.GT)
            D.ycore[p] = __dmul_rn(load_ref(theta, T.c_ref[D.bas[p]]), D.x[p]);
        if (Bt.x_basic)
            for (long long p = g.gtid; p < M; p += g.GT) Bt.x_basic[(size_t)lp * M + p] = D.x[p];
        if (Bt.basis)
            for (long long p = g.gtid; p < M; p += g.GT) Bt.basis[(size_t)lp * M + p] = D.bas[p];
        if (Bt.values)
            for (long long v = g.gtid; v < T.n_orig; v += g.GT) {
                const int wp = D.where[T.pos_index[v]], wn = D.where[T.neg_index[v]];
                const double pos = wp < 0 ? D.x[-1 - wp] : 0.0, neg = wn < 0 ? D.x[-1 - wn] : 0.0;
                Bt.values[(size_t)lp * T.n_orig + v] = __dsub_rn(pos, neg);
            }
        break;
    }
    default: break;
    }
}

// master: post a job, run it with everybody, wait for it (a real call: the job switch is not
// replicated at every call site)
__device__ __noinline__ void dispatch(G &g, const GridDev &D, const TemplateDev &T, const BatchDev &Bt, int type, int a0 = 0,
                                         int a1 = 0, int a2 = 0, int a3 = 0) {
    __syncthreads();
    if (g.tid == 0) {
        D.job[JI_TYPE] = type;
        D.job[JI_A0] = a0;
        D.job[JI_A1] = a1;
        D.job[JI_A2] = a2;
        D.job[JI_A3] = a3;
        D.job[JI_NR] = g.nr;
    }
    gsync(g, D);
    if (type != J_EXIT) {
        exec_job(g, D, T, Bt);
        gsync(g, D);
    }
}

// min-heap of positions an interchange has disturbed (one thread: lane 0 of the control warp)
__device__ __forceinline__ void heap_push(G &g, int v) {
    int n = g.sctl[10];
    if (n >= kHeapCap) {
        g.sctl[11] = 1; // overflow: the control warp goes back to scanning every position
        return;
    }
    int i = n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        const int pv = g.heap[p];
        if (pv <= v) break;
        g.heap[i] = pv;
        i = p;
    }
    g.heap[i] = v;
    g.sctl[10] = n;
}
__device__ __forceinline__ void heap_pop(G &g) {
    int n = g.sctl[10] - 1;
    g.sctl[10] = n;
    if (n <= 0) return;
    const int v = g.heap[n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && g.heap[c + 1] < g.heap[c]) ++c;
        if (g.heap[c] >= v) break;
        g.heap[i] = g.heap[c];
        i = c;
    }
    g.heap[i] = v;
}

__device__ __forceinline__ void g_swap_pos(G &g, const GridDev &D, int k, int mu) {
    if (mu == k) return;
    const int rk = D.rowAt[k], rm = D.rowAt[mu];
    D.rowAt[k] = rm;
    D.rowAt[mu] = rk;
    D.posOf[rm] = k;
    D.posOf[rk] = mu;
    heap_push(g, mu); // position mu no longer holds what its step expects
}

// master: reduce the per-CTA partial arg-max results of J_STATUS / J_RATIO (slot n)
__device__ __forceinline__ int reduce_partials(G &g, const GridDev &D, int n, double *key_out) {
    // every thread of the master walks the same short list (nblocks entries)
    double bk = 0.0;
    int bi = -1;
    for (int b = 0; b < g.nblocks; ++b) {
        const double k2 = D.pkey[b * 4 + n];
        const int i2 = D.pidx[b * 4 + n];
        if (beats(k2, i2, bk, bi)) {
            bk = k2;
            bi = i2;
        }
    }
    if (key_out) *key_out = bk;
    return bi;
}

// lu_solve on the grid; see dz_core.cu core_solve for the scheme.  Master only (it dispatches).
__device__ __forceinline__ bool grid_solve(G &g, const GridDev &D, const TemplateDev &T, const BatchDev &Bt, const bool transposed,
                                           const int arg, long long lp) {
    const int M = g.M, nr = g.nr, tid = g.tid, lane = g.lane, warp = g.warp;
    g.S = (long long)((nr + 1) | 1);
    g.MW = (nr + 31) >> 5;
    const long long S = g.S;
    const int MW = g.MW;
    const int *__restrict__ rmap = transposed ? D.pmap : D.rmapR;
    const int *__restrict__ cmap = transposed ? D.rmapR : D.pmap;
    const int *__restrict__ rlist = transposed ? D.plist : D.rlist;
    double *y = transposed ? D.vv : D.dxv;
    (void)lp;
    if (tid == 0) {
        D.job[JI_EXOTIC] = 0;
        g.sctl[10] = 0; // heap of disturbed positions: empty
        g.sctl[11] = 0; // ... not overflowed
    }
    dispatch(g, D, T, Bt, J_ZERO, nr, transposed ? 1 : 0);
    dispatch(g, D, T, Bt, J_SCATTER, transposed ? 1 : 0, arg);
    gtick(g, GP_ZERO_SCATTER);

    // ---- elimination (master; dense steps go to the grid) ----
    // WINDOWS.  In the transposed system nearly every core column is won by a row that has
    // nothing right of the pivot (the slack's row of B^T) and a zero right-hand side: such a
    // step records an interchange and changes no value.  So while W does not change, the
    // master finds the column maxima of the next kWin core columns at once (its threads walk
    // the rows' masks; maxima, tie counts and up to four tied rows per column go to shared
    // memory), and the control warp then commits step after step from them -- checking that
    // the stored candidates are still there and breaking ties by their CURRENT positions --
    // without a search and without a CTA barrier.  A step that changes W ends the window.
    int k = 0;
    bool epoch_valid = false;
    int epoch_lo = 0, epoch_hi = 0;
    int cc_next = 0, cooldown = 0; // transposed: core columns are real steps, in order
    int sp = 0;                    // cursor in the exceptional-position list (control warp)
    for (;;) {
        if (transposed && !epoch_valid && cooldown == 0 && nr - cc_next >= 8) {
            epoch_lo = cc_next;
            epoch_hi = min(nr, cc_next + kWin);
            if (tid < kWin) {
                g.wmax[tid] = 0ull;
                g.wcnt[tid] = 0;
            }
            if (tid == 0) g.sctl[8] = 0; // fast commits of this window
            __syncthreads();
            const int qa = epoch_lo >> 5, qb = (epoch_hi - 1) >> 5; // at most three mask words
            bool wbad = false;
            const int nwq = qb - qa + 1;
            for (int pass = 0; pass < 2; ++pass) {
                // work items are (row, mask word of the window): a dense row's words go to
                // neighbouring threads instead of all to one
                for (int it = tid; it < nr * nwq; it += g.NT) {
                    const int i = it / nwq, q = qa + (it - i * nwq);
                    unsigned word = D.rmask[(size_t)i * MW + q];
                    const int lo = epoch_lo - 32 * q, hi = epoch_hi - 32 * q; // keep bits [lo, hi)
                    if (lo > 0) word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
                    if (hi < 32) word = hi <= 0 ? 0u : (word & ((1u << hi) - 1u));
                    if (!word) continue; // nothing of this row in this part of the window
                    const int gid = rlist[i];
                    if (D.posOf[gid] < k) continue;
                    const double *__restrict__ row = D.W + (size_t)i * S;
                    while (word) { // eight independent loads at a time
                        int jj[8];
                        double vv8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            jj[e] = -1;
                            if (word) {
                                jj[e] = 32 * q + __ffs(word) - 1;
                                word &= word - 1;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) vv8[e] = jj[e] >= 0 ? row[jj[e]] : 0.0;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (jj[e] < 0) continue;
                            const double v = vv8[e];
                            wbad = wbad || !isfinite(v);
                            if (v == 0.0) continue;
                            const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v));
                            const int slot_j = jj[e] - epoch_lo;
                            if (pass == 0) {
                                // most entries are not the maximum: look before the atomic
                                if (key > *((volatile unsigned long long *)&g.wmax[slot_j])) atomicMax(&g.wmax[slot_j], key);
                            } else if (key == g.wmax[slot_j]) {
                                const int slot = atomicAdd(&g.wcnt[slot_j], 1);
                                if (slot < 4) {
                                    // everything a commit needs to know about this candidate
                                    g.wcand[4 * slot_j + slot] = i;
                                    g.wgid[4 * slot_j + slot] = gid;
                                    g.winert[4 * slot_j + slot] = (D.rlast[i] <= jj[e] && row[nr] == 0.0) ? 1 : 0;
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
            if (wbad) g.sctl[1] = 1;
            __syncthreads();
            if (g.sctl[1]) return false;
            epoch_valid = true;
            gtick(g, GP_EPOCH);
            gcount(g, GP_EPOCHS, 1);
        }
        if (warp == 0) {
            int irregular = 0, need_refresh = 0;
            const int *__restrict__ se = transposed ? D.set_ : D.seu;
            const int nse = g.nse[transposed ? 1 : 0];
            for (;;) { // bookkeeping steps, see dz_core.cu
                if (!g.sctl[11]) {
                    // JUMP to the next position whose step is not known to be a no-op: the next
                    // entry of this basis's exceptional list, or the smallest position an
                    // interchange of this elimination has disturbed (heap), whichever comes first
                    int nk = 0;
                    if (lane == 0) {
                        while (sp < nse && se[sp] < k) ++sp;
                        const int a = sp < nse ? se[sp] : M;
                        while (g.sctl[10] > 0 && g.heap[0] < k) heap_pop(g);
                        const int b = g.sctl[10] > 0 ? g.heap[0] : M;
                        nk = min(a, b);
                    }
                    nk = __shfl_sync(kFull, nk, 0);
                    sp = __shfl_sync(kFull, sp, 0);
                    k = min(nk, M - 1);
                    if (k >= M - 1) break;
                } else {
                    // (heap overflow) every position, a run of no-ops 32 at a time
                    const int kk = k + lane;
                    bool noop = false;
                    if (kk < M - 1) {
                        const int cc = cmap[kk];
                        const int u = transposed ? (cc < 0 ? D.spos[kk] : -1) : D.srow[kk];
                        noop = u >= 0 && D.rowAt[kk] == u;
                        if (noop && cc >= 0) D.pivr[cc] = rmap[u];
                    }
                    const unsigned m = __ballot_sync(kFull, noop);
                    const int run = (m == kFull) ? 32 : __ffs(~m) - 1;
                    k += run;
                    if (run == 32) continue;
                    if (k >= M - 1) break;
                }
                int adv = 0;
                if (lane == 0) {
                    const int cc = cmap[k];
                    const int u = transposed ? (cc < 0 ? D.spos[k] : -1) : D.srow[k];
                    if (u >= 0 && D.rowAt[k] == u) { // a no-op after all
                        if (cc >= 0) D.pivr[cc] = rmap[u];
                        adv = 1;
                    } else if (cc < 0) {
                        const int pu = u >= 0 ? D.posOf[u] : -1;
                        if (pu < k) {
                            irregular = 1;
                        } else {
                            g_swap_pos(g, D, k, pu);
                            adv = 1;
                        }
                    } else if (u >= 0 && D.posOf[u] >= k) {
                        g_swap_pos(g, D, k, D.posOf[u]);
                        D.pivr[cc] = rmap[u];
                        adv = 1;
                    }
                }
                adv = __shfl_sync(kFull, adv, 0);
                irregular = __shfl_sync(kFull, irregular, 0);
                __syncwarp();
                if (!adv && !irregular && epoch_valid) {
                    // a real step: commit it from the window's column maxima if it changes nothing
                    const int cc = cmap[k];
                    if (cc >= epoch_hi) {
                        need_refresh = 1; // window used up: the next one is computed before going on
                    } else if (cc >= epoch_lo) {
                        const int cnt = g.wcnt[cc - epoch_lo];
                        if (cnt >= 1 && cnt <= 4) {
                            const int row = lane < cnt ? g.wcand[4 * (cc - epoch_lo) + lane] : -1;
                            int pos = row >= 0 ? D.posOf[g.wgid[4 * (cc - epoch_lo) + lane]] : 0x7fffffff;
                            if (pos < k) pos = 0x7fffffff; // that candidate has been used up since
                            const int best = __reduce_min_sync(kFull, pos);
                            if (best != 0x7fffffff) {
                                const int src = __ffs(__ballot_sync(kFull, pos == best)) - 1;
                                const int pr = __shfl_sync(kFull, row, src);
                                const bool inert = g.winert[4 * (cc - epoch_lo) + src] != 0;
                                if (inert) {
                                    if (lane == 0) {
                                        g_swap_pos(g, D, k, best); // linalg.rs:107-114
                                        D.pivr[cc] = pr;
                                        g.sctl[8] += 1;
                                        if (g.prof) g.prof[GP_FAST_COMMITS] += 1;
                                    }
                                    adv = 1;
                                    ++cc_next;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (!adv) break;
                ++k;
            }
            if (lane == 0) {
                g.sctl[0] = k;
                g.sctl[1] = irregular;
                g.sctl[2] = 0; // candidate rows with a nonzero in the pivot column
                g.sctl[7] = cc_next;
                g.sctl[9] = need_refresh;
            }
        }
        __syncthreads();
        gtick(g, GP_BOOK_FAST);
        k = g.sctl[0];
        cc_next = g.sctl[7];
        if (g.sctl[1]) return false;
        if (k >= M - 1) break;
        if (g.sctl[9]) { // the window is used up
            epoch_valid = false;
            __syncthreads();
            continue;
        }
        // pivot search in core column cc over the rows at positions >= k (linalg.rs:98-105):
        // largest |a_ik|, ties to the smallest position; rows with a nonzero entry are listed
        const int cc = cmap[k];
        Cand<1> cd;
        cd.key[0] = 0.0;
        cd.idx[0] = -1; // idx = position (unique per row)
        bool bad = false;
        for (int base = tid; base < nr; base += 8 * g.NT) {
            // only rows whose pattern has this column can hold a nonzero there; eight rows per trip
            // so that mask words, row ids, positions and values are batches of independent loads
            bool on[8];
            int gid8[8], pos8[8];
            double v8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = base + e * g.NT;
                on[e] = i < nr && ((D.rmask[(size_t)i * MW + (cc >> 5)] >> (cc & 31)) & 1u);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) gid8[e] = on[e] ? rlist[base + e * g.NT] : 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                pos8[e] = on[e] ? D.posOf[gid8[e]] : -1;
                v8[e] = on[e] ? D.W[(size_t)(base + e * g.NT) * S + cc] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (!on[e] || pos8[e] < k) continue;
                const int i = base + e * g.NT, pos = pos8[e];
                const double v = v8[e];
                bad = bad || !isfinite(v);
                if (v != 0.0) {
                    const double av = fabs(v);
                    if (beats(av, pos, cd.key[0], cd.idx[0])) {
                        cd.key[0] = av;
                        cd.idx[0] = pos;
                    }
                    D.list[atomicAdd(&g.sctl[2], 1)] = i;
                }
            }
        }
        block_argmax<1>(cd, g.red_key, g.red_idx, g.parity, g.NW, g.tid, false);
        if (bad) g.sctl[1] = 1;
        __syncthreads();
        if (g.sctl[1]) return false;
        const int n_list = g.sctl[2];
        int pr;
        double pv;
        if (cd.idx[0] < 0) { // no nonzero candidate: the incumbent stays, pivot 0, step skipped (linalg.rs:117)
            pr = rmap[D.rowAt[k]];
            if (pr < 0) return false;
            pv = 0.0;
            if (tid == 0) D.pivr[cc] = pr;
        } else {
            const int ppos = cd.idx[0];
            pr = rmap[D.rowAt[ppos]];
            pv = D.W[(size_t)pr * S + cc];
            __syncthreads(); // everybody has read rowAt before the interchange
            if (tid == 0) {
                g_swap_pos(g, D, k, ppos); // linalg.rs:107-114
                D.pivr[cc] = pr;
            }
        }
        g.stat_real += 1;
        gtick(g, GP_SEARCH);
        gcount(g, GP_SLOW_STEPS, 1);
        if (transposed) ++cc_next; // this core column is done
        if (cooldown > 0) --cooldown;
        if (pv != 0.0 && n_list > 1) {
            // compact the pivot row right of the pivot: pattern columns + values, mask words
            __syncthreads();
            if (tid == 0) {
                g.sctl[5] = 0;
                g.sctl[6] = 0;
            }
            __syncthreads();
            {
                const double *__restrict__ prow = D.W + (size_t)pr * S;
                const unsigned *__restrict__ pmask = D.rmask + (size_t)pr * MW;
                const int q0 = (cc + 1) >> 5;
                bool badrow = false;
                for (int q = q0 + tid; q < MW; q += g.NT) {
                    unsigned word = pmask[q];
                    if (q == q0) {
                        const int lo = cc + 1 - 32 * q0;
                        word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
                    }
                    if (!word) continue;
                    const int wl = atomicAdd(&g.sctl[6], 1);
                    D.pwq[wl] = q;
                    D.pwb[wl] = word;
                    int at = atomicAdd(&g.sctl[5], __popc(word));
                    while (word) {
                        const int b = __ffs(word) - 1;
                        word &= word - 1;
                        const int j = 32 * q + b;
                        const double u = prow[j];
                        badrow = badrow || !isfinite(u);
                        D.pcols[at] = j;
                        D.pvals[at] = u;
                        ++at;
                    }
                }
                if (badrow) g.sctl[1] = 1;
            }
            __syncthreads();
            if (g.sctl[1]) return false;
            const int ncols = g.sctl[5], nwords = g.sctl[6];
            if (epoch_valid && (ncols > 0 || D.W[(size_t)pr * S + nr] != 0.0)) {
                epoch_valid = false; // W changes: the column maxima are stale
                if (g.sctl[8] < 2) cooldown = 4; // the window did not pay: a few plain steps first
            }
            if ((long long)n_list * max(ncols, 32) <= 64 * g.NW) { // small step: the master's own warps, no grid barrier
                update_rows(g, D, warp, g.NW, n_list, cc, pr, pv, ncols, nwords);
                __syncthreads();
            } else {
                if (tid == 0) {
                    D.job[JI_NLIST] = n_list;
                    D.job[JI_NCOLS] = ncols;
                    D.job[JI_NWORDS] = nwords;
                    D.jobd[JD_0] = pv;
                }
                dispatch(g, D, T, Bt, J_UPDATE, k, cc, pr);
                g.stat_grid += 1;
            }
            if (D.job[JI_EXOTIC] == 1) return false;
        }
        __syncthreads();
        gtick(g, GP_UPDATE);
        ++k;
    }
    // the row left at the last position is the pivot row of the last column
    {
        const int cc = cmap[M - 1], r = D.rowAt[M - 1];
        if (cc >= 0) {
            if (rmap[r] < 0) return false;
            if (tid == 0) D.pivr[cc] = rmap[r];
        } else {
            const int u = transposed ? D.spos[M - 1] : D.srow[M - 1];
            if (u != r) return false;
        }
    }
    // ---- back substitution (linalg.rs:292-297) ----
    dispatch(g, D, T, Bt, J_BACK_PREP);
    gtick(g, GP_BACK_PREP);
    if (tid == 0) D.job[JI_CURSOR] = 0;
    dispatch(g, D, T, Bt, J_BACKSUB);
    const bool nonfinite = D.job[JI_EXOTIC] == 2;
    dispatch(g, D, T, Bt, J_YSCATTER, transposed ? 1 : 0);
    gtick(g, GP_BACK_CHAIN);
    if (nonfinite) {
        // see dz_core.cu: the literal arithmetic multiplies a non-finite component into every
        // earlier row, by an exact zero wherever the row has no entry in that column
        if (tid == 0) {
            bool any_nan = false;
            int n_inf = 0;
            int *inf_list = D.list;
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            for (int kcol = M - 1; kcol >= 0; --kcol) {
                const int cc = cmap[kcol];
                double v = y[kcol];
                if (any_nan) {
                    v = qnan;
                } else if (n_inf > 0) {
                    bool all_nz = cc >= 0;
                    if (cc >= 0) {
                        const double *row = D.W + (size_t)D.pivr[cc] * S;
                        for (int t = 0; t < n_inf && all_nz; ++t) {
                            const double e = row[inf_list[t]];
                            all_nz = e != 0.0; // products were formed in place? no: W still holds U here
                        }
                    }
                    if (!all_nz) v = qnan;
                }
                y[kcol] = v;
                if (v != v)
                    any_nan = true;
                else if (!isfinite(v) && cc >= 0)
                    inf_list[n_inf++] = cc;
                else if (!isfinite(v))
                    any_nan = true;
            }
        }
        __syncthreads();
    }
    return true;
}

__global__ void __launch_bounds__(512, 1)
dz_grid_kernel(const TemplateDev T, const BatchDev Bt, const GridDev D) {
#ifdef DZ_EMU
    unsigned char *smem_raw = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    G g;
    g.M = T.M;
    g.Nn = T.Nn;
    g.NT = (int)blockDim.x;
    g.NW = g.NT >> 5;
    g.tid = (int)threadIdx.x;
    g.lane = g.tid & 31;
    g.warp = g.tid >> 5;
    g.nblocks = (int)gridDim.x;
    g.blk = (int)blockIdx.x;
    g.gtid = (long long)g.blk * g.NT + g.tid;
    g.GT = (long long)g.nblocks * g.NT;
    g.gwarp = (int)(g.gtid >> 5);
    g.GW = (int)(g.GT >> 5);
    g.parity = 0;
    g.nr = 0;
    g.S = 1;
    g.MW = 1;
    g.n_lu = g.n_solve = g.n_price = 0;
    g.stat_core = g.stat_real = g.stat_grid = 0;
    {
        double *dp = reinterpret_cast<double *>(smem_raw);
        g.sbuf = dp, dp += kBufTerms;
        g.wmax = reinterpret_cast<unsigned long long *>(dp), dp += kWin;
        g.red_key = dp, dp += 2 * 4 * kMaxWarps;
        int *ip = reinterpret_cast<int *>(dp);
        g.red_idx = ip, ip += 2 * 4 * kMaxWarps;
        g.scan = ip, ip += 4 * kMaxWarps;
        g.sctl = ip, ip += 16;
        g.heap = ip, ip += kHeapCap;
        g.wcnt = ip, ip += kWin;
        g.wcand = ip, ip += 4 * kWin;
        g.wgid = ip, ip += 4 * kWin;
        g.winert = ip, ip += 4 * kWin;
        g.prof = (Bt.prof && g.blk == 0) ? reinterpret_cast<long long *>(ip) : nullptr;
    }
    const int M = g.M, Nn = g.Nn, tid = g.tid;

    if (g.blk != 0) { // workers: the job loop
        for (;;) {
            gsync(g, D);
            if (D.job[JI_TYPE] == J_EXIT) break;
            g.nr = D.job[JI_NR];
            g.S = (long long)((g.nr + 1) | 1);
            g.MW = (g.nr + 31) >> 5;
            exec_job(g, D, T, Bt);
            gsync(g, D);
        }
    } else {
        const long long max_pivots = Bt.max_pivots;
        for (long long lp = 0; lp < Bt.B; ++lp) {
            if (tid == 0) D.job[JI_A4] = (int)lp;
            dispatch(g, D, T, Bt, J_INIT);
            dispatch(g, D, T, Bt, J_INIT2);
            int status = DZ_OPTIMAL;
            long long pivots = 0, n_primal = 0;
            unsigned long long hash = 0xcbf29ce484222325ULL, n_upd = 0;
            bool handed_over = false;
            if (g.prof) {
                if (tid < 16) g.prof[tid] = 0;
                __syncthreads();
                g.t_last = clock64();
            }
            long long dirty = 0; // doubles of W any solve of this LP has used
            g.stat_core = g.stat_real = g.stat_grid = 0;
            while (true) {
                // ---- status(), simplex.rs:274-306 ----
                dispatch(g, D, T, Bt, J_STATUS);
                int q0 = reduce_partials(g, D, 0, nullptr), p0 = reduce_partials(g, D, 2, nullptr);
                {
                    const int f1 = reduce_partials(g, D, 1, nullptr), f3 = reduce_partials(g, D, 3, nullptr);
                    if (f1 >= 0) { // the reference's reduce keeps a NaN first element
                        const double r = __ddiv_rn(-D.z[f1], D.zb[f1]);
                        if (r != r) q0 = f1;
                    }
                    if (f3 >= 0) {
                        const double r = __ddiv_rn(-D.x[f3], D.xb[f3]);
                        if (r != r) p0 = f3;
                    }
                }
                gtick(g, GP_STATUS);
                bool primal_step;
                double mu;
                if (q0 >= 0 && p0 >= 0) {
                    const double primal = __ddiv_rn(-D.x[p0], D.xb[p0]);
                    const double dual = __ddiv_rn(-D.z[q0], D.zb[q0]);
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
                    mu = __ddiv_rn(-D.z[q0], D.zb[q0]);
                } else if (p0 >= 0) {
                    primal_step = false;
                    mu = __ddiv_rn(-D.x[p0], D.xb[p0]);
                } else {
                    status = DZ_BREAKDOWN;
                    break;
                }
                if (pivots >= max_pivots) {
                    status = DZ_PIVOT_CAP;
                    break;
                }
                // ---- the coupled core of this basis ----
                dispatch(g, D, T, Bt, J_LISTS_COUNT);
                {
                    int run[4] = {0, 0, 0, 0};
                    for (int b = 0; b < g.nblocks; ++b) { // every master thread the same short scan
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (tid == 0) D.bcnt[(4 + q) * g.nblocks + b] = run[q];
                            run[q] += D.bcnt[q * g.nblocks + b];
                        }
                    }
                    const int nrow = run[0], ncol = run[1];
                    g.nse[0] = run[2];
                    g.nse[1] = run[3];
                    if (nrow != ncol) {
                        handed_over = true;
                        break;
                    }
                    g.nr = nrow;
                    g.S = (long long)((nrow + 1) | 1);
                    g.MW = (nrow + 31) >> 5;
                    if ((long long)nrow * g.S > D.w_cap) {
                        handed_over = true;
                        break;
                    }
                    dirty = max(dirty, (long long)nrow * g.S);
                    g.stat_core += 2 * (unsigned long long)nrow * (unsigned long long)g.S; // two solves
                }
                dispatch(g, D, T, Bt, J_LISTS_WRITE);
                gtick(g, GP_LISTS);
                int p = p0, q = q0;
                bool failed = false;
                for (int pass = 0; pass < 2; ++pass) {
                    const bool transposed = (pass == 0) != primal_step;
                    if (!grid_solve(g, D, T, Bt, transposed, transposed ? p : D.nb[q], lp)) {
                        handed_over = true;
                        break;
                    }
                    if (transposed) {
                        dispatch(g, D, T, Bt, J_PRICE);
                        gtick(g, GP_PRICE);
                    }
                    if (pass == 0) {
                        if (tid == 0) D.jobd[JD_0] = mu;
                        dispatch(g, D, T, Bt, J_RATIO, primal_step ? 1 : 0);
                        const int r = reduce_partials(g, D, 0, nullptr);
                        gtick(g, GP_RATIO);
                        if (primal_step) {
                            p = r;
                            if (p < 0) {
                                status = DZ_UNBOUNDED;
                                failed = true;
                            }
                        } else {
                            q = r;
                            if (q < 0) {
                                status = DZ_INFEASIBLE;
                                failed = true;
                            }
                        }
                        if (failed) break;
                    }
                }
                if (failed || handed_over) break;
                // ---- Simplex::pivot, simplex.rs:253-268 ----
                const int leaving = D.bas[p], entering = D.nb[q];
                double t, s, t_bar, s_bar;
                {
                    const double xp = D.x[p], dxp = D.dxv[p], zq = D.z[q], dzq = D.dzv[q];
                    const double xbp = D.xb[p], zbq = D.zb[q];
                    t = (xp == 0.0 && dxp == 0.0) ? 0.0 : __ddiv_rn(xp, dxp);
                    s = (zq == 0.0 && dzq == 0.0) ? 0.0 : __ddiv_rn(zq, dzq);
                    t_bar = (xbp == 0.0 && dxp == 0.0) ? 0.0 : __ddiv_rn(xbp, dxp);
                    s_bar = (zbq == 0.0 && dzq == 0.0) ? 0.0 : __ddiv_rn(zbq, dzq);
                }
                if (!(isfinite(t) && isfinite(s) && isfinite(t_bar) && isfinite(s_bar))) {
                    status = DZ_BREAKDOWN; // safe_divide assert, simplex.rs:466
                    break;
                }
                __syncthreads();
                if (tid == 0) {
                    D.jobd[JD_0] = t;
                    D.jobd[JD_1] = t_bar;
                    D.jobd[JD_2] = s;
                    D.jobd[JD_3] = s_bar;
                }
                dispatch(g, D, T, Bt, J_VECUPD, p, q, leaving, entering);
                n_upd += 4ull * (M + Nn);
                if (tid == 0) { // swap, simplex.rs:239-251
                    const int srl = T.slack_row[leaving], sre = T.slack_row[entering];
                    D.bas[p] = entering;
                    D.nb[q] = leaving;
                    D.where[entering] = -1 - p;
                    D.where[leaving] = q;
                    if (srl >= 0) D.spos[srl] = -1;
                    if (sre >= 0) D.spos[sre] = p;
                    D.srow[p] = sre;
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
                __syncthreads();
                gtick(g, GP_VECUPD);
            }
            if (g.prof) {
                __syncthreads();
                if (tid < 16) Bt.prof[(size_t)lp * 16 + tid] = g.prof[tid];
            }
            // leave the working matrix all-zero: the general kernel's interval mode, which
            // takes over a handed-over LP in the same workspace, expects that
            __syncthreads();
            if (tid == 0) D.jobd[JD_1] = (double)dirty;
            dispatch(g, D, T, Bt, J_ZERO, 0, 0, 1);
            if (handed_over) {
                __syncthreads();
                if (tid == 0) g.sctl[4] = (int)atomicAdd(Bt.exo_count, 1u);
                __syncthreads();
                const int slot = g.sctl[4];
                unsigned char *st = Bt.exo_state + (size_t)slot * Bt.exo_stride;
                double *sd = reinterpret_cast<double *>(st);
                for (int i = tid; i < M; i += g.NT) {
                    sd[i] = D.x[i];
                    sd[M + i] = D.xb[i];
                }
                for (int i = tid; i < Nn; i += g.NT) {
                    sd[2 * M + i] = D.z[i];
                    sd[2 * M + Nn + i] = D.zb[i];
                }
                long long *sl = reinterpret_cast<long long *>(sd + 2 * M + 2 * Nn);
                int *si = reinterpret_cast<int *>(sl + 3);
                for (int i = tid; i < M; i += g.NT) si[i] = D.bas[i];
                for (int i = tid; i < Nn; i += g.NT) si[M + i] = D.nb[i];
                if (tid == 0) {
                    Bt.exo_list[slot] = (int)lp;
                    sl[0] = pivots;
                    sl[1] = n_primal;
                    sl[2] = (long long)hash;
                }
            } else {
                dispatch(g, D, T, Bt, J_OBJ);
                const unsigned long long core_doubles = g.stat_core, real_steps = g.stat_real, grid_steps = g.stat_grid;
                if (tid == 0) {
                    const double *__restrict__ theta = Bt.theta + (size_t)lp * Bt.n_theta;
                    double obj = 0.0;
                    for (int p = 0; p < M; ++p) obj = __dadd_rn(obj, D.ycore[p]);
                    obj = __dadd_rn(load_ref(theta, T.c0_ref), obj);
                    Bt.status[lp] = status;
                    Bt.pivots[lp] = (int)pivots;
                    Bt.n_primal[lp] = (int)n_primal;
                    Bt.trace_hash[lp] = hash;
                    Bt.objective[lp] = obj;
                    if (Bt.work) {
                        double *w = Bt.work + (size_t)lp * 8;
                        atomicAdd(&w[3], (double)n_upd);
                        w[4] = (double)core_doubles;
                        w[5] = (double)real_steps;
                        w[6] = (double)grid_steps;
                    }
                }
            }
            __syncthreads();
        }
        dispatch(g, D, T, Bt, J_EXIT);
    }
    if (Bt.work && Bt.B == 1) { // executed flop counts (attributed to the LP when there is one)
        double *w = Bt.work;
        if (g.n_lu) atomicAdd(&w[0], (double)g.n_lu);
        if (g.n_solve) atomicAdd(&w[1], (double)g.n_solve);
        if (g.n_price) atomicAdd(&w[2], (double)g.n_price);
    }
}

} // namespace

// Workspace layout of the grid kernel for an m_int x n_int template with nnz entries.
// Returns the number of bytes; fills `d` with pointers relative to `base` (may be null
// to size only).
size_t grid_workspace(int M, int Nn, long long nnz, int nblocks, long long w_cap_doubles, unsigned char *base,
                      GridDev *d) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *p = base ? base + off : nullptr;
        off = (off + bytes + 255) & ~(size_t)255;
        return p;
    };
    const size_t Ms = (size_t)M, Ns = (size_t)Nn;
    GridDev g{};
    g.x = (double *)take(8 * Ms);
    g.xb = (double *)take(8 * Ms);
    g.dxv = (double *)take(8 * Ms);
    g.vv = (double *)take(8 * Ms);
    g.ycore = (double *)take(8 * Ms);
    g.z = (double *)take(8 * Ns);
    g.zb = (double *)take(8 * Ns);
    g.dzv = (double *)take(8 * Ns);
    g.lval = (double *)take(8 * (size_t)std::max<long long>(nnz, 1));
    g.pkey = (double *)take(8 * 4 * (size_t)nblocks);
    g.jobd = (double *)take(8 * JD_WORDS);
    g.bas = (int *)take(4 * Ms);
    g.nb = (int *)take(4 * Ns);
    g.rowAt = (int *)take(4 * Ms);
    g.posOf = (int *)take(4 * Ms);
    g.rowcnt = (int *)take(4 * Ms);
    g.srow = (int *)take(4 * Ms);
    g.spos = (int *)take(4 * Ms);
    g.rmapR = (int *)take(4 * Ms);
    g.rlist = (int *)take(4 * Ms);
    g.pmap = (int *)take(4 * Ms);
    g.plist = (int *)take(4 * Ms);
    g.pivr = (int *)take(4 * Ms);
    g.pend = (int *)take(4 * Ms);
    g.list = (int *)take(4 * (Ms + 32));
    g.pcols = (int *)take(4 * (Ms + 32));
    g.pwq = (int *)take(4 * (Ms / 32 + 32));
    g.pwb = (unsigned *)take(4 * (Ms / 32 + 32));
    g.pvals = (double *)take(8 * (Ms + 32));
    g.rlast = (int *)take(4 * Ms);
    g.done = (int *)take(4 * Ms);
    g.where = (int *)take(4 * (Ms + Ns));
    g.pidx = (int *)take(4 * 4 * (size_t)nblocks);
    g.bcnt = (int *)take(4 * 8 * (size_t)nblocks);
    g.seu = (int *)take(4 * Ms);
    g.set_ = (int *)take(4 * Ms);
    g.job = (int *)take(4 * JI_WORDS);
    g.bar = (unsigned *)take(4 * 4);
    g.rmask = (unsigned *)take(4 * Ms * (size_t)((M + 31) / 32));
    g.W = (double *)take(8 * (size_t)w_cap_doubles);
    g.w_cap = w_cap_doubles;
    if (d) *d = g;
    return off;
}

size_t grid_smem_bytes() {
    return (size_t)kBufTerms * 8 + kWin * 8 + 2 * 4 * kMaxWarps * 8 + (2 * 4 * kMaxWarps + 4 * kMaxWarps + 16 + 13 * kWin + kHeapCap) * 4 + 16 * 8 + 16;
}

int launch_grid(const TemplateDev &T, const BatchDev &Bt, const GridDev &D, const LaunchPlan &plan, void *stream,
                std::string *err) {
    cudaStream_t st = (cudaStream_t)stream;
#ifdef DZ_EMU
    (void)st;
    (void)err;
    return emu::launch_coop(dz_grid_kernel, plan.grid, plan.block, (size_t)plan.smem_bytes, T, Bt, D) == 0 ? DZ_OK
                                                                                                           : DZ_ERR_CUDA;
#else
    void *args[] = {(void *)&T, (void *)&Bt, (void *)&D};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)dz_grid_kernel, dim3(plan.grid), dim3(plan.block), args,
                                                (size_t)plan.smem_bytes, st);
    if (e != cudaSuccess) {
        *err = std::string("dz_grid_kernel cooperative launch: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    return DZ_OK;
#endif
}

} // namespace dz
