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
    J_BACK_PREP, J_YSCATTER, J_PRICE, J_RATIO, J_VECUPD, J_OBJ
};
// job descriptor words (ints) and doubles
enum { JI_TYPE = 0, JI_A0, JI_A1, JI_A2, JI_A3, JI_A4, JI_NR, JI_EXOTIC, JI_NLIST, JI_WORDS = 16 };
enum { JD_0 = 0, JD_1, JD_2, JD_3, JD_WORDS = 8 };
constexpr int kBufTerms = 2048; // ordered products of one back-substitution row per round (shared memory)

struct G {
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
    int parity;
    unsigned long long n_lu, n_solve, n_price;
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

__device__ __forceinline__ double fast_div(double v, double pv) {
    return (pv == 1.0) ? v : ((pv == -1.0) ? -v : __ddiv_rn(v, pv));
}

// Elimination update of the candidate rows list[first], list[first + stride], ... (one warp
// per row): l = a_ik / pivot, then a_ij -= l * a_kj over the nonzero pattern of the pivot row
// right of the pivot column, the right-hand side included (linalg.rs:118-124, :288-290).
__device__ __forceinline__ void update_rows(G &g, const GridDev &D, int first, int stride, int n_list, int k, int cc, int pr,
                                            double pv) {
    const int nr = g.nr, MW = g.MW, lane = g.lane;
    const long long S = g.S;
    const double *__restrict__ prow = D.W + (size_t)pr * S;
    const unsigned *__restrict__ pmask = D.rmask + (size_t)pr * MW;
    const double urhs = prow[nr];
    const int q0 = (cc + 1) >> 5;
    bool bad = !isfinite(urhs);
    for (int e = first; e < n_list; e += stride) {
        const int i = D.list[e];
        if (i == pr) continue;
        double *__restrict__ row = D.W + (size_t)i * S;
        unsigned *__restrict__ imask = D.rmask + (size_t)i * MW;
        const double v = row[cc];
        const double l = fast_div(v, pv);
        bad = bad || !isfinite(l);
        unsigned long long cnt = 0;
        for (int wb = q0; wb < MW; wb += 32) {
            const int q = wb + lane;
            unsigned word = (q < MW) ? pmask[q] : 0u;
            if (q == q0) {
                const int lo = cc + 1 - 32 * q0;
                word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
            }
            if (word && q < MW) imask[q] |= word; // fill pattern
            unsigned nzw = __ballot_sync(kFull, word != 0u);
            while (nzw) {
                const int wq = __ffs(nzw) - 1;
                nzw &= nzw - 1;
                const unsigned bits = __shfl_sync(kFull, word, wq);
                if ((bits >> lane) & 1u) {
                    const int j = 32 * (wb + wq) + lane;
                    const double u = prow[j];
                    bad = bad || !isfinite(u);
                    row[j] = __dsub_rn(row[j], __dmul_rn(l, u));
                }
                cnt += 2ull * __popc(bits);
            }
        }
        if (lane == 0) {
            if (urhs != 0.0) {
                row[nr] = __dsub_rn(row[nr], __dmul_rn(l, urhs));
                cnt += 2;
            }
            g.n_lu += cnt + 1;
        }
    }
    if (__ballot_sync(kFull, bad) && lane == 0) D.job[JI_EXOTIC] = 1;
    (void)k;
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
        // The coupled core (see dz_core.cu core_lists): CTA b owns the index range [b*chunk,
        // (b+1)*chunk), so that the concatenation of the CTAs' lists is in index order.
        const int chunk = (M + g.nblocks - 1) / g.nblocks;
        const int lo = g.blk * chunk, hi = min(M, lo + chunk);
        const unsigned lt = (1u << g.lane) - 1u;
        int offr = type == J_LISTS_WRITE ? D.bcnt[2 * g.nblocks + g.blk] : 0;
        int offc = type == J_LISTS_WRITE ? D.bcnt[3 * g.nblocks + g.blk] : 0;
        int nrow = 0, ncol = 0;
        for (int base = lo; base < hi; base += g.NT) {
            const int i = base + g.tid;
            bool pr = false, pc = false;
            if (i < hi) {
                pr = D.rowcnt[i] > 0;
                const int sr = D.srow[i];
                pc = sr < 0 || D.rowcnt[sr] > 0;
            }
            const unsigned mr = __ballot_sync(kFull, pr), mc = __ballot_sync(kFull, pc);
            if (g.lane == 0) {
                g.scan[g.warp] = __popc(mr);
                g.scan[kMaxWarps + g.warp] = __popc(mc);
            }
            __syncthreads();
            int wr = 0, wc = 0, tr = 0, tc = 0;
            for (int w = 0; w < g.NW; ++w) {
                const int a = g.scan[w], b = g.scan[kMaxWarps + w];
                if (w < g.warp) {
                    wr += a;
                    wc += b;
                }
                tr += a;
                tc += b;
            }
            if (type == J_LISTS_WRITE && i < hi) {
                const int ir = pr ? offr + nrow + wr + __popc(mr & lt) : -1;
                D.rmapR[i] = ir;
                if (pr) D.rlist[ir] = i;
                const int ic = pc ? offc + ncol + wc + __popc(mc & lt) : -1;
                D.pmap[i] = ic;
                if (pc) D.plist[ic] = i;
            }
            nrow += tr;
            ncol += tc;
            __syncthreads();
        }
        if (type == J_LISTS_COUNT && g.tid == 0) {
            D.bcnt[g.blk] = nrow;
            D.bcnt[g.nblocks + g.blk] = ncol;
        }
        break;
    }
    case J_ZERO: { // a0 = rows of W and of rmask to clear; also the tables of a fresh elimination
        // a2 = 1: clear everything any solve of this LP may have touched (count in jobd[JD_1])
        const long long nw = a2 ? (long long)D.jobd[JD_1] : (long long)a0 * g.S;
        const long long nm = a2 ? 0 : (long long)a0 * g.MW;
        for (long long e = g.gtid; e < nw; e += g.GT) D.W[e] = 0.0;
        for (long long e = g.gtid; e < nm; e += g.GT) D.rmask[e] = 0u;
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
            for (int e = T.col_ptr[col] + g.lane; e < T.col_ptr[col + 1]; e += 32) {
                const double val = D.lval[e];
                if (val == 0.0) continue; // exact zeros are not stored (linalg.rs:261)
                const int ir = D.rmapR[T.row_idx[e]];
                const int wi = transposed ? cc : ir, wj = transposed ? ir : cc;
                D.W[(size_t)wi * S + wj] = val;
                atomicOr(&D.rmask[(size_t)wi * MW + (wj >> 5)], 1u << (wj & 31));
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
        update_rows(g, D, g.gwarp, g.GW, D.job[JI_NLIST], a0, a1, a2, D.jobd[JD_0]);
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
                    if (!isfinite(yi)) D.job[JI_EXOTIC] = 2; // non-finite component: see the master
                    g.n_solve += 1;
                }
            }
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

__device__ __forceinline__ void g_swap_pos(const GridDev &D, int k, int mu) {
    const int rk = D.rowAt[k], rm = D.rowAt[mu];
    D.rowAt[k] = rm;
    D.rowAt[mu] = rk;
    D.posOf[rm] = k;
    D.posOf[rk] = mu;
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
        D.job[14] = (int)S;  // workers recompute S and MW from nr; kept for debugging
    }
    dispatch(g, D, T, Bt, J_ZERO, nr, transposed ? 1 : 0);
    dispatch(g, D, T, Bt, J_SCATTER, transposed ? 1 : 0, arg);

    // ---- elimination (master; dense steps go to the grid) ----
    int k = 0;
    for (;;) {
        if (warp == 0) {
            int irregular = 0;
            for (;;) { // bookkeeping steps, see dz_core.cu
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
                int adv = 0;
                if (lane == 0) {
                    const int cc = cmap[k];
                    const int u = transposed ? (cc < 0 ? D.spos[k] : -1) : D.srow[k];
                    if (cc < 0) {
                        const int pu = u >= 0 ? D.posOf[u] : -1;
                        if (pu < k) {
                            irregular = 1;
                        } else {
                            g_swap_pos(D, k, pu);
                            adv = 1;
                        }
                    } else if (u >= 0 && D.posOf[u] >= k) {
                        g_swap_pos(D, k, D.posOf[u]);
                        D.pivr[cc] = rmap[u];
                        adv = 1;
                    }
                }
                adv = __shfl_sync(kFull, adv, 0);
                irregular = __shfl_sync(kFull, irregular, 0);
                __syncwarp();
                if (!adv) break;
                ++k;
            }
            if (lane == 0) {
                g.sctl[0] = k;
                g.sctl[1] = irregular;
                g.sctl[2] = 0; // candidate rows with a nonzero in the pivot column
            }
        }
        __syncthreads();
        k = g.sctl[0];
        if (g.sctl[1]) return false;
        if (k >= M - 1) break;
        // pivot search in core column cc over the rows at positions >= k (linalg.rs:98-105):
        // largest |a_ik|, ties to the smallest position; rows with a nonzero entry are listed
        const int cc = cmap[k];
        Cand<1> cd;
        cd.key[0] = 0.0;
        cd.idx[0] = -1; // idx = position (unique per row)
        bool bad = false;
        for (int i = tid; i < nr; i += g.NT) {
            const int pos = D.posOf[rlist[i]];
            if (pos >= k) {
                const double v = D.W[(size_t)i * S + cc];
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
                g_swap_pos(D, k, ppos); // linalg.rs:107-114
                D.pivr[cc] = pr;
            }
        }
        g.stat_real += 1;
        if (pv != 0.0 && n_list > 1) {
            if (n_list <= 2 * g.NW) { // narrow step: the master's own warps, no grid barrier
                __syncthreads();
                update_rows(g, D, warp, g.NW, n_list, k, cc, pr, pv);
                __syncthreads();
            } else {
                if (tid == 0) {
                    D.job[JI_NLIST] = n_list;
                    D.jobd[JD_0] = pv;
                }
                dispatch(g, D, T, Bt, J_UPDATE, k, cc, pr);
                g.stat_grid += 1;
            }
            if (D.job[JI_EXOTIC] == 1) return false;
        }
        __syncthreads();
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
    bool nonfinite = D.job[JI_EXOTIC] == 2;
    for (int cc = nr - 1; cc >= 0; --cc) {
        if (!D.pend[cc]) continue;
        const int i = D.pivr[cc];
        const double *__restrict__ row = D.W + (size_t)i * S;
        const unsigned *__restrict__ mask = D.rmask + (size_t)i * MW;
        double s = row[nr];
        const double d = row[cc];
        const int q0 = (cc + 1) >> 5;
        // the CTA forms the products u_ij * y_j of the row's pattern in column order, up to
        // kBufTerms words' worth at a time; warp 0's lanes then subtract them one after the
        // other (every lane the same chain, so the result is in every lane)
        for (int qb = q0; qb < MW; qb += g.NT) {
            const int q = qb + tid;
            unsigned word = (q < MW) ? mask[q] : 0u;
            if (q == q0) {
                const int lo = cc + 1 - 32 * q0;
                word = lo >= 32 ? 0u : (word & ~((1u << lo) - 1u));
            }
            // exclusive prefix of the popcounts over the CTA
            const int pc = __popc(word);
            int incl = pc;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, off);
                if (lane >= off) incl += t;
            }
            if (lane == 31) g.scan[warp] = incl;
            __syncthreads();
            int base = 0, total = 0;
            for (int w = 0; w < g.NW; ++w) {
                const int a = g.scan[w];
                if (w < warp) base += a;
                total += a;
            }
            int at = base + incl - pc;
            // a thread's 32 columns may straddle the buffer: rounds of kBufTerms terms
            for (int r0 = 0; r0 < total; r0 += kBufTerms) {
                unsigned wbits = word;
                int pos = at;
                while (wbits) {
                    const int b = __ffs(wbits) - 1;
                    wbits &= wbits - 1;
                    if (pos >= r0 && pos < r0 + kBufTerms) {
                        const int j = 32 * q + b;
                        g.sbuf[pos - r0] = __dmul_rn(row[j], D.ycore[j]);
                    }
                    ++pos;
                }
                __syncthreads();
                if (warp == 0) {
                    const int n = min(kBufTerms, total - r0);
                    int t = 0;
                    for (; t + 4 <= n; t += 4) {
                        const double p0 = g.sbuf[t], p1 = g.sbuf[t + 1], p2 = g.sbuf[t + 2], p3 = g.sbuf[t + 3];
                        s = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(s, p0), p1), p2), p3);
                    }
                    for (; t < n; ++t) s = __dsub_rn(s, g.sbuf[t]);
                    if (lane == 0) g.n_solve += 2ull * n;
                }
                __syncthreads();
            }
            __syncthreads(); // scan[] is rewritten by the next batch of words
        }
        if (warp == 0) {
            const double yi = (d == 1.0) ? s : __ddiv_rn(s, d);
            if (lane == 0) {
                D.ycore[cc] = yi;
                g.sctl[3] = isfinite(yi) ? 0 : 1;
                g.n_solve += 1;
            }
        }
        __syncthreads();
        nonfinite = nonfinite || g.sctl[3] != 0;
        __syncthreads();
    }
    dispatch(g, D, T, Bt, J_YSCATTER, transposed ? 1 : 0);
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
        g.red_key = dp, dp += 2 * 4 * kMaxWarps;
        int *ip = reinterpret_cast<int *>(dp);
        g.red_idx = ip, ip += 2 * 4 * kMaxWarps;
        g.scan = ip, ip += 2 * kMaxWarps;
        g.sctl = ip, ip += 16;
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
                    int nrow = 0, ncol = 0;
                    for (int b = 0; b < g.nblocks; ++b) { // every master thread the same short scan
                        if (tid == 0) {
                            D.bcnt[2 * g.nblocks + b] = nrow;
                            D.bcnt[3 * g.nblocks + b] = ncol;
                        }
                        nrow += D.bcnt[b];
                        ncol += D.bcnt[g.nblocks + b];
                    }
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
                int p = p0, q = q0;
                bool failed = false;
                for (int pass = 0; pass < 2; ++pass) {
                    const bool transposed = (pass == 0) != primal_step;
                    if (!grid_solve(g, D, T, Bt, transposed, transposed ? p : D.nb[q], lp)) {
                        handed_over = true;
                        break;
                    }
                    if (transposed) dispatch(g, D, T, Bt, J_PRICE);
                    if (pass == 0) {
                        if (tid == 0) D.jobd[JD_0] = mu;
                        dispatch(g, D, T, Bt, J_RATIO, primal_step ? 1 : 0);
                        const int r = reduce_partials(g, D, 0, nullptr);
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
    g.where = (int *)take(4 * (Ms + Ns));
    g.pidx = (int *)take(4 * 4 * (size_t)nblocks);
    g.bcnt = (int *)take(4 * 4 * (size_t)nblocks);
    g.job = (int *)take(4 * JI_WORDS);
    g.bar = (unsigned *)take(4 * 4);
    g.rmask = (unsigned *)take(4 * Ms * (size_t)((M + 31) / 32));
    g.W = (double *)take(8 * (size_t)w_cap_doubles);
    g.w_cap = w_cap_doubles;
    if (d) *d = g;
    return off;
}

size_t grid_smem_bytes() { return (size_t)kBufTerms * 8 + 2 * 4 * kMaxWarps * 8 + (2 * 4 * kMaxWarps + 2 * kMaxWarps + 16) * 4 + 16; }

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
