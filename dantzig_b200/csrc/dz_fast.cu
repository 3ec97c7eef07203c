// dz_fast.cu -- OPT-IN fast numerics for batches of small LPs (sm_100a).  NOT the parity path.
//
// dz_options.numerics == DZ_NUMERICS_FAST selects this kernel; the default (EXACT) never
// reaches it.  It runs the same algorithm as the reference -- the parametric self-dual
// simplex of /root/reference/src/simplex.rs:274-343 with the same entering / leaving rules
// and the same "first index wins" tie-breaks (:423-461) and a FRESH factorisation of the
// basis on every pivot (no factor reuse between pivots, like :226-236) -- but it gives up
// the reference's floating-point operation order:
//
//   * ONE factorisation per pivot serves both FTRAN (simplex.rs:226-229) and BTRAN
//     (:231-236); the reference factorises B and an explicit B^T separately.
//   * Fused multiply-add everywhere; reciprocal-multiply instead of division in the
//     elimination.
//   * Tolerance-guarded tests where the reference has none: the ratio test ignores direction
//     components below 1e-9 and treats a negative perturbed value as zero (the reference
//     tests `ratio > 0.0` only, simplex.rs:455), and optimality is accepted at mu* <= 1e-9
//     (the reference: 1e-12, :283).  On well-posed LPs the pivots are the reference's; on
//     the LPs where the reference's arithmetic ends in a false "infeasible"/"unbounded" or
//     in the safe_divide panic (about half of config 1's instances) this path still ends at
//     the optimum (checked against HiGHS in tests/test_fast_mode.py).
//   * The factorisation is restricted to the part of the basis that is not an identity
//     block.  With the basis columns split into slacks (unit columns) and structurals,
//     and the rows into S (slack basic) and R (slack nonbasic),
//
//         B = [ K  0 ]   rows R            K = A[R, structural basis columns]   (k x k)
//             [ F  I ]   rows S            F = A[S, structural basis columns]
//
//     so  B d = a   is   K d_C = a_R,  d_S = a_S - F d_C,   and
//         B^T v = e_p is  v_S = (e_p)_S, K^T v_R = (e_p)_C - F^T v_S.
//     K (config 5: k ~ 68 of m_int = 192; config 2: ~33 of 96) lives in shared memory.
//   * K is inverted in place by Gauss-Jordan elimination with partial pivoting (largest
//     magnitude, smallest row index on ties), rows never physically swapped and never
//     scaled: storage column j is reused for the identity column of step j's pivot row,
//     so K^-1[a][b] = rinv[a] * U[pr[a]][prinv[b]].  Both solves are then dense
//     matrix-vector products over all threads: no triangular-solve dependency chain.
//     For k <= 32 / 64 / 128 the elimination is tiled over one / four / sixteen warps
//     (thread (ty, tx): rows ty + TY * a, columns tx + 16 * b, the pivot row's entries in
//     registers, warp 0 searching), two barriers per step; larger k, or a K that does not
//     fit shared memory, falls back to a generic loop.
//
// Results are deterministic (fixed summation orders, no floating-point atomics) but not
// bit-identical to the reference: on well-posed LPs status and objective agree with the
// exact path (tests/test_fast_mode.py compares at 1e-9 and reports pivot-count deltas);
// on LPs where the reference's tolerance-free ratio tests break down the two may part.
//
// One CTA per LP (persistent, work queue), 128 threads for m_int <= 128, else 512.

#include "dz_device.cuh"

#include <algorithm>
#include <cstdio>
#include <string>

namespace dz {

namespace {

enum { FC_LP = 0, FC_WORDS = 8 };
constexpr double kFastPivTol = 1e-9; // ratio tests of the fast path ignore |direction| below this
constexpr double kFastOptTol = 1e-9; // ... and its optimality test accepts a parameter mu* up to this

// Slots of the optional per-LP cycle profile (BatchDev::prof, 16 per LP).
enum { FP_STATUS = 0, FP_LISTS, FP_BUILD, FP_GJ, FP_FTRAN, FP_BTRAN, FP_PRICE, FP_RATIO, FP_UPDATE,
       FP_K_SUM, FP_SOLVES, FP_ROWS_UPD, FP_K_GLOBAL, FP_K_MAX };

struct Fast {
    int M, Nn, NT, NW, tid, lane, warp;
    long long *prof; // shared-memory accumulators, or null
    long long t_last;
    // per-LP state, shared memory
    double *x, *xb, *dxv, *vv, *z, *zb, *dzv;
    double *bK, *tK, *rinv, *uK; // [M] work vectors of the two solves
    double *acc;                 // [NW][M] per-warp partial sums of F d_C (fast_ftran)
    double *red_key;
    int *red_idx;
    int parity;
    int *bas, *nb;
    int *where;         // [Nint] column -> basis position p >= 0, or -1 - (nonbasic slot)
    int *srow;          // [M] position -> row of the slack that sits there, -1 if structural
    int *spos;          // [M] row -> position of its slack, -1 if the slack is nonbasic
    int *rmapK, *rlistK; // row -> row of K / back
    int *cmapK, *clistK; // basis position -> column of K / back
    int *pr, *prinv;    // elimination step -> its pivot row of K / back
    int *pdone;         // row of K already used as a pivot row
    int *kref;          // [M] column of K -> signed theta reference of its row-0 entry (dense-rows templates)
    int *scan, *ctl;
    // the k x k working matrix
    double *Ks;
    int capK;
    double *Kg;
    double *K;
    int k, S;
    bool blocked; // blocked tensor-core elimination (opt-in A/B switch) instead of the step-by-step tiled loop
    unsigned long long n_lu, n_solve, n_price;
};

__device__ __forceinline__ void ftick(Fast &c, int slot) {
    if (c.prof && c.tid == 0) {
        const long long now = clock64();
        c.prof[slot] += now - c.t_last;
        c.t_last = now;
    }
}

// find_first_pivot (simplex.rs:423-437) on both sides, one barrier.
__device__ __forceinline__ void fast_find_first_both(Fast &c, int &q0, int &p0) {
    Cand<4> cd;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        cd.key[n] = 0.0;
        cd.idx[n] = -1;
    }
    for (int k = c.tid; k < c.Nn; k += c.NT) {
        const double yb = c.zb[k];
        if (yb > 0.0) {
            const double ratio = -c.z[k] / yb;
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
            const double ratio = -c.x[k] / yb;
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
        const double r = -c.z[cd.idx[1]] / c.zb[cd.idx[1]];
        if (r != r) q0 = cd.idx[1];
    }
    p0 = cd.idx[2];
    if (cd.idx[3] >= 0) {
        const double r = -c.x[cd.idx[3]] / c.xb[cd.idx[3]];
        if (r != r) p0 = cd.idx[3];
    }
}

// find_second_pivot (simplex.rs:439-461).
__device__ __forceinline__ int fast_find_second(Fast &c, double mu, const double *y, const double *yb,
                                                const double *dy, int len) {
    Cand<1> cd;
    cd.key[0] = 0.0;
    cd.idx[0] = -1;
    for (int k = c.tid; k < len; k += c.NT) {
        // tolerance-guarded (the reference tests ratio > 0.0 and nothing else, simplex.rs:455): the
        // perturbed value y + mu * ybar is nonnegative by the method's invariant, so a negative one is
        // rounding noise and counts as zero, and a direction component within kFastPivTol of zero does
        // not block the step.  On well-posed LPs this selects the reference's index.
        const double d = dy[k];
        if (d > kFastPivTol) {
            double denom = fma(mu, yb[k], y[k]);
            if (denom < 0.0) denom = 0.0;
            const double ratio = d / denom; // +inf on a degenerate step: first index wins
            if (beats(ratio, k, cd.key[0], cd.idx[0])) {
                cd.key[0] = ratio;
                cd.idx[0] = k;
            }
        }
    }
    block_argmax<1>(cd, c.red_key, c.red_idx, c.parity, c.NW, c.tid, false);
    return cd.idx[0];
}

// Rows R (slack nonbasic) and structural basis positions, in ascending order: the rows and
// columns of K.  The two counts are equal for every basis (m_int rows, m_int positions,
// one basic slack per row of S).
__device__ __forceinline__ void fast_lists(Fast &c) {
    const int M = c.M;
    const unsigned lt = (1u << c.lane) - 1u;
    int nrow = 0, ncol = 0;
    for (int base = 0; base < M; base += c.NT) {
        const int i = base + c.tid;
        const bool pr = i < M && c.spos[i] < 0, pc = i < M && c.srow[i] < 0;
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
            c.rmapK[i] = ir;
            if (pr) c.rlistK[ir] = i;
            const int ic = pc ? offc + __popc(mc & lt) : -1;
            c.cmapK[i] = ic;
            if (pc) c.clistK[ic] = i;
        }
        __syncthreads();
    }
    c.k = nrow; // == ncol
}

// Per-warp candidate of the pivot search in column j: largest |K[i][j]| over this warp's rows
// that are not pivot rows yet (`skip` is the row chosen in the step being finished), smallest
// row on ties.  Lane 0 publishes it in the reduction slots of `parity`.
__device__ __forceinline__ void gj_candidate(Fast &c, int j, int skip) {
    const int i = c.warp + c.lane * c.NW; // k <= 32 * NW by the launch rule
    double key = -1.0;
    int idx = -1;
    if (i < c.k && i != skip && !c.pdone[i]) {
        key = fabs(c.K[i * c.S + j]);
        idx = i;
        if (!(key == key)) key = 1.7976931348623157e308; // a NaN column breaks down below
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double k2 = __shfl_xor_sync(kFull, key, off);
        const int i2 = __shfl_xor_sync(kFull, idx, off);
        if (beats(k2, i2, key, idx)) {
            key = k2;
            idx = i2;
        }
    }
    if (c.lane == 0) {
        c.red_key[c.parity * kMaxWarps + c.warp] = key;
        c.red_idx[c.parity * kMaxWarps + c.warp] = idx;
    }
}

// Generic in-place Gauss-Jordan inversion of K where it lies (shared memory or the HBM workspace):
// any k, two CTA barriers per step (the pivot search of column j+1 is folded into the update of
// step j).  The fallback for k beyond the register-tiled classes below.
__device__ __forceinline__ bool gj_generic(Fast &c) {
    const int k = c.k, S = c.S, tid = c.tid, lane = c.lane, warp = c.warp, NW = c.NW;
    double *K = c.K; // never __restrict__/const: K may live in the HBM workspace and is rewritten all the time
    for (int i = tid; i < k; i += c.NT) c.pdone[i] = 0;
    __syncthreads();
    int par = 0;
    c.parity = par; // gj_candidate reads it
    gj_candidate(c, 0, -1);
    bool ok = true;
    unsigned long long rows_upd = 0;
    for (int j = 0; j < k; ++j) {
        __syncthreads(); // the updates of step j-1 and the candidates of column j are visible
        double bk = c.red_key[par * kMaxWarps];
        int ip = c.red_idx[par * kMaxWarps];
        for (int w = 1; w < NW; ++w) {
            const double k2 = c.red_key[par * kMaxWarps + w];
            const int i2 = c.red_idx[par * kMaxWarps + w];
            if (beats(k2, i2, bk, ip)) {
                bk = k2;
                ip = i2;
            }
        }
        par ^= 1;
        c.parity = par;
        if (!(bk > 0.0) || !(bk < 1.0e300)) { // every thread sees the same candidate
            ok = false;
            break;
        }
        const double r = 1.0 / K[ip * S + j];
        __syncthreads(); // everyone has the pivot before the pivot row's owner overwrites it below
        if (tid == 0) {
            c.pr[j] = ip;
            c.prinv[ip] = j;
            c.rinv[j] = r;
            c.pdone[ip] = 1;
        }
        const double *prow = K + ip * S;
        for (int i = warp; i < k; i += NW) {
            double *row = K + i * S;
            if (i == ip) {
                __syncwarp();
                if (lane == 0) row[j] = 1.0; // the identity column of this pivot row takes column j's place
                continue;
            }
            const double f = row[j];
            if (f == 0.0) continue; // warp-uniform
            const double g = f * r;
            __syncwarp(); // every lane has read row[j]
            for (int cc = lane; cc < k; cc += 32) row[cc] = (cc == j) ? -g : fma(-g, prow[cc], row[cc]);
            ++rows_upd;
        }
        __syncwarp();
        if (j + 1 < k) gj_candidate(c, j + 1, ip);
    }
    __syncthreads();
    if (lane == 0) c.n_lu += rows_upd * (unsigned long long)(2 * k + 1);
    return ok;
}

// 1 / x to ~1 ulp without the division sequence: hardware approximation + two Newton steps.
__device__ __forceinline__ double fast_rcp(double x) {
#ifdef DZ_EMU
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#endif
}

// D = A * B + C on the FP64 tensor cores: one m8n8k4 tile per warp.  Fragments (PTX ISA, mma.m8n8k4
// .f64): with g = lane >> 2 and t = lane & 3, a = A[g][t], b = B[t][g], (c0, c1) = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b) {
#ifdef DZ_EMU // test-only restatement of the fragment layout on the SIMT emulator
    const int lane = (int)threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double d0 = c0, d1 = c1;
    for (int kk = 0; kk < 4; ++kk) {
        const double ak = __shfl_sync(kFull, a, g * 4 + kk);
        const double b0 = __shfl_sync(kFull, b, (2 * t) * 4 + kk);
        const double b1 = __shfl_sync(kFull, b, (2 * t + 1) * 4 + kk);
        d0 = fma(ak, b0, d0);
        d1 = fma(ak, b1, d1);
    }
    c0 = d0;
    c1 = d1;
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// BLOCKED in-place Gauss-Jordan inversion of K in shared memory (k <= 16 * CB), four elimination
// steps per pass over the matrix.  Within a panel of four columns every step is carried out on
// the panel's columns only: warp 0 searches the step's column (REDUX on the magnitude bits,
// smallest row on ties) and publishes the pivot row and the reciprocal of the pivot; then the
// threads form the multipliers (G, k x 4, stored negated), update the panel's four columns in
// place, and record the pivot row as it stands at that step outside the panel (U, 4 x columns,
// zero on the panel's columns).  After the four steps every warp applies the rank-4 update
// K += (-G) U to the whole matrix as 8 x 8 tiles on the FP64 TENSOR CORES (mma.m8n8k4.f64: the
// panel width is the instruction's k).  The k x k update runs once per four steps.  Same pivots as
// the step-by-step loop.  MEASURED (profiles/r02_fast_mode.md): no faster than the step-by-step
// loop at k ~ 34 / 68 (config 2: 24.9 against 26.0 kLP/s, config 5: 1.47 against 1.51 kLP/s) -- a step
// is a chain of short dependent phases (search ~900 cycles, panel update ~900, a quarter of the
// tile pass ~1000 at config 5), not the k x k traffic -- so it is an opt-in switch
// (dz_options.worker_warps == 2 together with DZ_NUMERICS_FAST), not the default.
template <int CB>
__device__ __forceinline__ void gj_blocked(double *K, const int k, const int tid, const int NT, double *Gbuf,
                                           double *Ubuf, double *colj, double *pub, int *pr, int *prinv,
                                           double *rinv, int *okflag, long long *prof) {
    long long tq = (prof && tid == 0) ? clock64() : 0;
    constexpr int S = 16 * CB + 1;
    const int lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const int ncol = ((k + 15) >> 4) << 4; // columns in use, padded (the padding columns are zero)
    unsigned long long pd0 = 0ull, pd1 = 0ull; // warp 0: rows already used as pivot rows
    bool ok = true;                            // warp 0
    for (int j0 = 0; j0 < k; j0 += 4) {
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const int col = j0 + s;
            if (warp == 0) {
                int p = -1;
                double r = 0.0;
                if (col < k && ok) {
                    double kb = 0.0;
                    int ib = 0x7fffffff;
#pragma unroll
                    for (int q = 0; q < (CB + 1) / 2; ++q) {
                        const int i = lane + 32 * q;
                        if (i < k) {
                            const double v0 = K[i * S + col];
                            colj[i] = v0;
                            const bool done = q < 2 ? ((pd0 >> (i & 63)) & 1ull) : ((pd1 >> (i & 63)) & 1ull);
                            double av = fabs(v0);
                            if (!(av == av)) av = 1.7976931348623157e308;
                            if (!done && av > kb) {
                                kb = av;
                                ib = i;
                            }
                        }
                    }
                    const unsigned hi = (unsigned)__double2hiint(kb), lo = (unsigned)__double2loint(kb);
                    const unsigned mh = __reduce_max_sync(kFull, hi);
                    const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
                    p = __reduce_min_sync(kFull, (hi == mh && lo == ml) ? ib : 0x7fffffff);
                    if ((mh | ml) == 0u || mh >= 0x7e000000u || p >= k) { // zero, huge or non-finite pivot
                        ok = false;
                        p = -1;
                    } else {
                        __syncwarp();
                        r = fast_rcp(colj[p]);
                        if (p < 64) pd0 |= 1ull << p; else pd1 |= 1ull << (p - 64);
                    }
                }
                if (lane == 0) {
                    pub[0] = r;
                    reinterpret_cast<int *>(pub + 1)[0] = p;
                    if (p >= 0) {
                        pr[col] = p;
                        prinv[p] = col;
                        rinv[col] = r;
                    }
                }
            }
            if (prof && tid == 0) {
                const long long t_ = clock64();
                prof[11] += t_ - tq;
                tq = t_;
            }
            __syncthreads();
            if (prof && tid == 0) {
                const long long t_ = clock64();
                prof[12] += t_ - tq;
                tq = t_;
            }
            {
                const double r = pub[0];
                const int p = reinterpret_cast<const int *>(pub + 1)[0];
                if (p >= 0) {
                    // the pivot row as it stands at this step, outside the panel's columns: the rows of
                    // the earlier steps of this panel have not been applied to it yet
                    for (int c = tid; c < ncol; c += NT) {
                        double u = 0.0;
                        if (c < j0 || c >= j0 + 4) {
                            u = K[p * S + c];
                            for (int t = 0; t < s; ++t) u = fma(Gbuf[p * 4 + t], Ubuf[t * ncol + c], u);
                        }
                        Ubuf[s * ncol + c] = u;
                    }
                    // multipliers; the panel's columns in place (column `col` becomes the identity
                    // column of the pivot row, with the multipliers negated)
                    for (int i = tid; i < k; i += NT) {
                        if (i == p) {
                            Gbuf[i * 4 + s] = 0.0;
                            K[i * S + col] = 1.0;
                        } else {
                            const double g = colj[i] * r;
                            Gbuf[i * 4 + s] = -g;
                            K[i * S + col] = -g;
                            if (g != 0.0) {
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    if (t != s && j0 + t < k) K[i * S + j0 + t] = fma(-g, K[p * S + j0 + t], K[i * S + j0 + t]);
                            }
                        }
                    }
                } else { // past the last column, or after a breakdown: a step that changes nothing
                    for (int c = tid; c < ncol; c += NT) Ubuf[s * ncol + c] = 0.0;
                    for (int i = tid; i < k; i += NT) Gbuf[i * 4 + s] = 0.0;
                }
            }
            __syncthreads();
            if (prof && tid == 0) {
                const long long t_ = clock64();
                prof[14] += t_ - tq;
                tq = t_;
            }
        }
        {
            const int nct = ncol >> 3, nrt = (k + 7) >> 3;
            const int g = lane >> 2, t = lane & 3;
            for (int tile = warp; tile < nrt * nct; tile += NW) {
                const int rt = tile / nct, ct = tile - rt * nct;
                const int row = 8 * rt + g, c = 8 * ct + 2 * t;
                const double a = row < k ? Gbuf[row * 4 + t] : 0.0;
                const double b = Ubuf[t * ncol + 8 * ct + g];
                double c0 = 0.0, c1 = 0.0;
                if (row < k) {
                    c0 = K[row * S + c];
                    c1 = K[row * S + c + 1];
                }
                dmma_m8n8k4(c0, c1, a, b);
                if (row < k) {
                    K[row * S + c] = c0;
                    K[row * S + c + 1] = c1;
                }
            }
        }
        __syncthreads();
        if (prof && tid == 0) {
            const long long t_ = clock64();
            prof[15] += t_ - tq;
            tq = t_;
        }
    }
    if (tid == 0) *okflag = ok ? 1 : 0;
}

// Tiled in-place Gauss-Jordan inversion of K in shared memory, for k <= 16 * CB.  The first
// TY * 16 threads work: thread (ty, tx) updates rows ty, ty + TY, ... and the CB columns
// tx + 16 * b, the pivot row's entries held in registers for the whole step (K is stored with
// the stride 16 * CB + 1 and zero padding columns, so the column loop needs no bounds).
// Per step: warp 0 searches column j (REDUX on the magnitude bits, smallest row on ties),
// snapshots it, and publishes the pivot row and the reciprocal of the pivot; one barrier;
// every thread updates its rows; one barrier.  TY == 2 is a single warp: __syncwarp instead
// of CTA barriers, and the other warps skip the call.  A singular K does not leave the loop
// early (the barriers stay matched); *okflag reports it.
template <int TY, int CB>
__device__ __forceinline__ void gj_tiled(double *K, const int k, const int tid, double *colj, double *pub,
                                         int *pr, int *prinv, double *rinv, int *okflag, long long *prof) {
    constexpr int S = 16 * CB + 1;
    long long tq = (prof && tid == 0) ? clock64() : 0;
    const int lane = tid & 31, warp = tid >> 5;
    const bool part = tid < TY * 16;
    const int tx = tid & 15, ty = tid >> 4;
    unsigned long long pd0 = 0ull, pd1 = 0ull; // warp 0: rows already used as pivot rows
    bool ok = true;
    for (int j = 0; j < k; ++j) {
        if (warp == 0) {
            double kb = 0.0;
            int ib = 0x7fffffff;
#pragma unroll
            for (int q = 0; q < (CB + 1) / 2; ++q) {
                const int i = lane + 32 * q;
                if (i < k) {
                    const double v0 = K[i * S + j];
                    colj[i] = v0;
                    const bool done = q < 2 ? ((pd0 >> (i & 63)) & 1ull) : ((pd1 >> (i & 63)) & 1ull);
                    double v = fabs(v0);
                    if (!(v == v)) v = 1.7976931348623157e308;
                    if (!done && v > kb) {
                        kb = v;
                        ib = i;
                    }
                }
            }
            const unsigned hi = (unsigned)__double2hiint(kb), lo = (unsigned)__double2loint(kb);
            const unsigned mh = __reduce_max_sync(kFull, hi);
            const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
            int ip = __reduce_min_sync(kFull, (hi == mh && lo == ml) ? ib : 0x7fffffff);
            if (prof && tid == 0) {
                const long long t = clock64();
                prof[11] += t - tq;
                tq = t;
            }
            double r = 0.0;
            if ((mh | ml) == 0u || mh >= 0x7e000000u || ip >= k) { // zero, huge or non-finite pivot
                ok = false;
                ip = -1;
            } else {
                __syncwarp();
                r = fast_rcp(colj[ip]);
                if (ip < 64) pd0 |= 1ull << ip; else pd1 |= 1ull << (ip - 64);
            }
            if (lane == 0) {
                pub[0] = r;
                reinterpret_cast<int *>(pub + 1)[0] = ip;
                if (ip >= 0) {
                    pr[j] = ip;
                    prinv[ip] = j;
                    rinv[j] = r;
                }
            }
        }
        if (TY == 2) __syncwarp(); else __syncthreads();
        if (prof && tid == 0) {
            const long long t = clock64();
            prof[12] += t - tq;
            tq = t;
        }
        if (part) {
            const double r = pub[0];
            const int ip = reinterpret_cast<const int *>(pub + 1)[0];
            if (ip >= 0) {
                const int nb = (k + 15) >> 4; // column blocks in use (uniform)
                double u[CB];
#pragma unroll
                for (int b = 0; b < CB; ++b) u[b] = b < nb ? K[ip * S + tx + 16 * b] : 0.0;
                const bool fix = tx == (j & 15);
                for (int i = ty; i < k; i += TY) {
                    if (i == ip) {
                        if (fix) K[i * S + j] = 1.0; // the identity column of the pivot row takes column j's place
                        continue;
                    }
                    const double f = colj[i];
                    if (f == 0.0) continue;
                    const double g = f * r;
                    double *row = K + i * S + tx;
#pragma unroll
                    for (int b = 0; b < CB; ++b)
                        if (b < nb) row[16 * b] = fma(-g, u[b], row[16 * b]);
                    if (fix) K[i * S + j] = -g;
                }
            }
        }
        if (prof && tid == 0) {
            const long long t = clock64();
            prof[14] += t - tq;
            tq = t;
        }
        if (TY == 2) __syncwarp(); else __syncthreads();
        if (prof && tid == 0) {
            const long long t = clock64();
            prof[15] += t - tq;
            tq = t;
        }
    }
    if (tid == 0) *okflag = ok ? 1 : 0;
}

// K = A[R, structural basis columns], zero-filled and scattered from the CSC template: one warp per
// column, lanes on its entries.
__device__ __forceinline__ void fast_gather(Fast &c, const TemplateDev &T, const double *__restrict__ theta, double *K) {
    const int k = c.k, S = c.S;
    const int md = T.dense_md, st = T.dense_stride;
    for (int e = c.tid; e < k * S; e += c.NT) K[e] = 0.0;
    if (md > 0)
        for (int kc = c.tid; kc < k; kc += c.NT) c.kref[kc] = T.dense_ref[c.bas[c.clistK[kc]]];
    __syncthreads();
    if (md > 0) {
        // dense-rows template: rows 0 .. md-1 of every structural column sit at computed addresses of
        // theta (A[r][j], row-major): one warp per row of R, lanes on the columns of K -- no loads of
        // row_idx / val_ref, neighbouring lanes read neighbouring words
        for (int r = c.warp; r < md; r += c.NW) {
            const int kr = c.rmapK[r];
            if (kr < 0) continue; // warp-uniform
            for (int kc = c.lane; kc < k; kc += 32) {
                const int ref = c.kref[kc];
                const double a = __ldg(theta + (ref >> 1) + r * st);
                K[kr * S + kc] = (ref & 1) ? -a : a;
            }
        }
    }
    for (int kc = c.warp; kc < k; kc += c.NW) { // (the rest of) every column through the template
        const int col = c.bas[c.clistK[kc]];
        const int e1 = T.col_ptr[col + 1];
        for (int e = T.col_ptr[col] + md + c.lane; e < e1; e += 32) {
            const int kr = c.rmapK[T.row_idx[e]];
            if (kr >= 0) K[kr * S + kc] = load_ref(theta, T.val_ref[e]);
        }
    }
    __syncthreads();
}

// K (the non-identity block of the current basis) gathered into shared memory and inverted
// in place.  Returns false when K is singular to working precision.
template <int NTH>
__device__ __forceinline__ bool fast_factor(Fast &c, const TemplateDev &T, const double *__restrict__ theta) {
    const int k = c.k, tid = c.tid;
    // size class of the tiled elimination (0: generic loop) and the row stride that goes with it
    const int cls = k <= 32 ? 1 : ((k <= 64 && c.NT >= 128) ? 2 : ((k <= 128 && c.NT >= 512) ? 3 : 0));
    const int S = cls == 1 ? 33 : (cls == 2 ? 65 : (cls == 3 ? 129 : (k | 1)));
    c.S = S;
    c.K = (k * S <= c.capK) ? c.Ks : c.Kg;
    const int use_cls = c.K == c.Ks ? cls : 0; // the tiled classes work in shared memory only
    if (c.K == c.Ks)
        fast_gather(c, T, theta, c.Ks);
    else
        fast_gather(c, T, theta, c.Kg);
    ftick(c, FP_BUILD);
    if (c.prof && tid == 0) {
        c.prof[FP_SOLVES] += 1;
        c.prof[FP_K_SUM] += k;
        if (k > c.prof[FP_K_MAX]) c.prof[FP_K_MAX] = k;
    }
    if (k == 0) return true;
    bool ok;
    if (use_cls == 0) {
        ok = gj_generic(c);
    } else {
        // (c.Ks, not K: the compiler then knows the address space and emits LDS/STS instead of
        // generic loads and stores)
        // blocked (tensor-core) elimination, when asked for and its scratch fits the solve vectors that are idle
        // during the factorisation: G (k x 4) in acc [NW x M], U (4 x padded columns) in dxv..tK [4 x M]
        const bool blocked = c.blocked && 4 * (((k + 15) >> 4) << 4) <= 4 * c.M;
        if (blocked) {
            if (use_cls == 1) {
                gj_blocked<2>(c.Ks, k, tid, c.NT, c.acc, c.dxv, c.uK, c.red_key, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
            } else if (use_cls == 2) {
                gj_blocked<4>(c.Ks, k, tid, c.NT, c.acc, c.dxv, c.uK, c.red_key, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
            } else if (NTH >= 512) {
                gj_blocked<8>(c.Ks, k, tid, c.NT, c.acc, c.dxv, c.uK, c.red_key, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
            }
        } else if (use_cls == 1) {
            gj_tiled<NTH / 16, 2>(c.Ks, k, tid, c.bK, c.uK, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
        } else if (use_cls == 2) {
            gj_tiled<NTH / 16, 4>(c.Ks, k, tid, c.bK, c.uK, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
        } else if (NTH >= 512) {
            gj_tiled<NTH / 16, 8>(c.Ks, k, tid, c.bK, c.uK, c.pr, c.prinv, c.rinv, &c.ctl[1], c.prof);
        }
        if (tid == 0) c.n_lu += 2ull * (unsigned long long)k * (unsigned long long)k * (unsigned long long)k;
        __syncthreads();
        ok = c.ctl[1] != 0;
    }
    __syncthreads(); // c.ctl[1] is rewritten by the next factorisation only after everyone has read it
    ftick(c, FP_GJ);
    return ok;
}

// d_C = K^-1 a_R: thread a forms row pr[a] of the stored array times the permuted right-hand side.
__device__ __forceinline__ void ftran_matvec(Fast &c, const double *K) {
    const int k = c.k, S = c.S;
    for (int a = c.tid; a < k; a += c.NT) {
        const double *row = K + c.pr[a] * S;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = 0;
        for (; j + 4 <= k; j += 4) {
            s0 = fma(row[j], c.tK[j], s0);
            s1 = fma(row[j + 1], c.tK[j + 1], s1);
            s2 = fma(row[j + 2], c.tK[j + 2], s2);
            s3 = fma(row[j + 3], c.tK[j + 3], s3);
        }
        for (; j < k; ++j) s0 = fma(row[j], c.tK[j], s0);
        c.dxv[c.clistK[a]] = ((s0 + s1) + (s2 + s3)) * c.rinv[a];
    }
}

// v_R = K^-T rhs: thread j forms storage column j (it belongs to row pr[j]) times the scaled rhs.
__device__ __forceinline__ void btran_matvec(Fast &c, const double *K) {
    const int k = c.k, S = c.S;
    for (int j = c.tid; j < k; j += c.NT) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int i = 0;
        for (; i + 4 <= k; i += 4) {
            s0 = fma(K[i * S + j], c.tK[i], s0);
            s1 = fma(K[(i + 1) * S + j], c.tK[i + 1], s1);
            s2 = fma(K[(i + 2) * S + j], c.tK[i + 2], s2);
            s3 = fma(K[(i + 3) * S + j], c.tK[i + 3], s3);
        }
        for (; i < k; ++i) s0 = fma(K[i * S + j], c.tK[i], s0);
        c.vv[c.rlistK[c.pr[j]]] = (s0 + s1) + (s2 + s3);
    }
}

// FTRAN: dx = B^-1 a_col (simplex.rs:226-229), dx by basis position.
__device__ __forceinline__ void fast_ftran(Fast &c, const TemplateDev &T, const double *__restrict__ theta, int col) {
    const int M = c.M, k = c.k, tid = c.tid;
    for (int i = tid; i < M; i += c.NT) {
        c.dxv[i] = 0.0;
        c.bK[i] = 0.0;
    }
    for (int i = tid; i < M * c.NW; i += c.NT) c.acc[i] = 0.0;
    __syncthreads();
    for (int e = T.col_ptr[col] + tid; e < T.col_ptr[col + 1]; e += c.NT) {
        const double val = load_ref(theta, T.val_ref[e]);
        const int r = T.row_idx[e];
        const int kr = c.rmapK[r];
        if (kr >= 0)
            c.bK[kr] = val;
        else
            c.dxv[c.spos[r]] = val; // a_S, corrected by F d_C below
    }
    __syncthreads();
    for (int j = tid; j < k; j += c.NT) c.tK[j] = c.bK[c.pr[j]];
    __syncthreads();
    if (c.K == c.Ks) // (shared-memory pointer: LDS instead of generic loads)
        ftran_matvec(c, c.Ks);
    else
        ftran_matvec(c, c.Kg);
    __syncthreads();
    // d_S = a_S - F d_C.  Warp w walks the basic structural columns kc = w, w + NW, ... (lanes on
    // the column's entries) and adds a_rc * d_c for the rows of S into ITS OWN accumulator row,
    // so no two warps ever touch the same word; the NW partial sums of a row are then added in
    // warp order.  Deterministic, no floating-point atomics, no row-major copy of A needed.
    const int md = T.dense_md, st = T.dense_stride;
    if (md > 0) {
        // dense-rows template: the rows of S among rows 0 .. md-1, one warp per row, lanes on the
        // columns of K (computed addresses, fixed summation tree)
        for (int r = c.warp; r < md; r += c.NW) {
            const int ps = c.spos[r];
            if (ps < 0) continue; // warp-uniform
            double s = 0.0;
            for (int kc = c.lane; kc < k; kc += 32) {
                const int ref = c.kref[kc];
                const double a = __ldg(theta + (ref >> 1) + r * st);
                s = fma((ref & 1) ? -a : a, c.dxv[c.clistK[kc]], s);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
            if (c.lane == 0) c.dxv[ps] -= s;
        }
    }
    {
        double *acc = c.acc + c.warp * M;
        for (int kc = c.warp; kc < k; kc += c.NW) {
            const int pc = c.clistK[kc];
            const double d = c.dxv[pc];
            if (d == 0.0) continue; // warp-uniform
            const int col = c.bas[pc];
            const int e1 = T.col_ptr[col + 1];
            for (int e = T.col_ptr[col] + md + c.lane; e < e1; e += 32) {
                const int r = T.row_idx[e];
                if (c.rmapK[r] < 0) acc[r] = fma(load_ref(theta, T.val_ref[e]), d, acc[r]);
            }
            __syncwarp(); // rows repeat from one column to the next, on other lanes
        }
    }
    __syncthreads();
    for (int r = tid; r < M; r += c.NT) {
        const int ps = c.spos[r];
        if (ps < 0) continue;
        double s = 0.0;
        for (int w = 0; w < c.NW; ++w) s += c.acc[w * M + r];
        c.dxv[ps] -= s;
    }
    if (c.lane == 0) c.n_solve += 2ull * k * k / c.NW;
    __syncthreads();
    ftick(c, FP_FTRAN);
}

// BTRAN: v = B^-T e_p (simplex.rs:231-234), v by row.
__device__ __forceinline__ void fast_btran(Fast &c, const TemplateDev &T, const double *__restrict__ theta, int p) {
    const int M = c.M, k = c.k, tid = c.tid;
    for (int i = tid; i < M; i += c.NT) {
        c.vv[i] = 0.0;
        c.bK[i] = 0.0;
    }
    __syncthreads();
    const int r0 = c.srow[p];
    if (r0 >= 0) { // the leaving variable is the slack of row r0: v_S = e_r0, rhs = -F[r0, :]
        for (int e = T.csr_ptr[r0] + tid; e < T.csr_ptr[r0 + 1]; e += c.NT) {
            const int pp = c.where[T.csr_col[e]];
            if (pp >= 0 && c.srow[pp] < 0) c.bK[c.cmapK[pp]] = -load_ref(theta, T.csr_ref[e]);
        }
        if (tid == 0) c.vv[r0] = 1.0;
    } else if (tid == 0) {
        c.bK[c.cmapK[p]] = 1.0;
    }
    __syncthreads();
    for (int i = tid; i < k; i += c.NT) {
        const int a = c.prinv[i];
        c.tK[i] = c.bK[a] * c.rinv[a];
    }
    __syncthreads();
    if (c.K == c.Ks)
        btran_matvec(c, c.Ks);
    else
        btran_matvec(c, c.Kg);
    if (c.lane == 0) c.n_solve += 2ull * k * k / c.NW;
    __syncthreads();
    ftick(c, FP_BTRAN);
}

template <int NTH>
__global__ void __launch_bounds__(NTH, NTH == 128 ? 4 : 1)
dz_fast_kernel(const TemplateDev T, const BatchDev Bt, const int capK) {
#ifdef DZ_EMU
    unsigned char *smem_raw = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    Fast c;
    c.M = T.M;
    c.Nn = T.Nn;
    c.NT = (int)blockDim.x;
    c.NW = c.NT >> 5;
    c.tid = (int)threadIdx.x;
    c.lane = c.tid & 31;
    c.warp = c.tid >> 5;
    c.parity = 0;
    c.blocked = Bt.resume == 1; // BatchDev::resume is unused by this kernel otherwise: 1 = blocked tensor-core elimination
    const int M = c.M, Nn = c.Nn, tid = c.tid;
    {
        double *dp = reinterpret_cast<double *>(smem_raw);
        c.Ks = dp, dp += capK;
        c.capK = capK;
        c.x = dp, dp += M;
        c.xb = dp, dp += M;
        c.dxv = dp, dp += M;
        c.vv = dp, dp += M;
        c.bK = dp, dp += M;
        c.tK = dp, dp += M;
        c.rinv = dp, dp += M;
        c.uK = dp, dp += M;
        c.acc = dp, dp += (size_t)M * c.NW;
        c.z = dp, dp += Nn;
        c.zb = dp, dp += Nn;
        c.dzv = dp, dp += Nn;
        c.prof = Bt.prof ? reinterpret_cast<long long *>(dp) : nullptr;
        dp += 16;
        c.red_key = dp, dp += 2 * 4 * kMaxWarps;
        int *ip = reinterpret_cast<int *>(dp);
        c.red_idx = ip, ip += 2 * 4 * kMaxWarps;
        c.scan = ip, ip += 2 * kMaxWarps;
        c.ctl = ip, ip += FC_WORDS;
        c.bas = ip, ip += M;
        c.srow = ip, ip += M;
        c.spos = ip, ip += M;
        c.rmapK = ip, ip += M;
        c.rlistK = ip, ip += M;
        c.cmapK = ip, ip += M;
        c.clistK = ip, ip += M;
        c.pr = ip, ip += M;
        c.prinv = ip, ip += M;
        c.pdone = ip, ip += M;
        c.kref = ip, ip += M;
        c.nb = ip, ip += Nn;
        c.where = ip, ip += T.Nint;
        c.Kg = Bt.gws + (size_t)blockIdx.x * Bt.gws_stride;
    }
    const long long max_pivots = Bt.max_pivots;

    for (;;) {
        __syncthreads();
        if (tid == 0) c.ctl[FC_LP] = (int)atomicAdd(Bt.next_lp, 1u);
        __syncthreads();
        const long long lp = (unsigned)c.ctl[FC_LP];
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
            c.where[col] = p;
            c.x[p] = load_ref(theta, T.b_ref[p]);
            c.xb[p] = 1.0;
            c.spos[p] = -1;
        }
        for (int k = tid; k < Nn; k += c.NT) {
            const int col = T.nonbasis0[k];
            c.nb[k] = col;
            c.where[col] = -1 - k;
            c.z[k] = -load_ref(theta, T.c_ref[col]);
            c.zb[k] = 1.0;
        }
        __syncthreads();
        for (int p = tid; p < M; p += c.NT) {
            const int sr = T.slack_row[c.bas[p]];
            c.srow[p] = sr;
            if (sr >= 0) c.spos[sr] = p;
        }
        __syncthreads();

        int status = DZ_OPTIMAL;
        long long pivots = 0, n_primal = 0;
        unsigned long long hash = 0xcbf29ce484222325ULL;

        while (true) {
            // ---- status(), simplex.rs:274-306 ----
            int q0, p0;
            fast_find_first_both(c, q0, p0);
            ftick(c, FP_STATUS);
            bool primal_step;
            double mu;
            if (q0 >= 0 && p0 >= 0) {
                const double primal = -c.x[p0] / c.xb[p0];
                const double dual = -c.z[q0] / c.zb[q0];
                // (the reference stops at 1e-12, simplex.rs:283; rounding residue at m_int in the hundreds is
                // larger than that and sends it into one more, ill-posed pivot: the false "infeasible")
                if (primal <= kFastOptTol && dual <= kFastOptTol) break;
                if (primal < dual) {
                    primal_step = true;
                    mu = dual;
                } else {
                    primal_step = false;
                    mu = primal;
                }
            } else if (q0 >= 0) {
                primal_step = true;
                mu = -c.z[q0] / c.zb[q0];
            } else if (p0 >= 0) {
                primal_step = false;
                mu = -c.x[p0] / c.xb[p0];
            } else {
                status = DZ_BREAKDOWN; // "unexpected code path", simplex.rs:304
                break;
            }
            if (pivots >= max_pivots) {
                status = DZ_PIVOT_CAP;
                break;
            }
            fast_lists(c);
            ftick(c, FP_LISTS);
            if (!fast_factor<NTH>(c, T, theta)) {
                status = DZ_BREAKDOWN; // singular basis
                break;
            }
            // primal_step (simplex.rs:308-318): dx = B^-1 a_j, ratio test on x, then dz;
            // dual_step   (simplex.rs:320-330): dz first, ratio test on z, then dx.
            int p = p0, q = q0;
            bool failed = false;
            for (int pass = 0; pass < 2; ++pass) {
                const bool transposed = (pass == 0) != primal_step;
                if (!transposed) {
                    fast_ftran(c, T, theta, c.nb[q]);
                } else {
                    fast_btran(c, T, theta, p);
                    // pricing: dz = -N^T v (simplex.rs:235, linalg.rs:199-207); four lanes per nonbasic
                    // column, their partial sums added by a fixed shuffle tree
                    for (int base = 0; base < Nn; base += c.NT >> 2) {
                        const int k = base + (tid >> 2), sub = tid & 3;
                        double s = 0.0;
                        bool own = false;
                        if (k < Nn) {
                            const int col = c.nb[k];
                            const int sr = T.slack_row[col];
                            if (sr >= 0) {
                                s = sub == 0 ? -c.vv[sr] : 0.0;
                                own = true;
                            } else {
                                const int tw = T.twin[col];
                                // the exact negative of an earlier nonbasic column: filled from it below
                                if (!(tw >= 0 && tw < col && c.where[tw] < 0)) {
                                    const int e0 = T.col_ptr[col], e1 = T.col_ptr[col + 1];
                                    double s1 = 0.0;
                                    int e = e0 + sub;
                                    if (T.dense_md > 0) {
                                        // dense-rows template: rows 0 .. md-1 at computed addresses of theta
                                        const int md = T.dense_md, st = T.dense_stride, ref = T.dense_ref[col];
                                        const double *a = theta + (ref >> 1);
                                        double s2 = 0.0, s3 = 0.0;
                                        int r = sub;
                                        for (; r + 12 < md; r += 16) {
                                            s = fma(__ldg(a + r * st), c.vv[r], s);
                                            s1 = fma(__ldg(a + (r + 4) * st), c.vv[r + 4], s1);
                                            s2 = fma(__ldg(a + (r + 8) * st), c.vv[r + 8], s2);
                                            s3 = fma(__ldg(a + (r + 12) * st), c.vv[r + 12], s3);
                                        }
                                        for (; r < md; r += 4) s = fma(__ldg(a + r * st), c.vv[r], s);
                                        s = (s + s1) + (s2 + s3);
                                        s = (ref & 1) ? s : -s; // dz = -a . v, and a odd reference is -theta
                                        s1 = 0.0;
                                        e += md;
                                    }
                                    for (; e + 4 < e1; e += 8) {
                                        const double a0 = load_ref(theta, T.val_ref[e]);
                                        const double a1 = load_ref(theta, T.val_ref[e + 4]);
                                        s = fma(a0, -c.vv[T.row_idx[e]], s);
                                        s1 = fma(a1, -c.vv[T.row_idx[e + 4]], s1);
                                    }
                                    if (e < e1) s = fma(load_ref(theta, T.val_ref[e]), -c.vv[T.row_idx[e]], s);
                                    s += s1;
                                    if (sub == 0) c.n_price += 2u * (unsigned)(e1 - e0);
                                    own = true;
                                }
                            }
                        }
                        s += __shfl_xor_sync(kFull, s, 1);
                        s += __shfl_xor_sync(kFull, s, 2);
                        if (own && sub == 0) c.dzv[k] = s;
                    }
                    __syncthreads();
                    // a nonbasic column whose exact negative (the other half of a split variable)
                    // is nonbasic too and comes first: dz is the twin's, negated
                    for (int k = tid; k < Nn; k += c.NT) {
                        const int col = c.nb[k];
                        const int tw = T.slack_row[col] < 0 ? T.twin[col] : -1;
                        if (tw >= 0 && tw < col) {
                            const int wt = c.where[tw];
                            if (wt < 0) c.dzv[k] = -c.dzv[-1 - wt];
                        }
                    }
                    __syncthreads();
                    ftick(c, FP_PRICE);
                }
                if (pass == 0) {
                    if (primal_step) {
                        p = fast_find_second(c, mu, c.x, c.xb, c.dxv, M);
                        if (p < 0) {
                            status = DZ_UNBOUNDED;
                            failed = true;
                        }
                    } else {
                        q = fast_find_second(c, mu, c.z, c.zb, c.dzv, Nn);
                        if (q < 0) {
                            status = DZ_INFEASIBLE;
                            failed = true;
                        }
                    }
                    ftick(c, FP_RATIO);
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
                t = (xp == 0.0 && dxp == 0.0) ? 0.0 : xp / dxp;
                s = (zq == 0.0 && dzq == 0.0) ? 0.0 : zq / dzq;
                t_bar = (xbp == 0.0 && dxp == 0.0) ? 0.0 : xbp / dxp;
                s_bar = (zbq == 0.0 && dzq == 0.0) ? 0.0 : zbq / dzq;
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
                    c.x[k] = fma(-t, d, c.x[k]);
                    c.xb[k] = fma(-t_bar, d, c.xb[k]);
                }
            }
            for (int k = tid; k < Nn; k += c.NT) {
                const double d = c.dzv[k];
                if (k == q) {
                    c.z[k] = s;
                    c.zb[k] = s_bar;
                } else {
                    c.z[k] = fma(-s, d, c.z[k]);
                    c.zb[k] = fma(-s_bar, d, c.zb[k]);
                }
            }
            n_upd += 4ull * (M + Nn);
            // swap (simplex.rs:239-251) and the structure that follows the basis
            if (tid == 0) {
                const int srl = T.slack_row[leaving], sre = T.slack_row[entering];
                c.bas[p] = entering;
                c.nb[q] = leaving;
                c.where[entering] = p;
                c.where[leaving] = -1 - q;
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
            {
                const unsigned long long w = (unsigned long long)(primal_step ? 0u : 1u) |
                                             ((unsigned long long)(unsigned)leaving << 1) |
                                             ((unsigned long long)(unsigned)entering << 32);
                hash = (hash ^ w) * 0x100000001b3ULL;
            }
            ++pivots;
            if (primal_step) ++n_primal;
            __syncthreads();
            ftick(c, FP_UPDATE);
        }

        __syncthreads();
        if (c.prof) {
            __syncthreads();
            if (tid < 16) Bt.prof[(size_t)lp * 16 + tid] = c.prof[tid];
        }
        if (Bt.work) { // executed flop counts
            double *w = Bt.work + (size_t)lp * 8;
            if (c.n_lu) atomicAdd(&w[0], (double)c.n_lu);
            if (c.n_solve) atomicAdd(&w[1], (double)c.n_solve);
            if (c.n_price) atomicAdd(&w[2], (double)c.n_price);
            if (tid == 0) atomicAdd(&w[3], (double)n_upd);
        }
        // ---- results: objective_value / solution, simplex.rs:345-371 ----
        if (tid == 0) {
            double obj = 0.0;
            for (int p = 0; p < M; ++p) obj = fma(load_ref(theta, T.c_ref[c.bas[p]]), c.x[p], obj);
            obj = load_ref(theta, T.c0_ref) + obj;
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
                const int wp = c.where[T.pos_index[v]], wn = c.where[T.neg_index[v]];
                const double pos = wp >= 0 ? c.x[wp] : 0.0, neg = wn >= 0 ? c.x[wn] : 0.0;
                Bt.values[(size_t)lp * T.n_orig + v] = pos - neg;
            }
        }
    }
}

size_t fast_fixed_smem_bytes(int M, int Nn, int Nint, int nwarps) {
    const size_t doubles = (8 + (size_t)nwarps) * (size_t)M + 3 * (size_t)Nn + 16 + 2 * 4 * kMaxWarps;
    const size_t ints = 2 * 4 * kMaxWarps + 2 * kMaxWarps + FC_WORDS + 11 * (size_t)M + (size_t)Nn + (size_t)Nint;
    return doubles * 8 + ints * 4 + 16;
}

template <int NTH>
cudaError_t launch_fast_one(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, cudaStream_t st) {
    auto kern = dz_fast_kernel<NTH>;
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

} // namespace

// Launch plan of the fast kernel: CTA per LP, K in shared memory up to a capacity that
// leaves `cps` CTAs per SM (larger K's go to the per-CTA HBM/L2 workspace).
int plan_fast(int device, int32_t M, int32_t Nn, int32_t Nint, int64_t B, int32_t cps_hint, LaunchPlan *plan,
              std::string *err) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        *err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    if (M > 512) {
        *err = "fast numerics: m_int > 512 is not supported (use the exact path)";
        return DZ_ERR_LIMIT;
    }
    const size_t max_smem = prop.sharedMemPerBlockOptin;
    const size_t per_sm = prop.sharedMemPerMultiprocessor;
    const int sms = prop.multiProcessorCount;
    const int block = M <= 128 ? 128 : 512;
    const size_t fixed = fast_fixed_smem_bytes(M, Nn, Nint, block / 32);
    // the tiled size classes store K with the strides 33 / 65 / 129 (fast_factor)
    const size_t stride_max = std::max<size_t>((size_t)(M | 1), block == 512 ? 129 : 65);
    const size_t full = (size_t)M * stride_max * 8;
    if (fixed + 1024 > max_smem) {
        *err = "fast numerics: the per-LP vectors do not fit shared memory";
        return DZ_ERR_LIMIT;
    }
    int cps = cps_hint > 0 ? cps_hint : (M <= 128 ? 4 : 1);
    if (block == 512) cps = 1; // the 128-register tile kernel fills the register file
    cps = std::max(1, std::min(cps, 2048 / block));
    size_t per_cta = std::min<size_t>(per_sm / cps - 1024, max_smem);
    if (per_cta < fixed + 64) per_cta = fixed + 64;
    size_t cap_bytes = std::min(per_cta - fixed, (full + 15) & ~(size_t)15) & ~(size_t)15;
    *plan = LaunchPlan();
    plan->fast_mode = true;
    plan->core_cap_w = (int32_t)(cap_bytes / 8);
    plan->home = 6;
    plan->block = block;
    plan->worker_warps = block / 32;
    plan->w_in_smem = cap_bytes >= full;
    plan->smem_bytes = (int32_t)(fixed + cap_bytes);
    plan->ctas_per_sm = cps;
    plan->gws_doubles_per_cta = (int64_t)((full + 15) / 8);
    int64_t grid = (int64_t)sms * cps;
    if (grid > B) grid = B;
    plan->grid = (int32_t)std::max<int64_t>(grid, 1);
    plan->teams = plan->grid;
    return DZ_OK;
}

int launch_fast(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream, std::string *err) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = plan.block == 128 ? launch_fast_one<128>(T, Bt, plan, st) : launch_fast_one<512>(T, Bt, plan, st);
    if (e != cudaSuccess) {
        *err = std::string("dz_fast_kernel launch: ") + cudaGetErrorString(e);
        return DZ_ERR_CUDA;
    }
    return DZ_OK;
}

} // namespace dz
