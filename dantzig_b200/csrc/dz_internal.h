// dz_internal.h -- shared host/device declarations of the dantzig_b200 library.
#ifndef DZ_INTERNAL_H
#define DZ_INTERNAL_H

#include "../../include/dantzig_b200.h"

#include <cstdint>
#include <string>
#include <vector>

namespace dz {

// Host-side lowered template (see dz_lower.cpp).
struct Template {
    int32_t m = 0, n_int = 0;
    std::vector<int64_t> col_ptr;
    std::vector<int32_t> row_idx, val_ref;
    std::vector<int32_t> c_ref, b_ref, basis0, nonbasis0;
    std::vector<int32_t> orig_var, pos_index, neg_index;
    // slack_row[col]: the row of a slack column (one entry, the constant +1.0), else -1.
    // twin[col]: the column that is this column's exact negative (pos/neg of one original
    // variable: same rows, same references with the opposite sign), else -1.
    std::vector<int32_t> slack_row, twin;
    uint64_t structure_hash = 0; // of the structural arrays the template was built from
    int32_t c0_ref = -1;
    // theta layout
    int32_t n_vars = 0, n_obj = 0, n_rows_user = 0;
    int64_t n_row_terms = 0;
    int64_t off_obj = 0, off_rowcoef = 0, off_rhs = 0, off_lb = 0, off_ub = 0, n_theta = 0;
};

int build_template(const dz_model *m, Template *t, std::string *err);
int pack_theta(const Template *t, const dz_model *m, double *theta, std::string *err);

// Device view of a template: plain pointers into one device allocation.
struct TemplateDev {
    int32_t M, Nint, Nn, n_orig;
    int32_t nnz;
    int32_t c0_ref;
    int32_t S;                // row stride of the working basis: (M + 1) | 1
    int32_t pad_;
    const int32_t *col_ptr;   // [Nint+1]
    const int32_t *row_idx;   // [nnz]
    const int32_t *val_ref;   // [nnz]
    const int32_t *c_ref;     // [Nint]
    const int32_t *b_ref;     // [M]
    const int32_t *basis0;    // [M]
    const int32_t *nonbasis0; // [Nn]
    const int32_t *pos_index; // [n_orig]
    const int32_t *neg_index; // [n_orig]
    const int32_t *slack_row; // [Nint]
    const int32_t *twin;      // [Nint]
    // row-major (CSR) view of the same pattern, used by the fast-numerics kernel (dz_fast.cu)
    const int32_t *csr_ptr;   // [M+1]
    const int32_t *csr_col;   // [nnz]
    const int32_t *csr_ref;   // [nnz]
    // DENSE-ROWS structure (fast-numerics kernel): when every structural column starts with entries
    // in rows 0 .. dense_md-1 whose theta references advance by the same stride from row to row
    // (a dense user matrix A[r][j] stored row-major in theta, all rows of one sense), the kernel
    // computes those addresses instead of loading row_idx / val_ref.  dense_ref[col] is the signed
    // reference of the column's row-0 entry (-1 for slacks); dense_md == 0 switches it off.
    int32_t dense_md, dense_stride;
    const int32_t *dense_ref; // [Nint]
};

// Device view of one batch: inputs and outputs, all in HBM.
struct BatchDev {
    int64_t B;
    const double *theta; // [B][n_theta]
    int64_t n_theta;
    int64_t max_pivots;
    int32_t trace_cap;
    int32_t *status;
    int32_t *pivots;
    int32_t *n_primal;
    unsigned long long *trace_hash;
    double *objective;
    double *values;  // [B][n_orig]
    double *x_basic; // [B][M]
    int32_t *basis;  // [B][M]
    int32_t *trace;  // [B][trace_cap][3] or null
    double *work;    // [B][8]
    long long *prof; // [B][16] per-phase cycles (optional, null = off)
    unsigned int *next_lp; // work-queue counter
    double *gws;     // global workspace for the basis when it does not fit in smem
    int64_t gws_stride; // doubles per CTA
    // Hand-over from the on-chip core kernel (dz_core.cu) to the general kernel: LP ids the
    // core kernel gave up, each with the state of its pivot loop at that point.
    int32_t *exo_list;        // [B]
    unsigned int *exo_count;  // number of entries
    unsigned char *exo_state; // [B][exo_stride]: x, xbar [M], z, zbar [Nn] f64; pivots, n_primal, hash i64;
                              // basis [M], nonbasis [Nn] i32
    int64_t exo_stride;
    int32_t resume;           // general kernel: 1 = work items are exo_list[0..*exo_count), continued from exo_state
    unsigned int *next_lp2;   // work-queue counter of that second launch
};

// Workspace of the whole-GPU single-LP kernel (dz_grid.cu): pointers into one HBM allocation.
struct GridDev {
    double *x, *xb, *dxv, *vv, *ycore, *z, *zb, *dzv;
    double *lval;  // [nnz] the LP's lowered values, resolved once from theta
    double *pkey;  // [4 * CTAs] partial arg-max keys
    double *jobd;  // job descriptor, doubles
    double *W;     // working core, dense rows of (nr + 1) | 1 doubles
    int *bas, *nb, *rowAt, *posOf, *rowcnt, *srow, *spos, *rmapR, *rlist, *pmap, *plist, *pivr, *pend;
    int *list;     // rows with a nonzero in the current pivot column
    int *pcols, *pwq; // compacted pivot row: pattern columns; mask word indices
    unsigned *pwb;    // ... and the mask words
    double *pvals;    // ... and the values
    int *rlast;                 // last column of each core row's nonzero pattern
    int *done;                  // back-substitution: component of this core column is final
    int *where;    // column -> nonbasic slot k >= 0, or -1 - (basis position)
    int *pidx, *bcnt, *job;
    int *seu, *set_; // positions (columns of B / of B^T) whose elimination step is not a no-op from the start
    unsigned *bar;   // grid barrier: arrivals, generation
    unsigned *rmask; // [nr][ceil(nr/32)] columns of each core row that may be nonzero
    long long w_cap; // doubles available at W
};

// Launch plan computed on the host (dz_kernel.cu).
struct LaunchPlan {
    int32_t grid = 0, block = 0, smem_bytes = 0, ctas_per_sm = 0, worker_warps = 1, home = 0;
    bool w_in_smem = true;
    bool warp_mode = false;    // one warp per LP instead of one CTA per LP
    int32_t smem_per_team = 0; // warp mode: shared-memory slab per warp
    int64_t teams = 0;         // CTAs (or warps) that own a workspace slab
    int64_t gws_doubles_per_cta = 0;
    bool grid_mode = false;    // whole-GPU single-LP kernel (dz_grid.cu), cooperative launch
    bool core_mode = false;    // on-chip coupled-core kernel (dz_core.cu)
    int32_t core_cap_w = 0;    // its shared-memory capacity for the working core, in doubles
    bool fast_mode = false;    // opt-in fast numerics (dz_fast.cu); core_cap_w = capacity for K
};

int plan_launch(int device, int32_t M, int32_t Nn, int64_t nnz, int64_t B, int32_t warps_hint,
                int32_t cps_hint, int32_t basis_home, LaunchPlan *plan, std::string *err);
// Enqueue the batched solve on `stream` (cudaStream_t passed as void*).
int launch_batch(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream,
                 std::string *err);
// dz_core.cu
size_t core_fixed_smem_bytes(int M, int Nn, int NQ);
int core_nq(int M); // rows-per-lane class of the core kernel (0: m_int too large for it)
int launch_core(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream,
                std::string *err);
// dz_fast.cu (opt-in fast numerics, dz_options.numerics == DZ_NUMERICS_FAST)
int plan_fast(int device, int32_t M, int32_t Nn, int32_t Nint, int64_t B, int32_t cps_hint, LaunchPlan *plan,
              std::string *err);
int launch_fast(const TemplateDev &T, const BatchDev &Bt, const LaunchPlan &plan, void *stream, std::string *err);
// dz_grid.cu
size_t grid_workspace(int M, int Nn, long long nnz, int nblocks, long long w_cap_doubles, unsigned char *base,
                      GridDev *d);
size_t grid_smem_bytes();
int launch_grid(const TemplateDev &T, const BatchDev &Bt, const GridDev &D, const LaunchPlan &plan, void *stream,
                std::string *err);
int measure_fp64_peak(int device, double *mul_sub_gflops, double *fma_gflops, std::string *err);

void set_error(const std::string &s);

} // namespace dz

#endif
