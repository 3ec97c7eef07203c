/*
 * dzo.h -- C interface of the CPU ORACLE for the dantzig simplex hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product (dantzig_b200/) never links, loads or
 * calls it.
 *
 * Parity status: PINNED.  The restatement reproduces every known-answer test
 * the reference holds for this path (src/linalg.rs:323-446, src/simplex.rs:
 * 485-796, tests/test_optimize.py, tests/test_exceptions.py); see
 * tests/test_oracle_kat.py.  The Rust reference itself cannot be built here
 * (no cargo/rustc in the image, no network), so there is no oracle/_ref.
 *
 * The restatement follows, function by function:
 *   lowering            src/simplex.rs:11-81,123-224, src/model.rs:11-22
 *   pivot loop          src/simplex.rs:226-343,410-468
 *   linear algebra      src/linalg.rs:8-10,88-128,180-207,219-234,282-299
 *   result extraction   src/simplex.rs:345-371
 * Arithmetic contract: IEEE binary64, round-to-nearest, no FMA contraction
 * (compiled with -ffp-contract=off), no re-association, true division.
 */
#ifndef DZO_H
#define DZO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Solver outcome.  PANIC mirrors a Rust panic on the reference path
 * (safe_divide assert simplex.rs:466, "unexpected code path" simplex.rs:304,
 * the 0x0 basis underflow linalg.rs:95).  PIVOT_CAP does not exist in the
 * reference (it would recurse forever / overflow the stack). */
enum {
    DZO_OPTIMAL = 0,
    DZO_UNBOUNDED = 1,
    DZO_INFEASIBLE = 2,
    DZO_PANIC = 3,
    DZO_PIVOT_CAP = 4
};

/* Arithmetic variants.  Both produce bit-identical pivot traces and
 * solutions (tests/test_oracle_variants.py); LITERAL does every operation the
 * reference does (two fresh dense LUs per pivot, no skipping) and is the
 * variant timed as the CPU baseline; SKIP elides operations whose operand is an
 * exact zero while every other operand is finite. */
enum { DZO_LITERAL = 0, DZO_SKIP = 1, DZO_SPARSE = 2 };
/* DZO_SPARSE: the SKIP variant's operations in the SKIP variant's order on rows
 * stored as ordered maps (no dense m x m array), so a lowered 70 000 x 170 000
 * transportation LP can be followed.  Falls back to dense SKIP for a solve in
 * which a non-finite value appears (when m*m fits), else reports DZO_PANIC. */

/*
 * Model as handed to rust.solve(objective, constraints) (src/lib.rs:16-27):
 * MAXIMISE obj_const + sum obj_coef[t]*var(obj_var[t]) subject to, for every
 * row r, sum_{t in [row_ptr[r],row_ptr[r+1])} row_coef[t]*var(row_var[t]) <= rhs[r].
 * Variables are entries of a table (index 0..n_vars-1) carrying optional
 * bounds (src/pyobjs.rs:12-19).  Term order is significant: it fixes the
 * lowered column order and so every tie-break (simplex.rs:126-176).
 */
typedef struct {
    int32_t n_vars;
    const uint8_t *has_lb; /* [n_vars] */
    const uint8_t *has_ub; /* [n_vars] */
    const double *lb;      /* [n_vars] */
    const double *ub;      /* [n_vars] */
    int32_t n_obj;
    const int32_t *obj_var; /* [n_obj] */
    const double *obj_coef; /* [n_obj] */
    double obj_const;
    int32_t n_rows;
    const int64_t *row_ptr; /* [n_rows+1] */
    const int32_t *row_var; /* [row_ptr[n_rows]] */
    const double *row_coef; /* [row_ptr[n_rows]] */
    const double *rhs;      /* [n_rows] */
} dzo_model;

typedef struct dzo_lowered dzo_lowered; /* opaque: result of Simplex::new */

/* Restatement of Simplex::new (simplex.rs:123-224).  Never fails on a
 * well-formed model; returns NULL on malformed indices. */
dzo_lowered *dzo_lower(const dzo_model *model);
void dzo_lowered_free(dzo_lowered *lp);

/* Build a lowered problem directly from computational-form arrays (used to
 * check the device path on inputs that did not come through dzo_lower).
 * CSC of the m x n_int constraint matrix, c[n_int], c0, b[m], initial basis[m]
 * and nonbasis[n_int-m] index lists (position ordered). */
dzo_lowered *dzo_lowered_from_arrays(int32_t m, int32_t n_int, const int64_t *col_ptr,
                                     const int32_t *row_idx, const double *val, const double *c,
                                     double c0, const double *b, const int32_t *basis,
                                     const int32_t *nonbasis);

/* Dimensions: m_int rows, n_int columns, nnz of the CSC, number of distinct
 * original variables seen (first-appearance order). */
void dzo_lowered_dims(const dzo_lowered *lp, int32_t *m, int32_t *n_int, int64_t *nnz,
                      int32_t *n_orig);
/* Copy out the lowered arrays; any pointer may be NULL. */
void dzo_lowered_get(const dzo_lowered *lp, int64_t *col_ptr, int32_t *row_idx, double *val,
                     double *c, double *c0, double *b, int32_t *basis0, int32_t *nonbasis0,
                     int32_t *orig_var, int32_t *pos_index, int32_t *neg_index);

typedef struct {
    int32_t status;
    int64_t pivots;
    int64_t n_primal;
    int64_t n_dual;
    uint64_t trace_hash; /* FNV-1a style hash over (kind, leaving, entering) */
    double objective;    /* c0 + sum_p c[basis[p]]*x[p], p ascending (simplex.rs:345-352) */
} dzo_result;

/*
 * Run Simplex::solve (simplex.rs:332-343) on a lowered problem.
 *  variant     DZO_LITERAL or DZO_SKIP
 *  max_pivots  <=0: unlimited
 *  x_basic     [m]   out, may be NULL: final x (position ordered)
 *  basis       [m]   out, may be NULL: final basic variable indices
 *  values      [n_orig] out, may be NULL: pos - neg per original variable
 *              (simplex.rs:354-371), in first-appearance order
 *  trace       [3*trace_cap] out, may be NULL: (kind 0=primal 1=dual, leaving i, entering j)
 */
int dzo_solve(const dzo_lowered *lp, int variant, int64_t max_pivots, dzo_result *res,
              double *x_basic, int32_t *basis, double *values, int32_t *trace, int64_t trace_cap);

/* Dense helpers restating linalg.rs for the known-answer tests. */
/* Matrix::factorize (linalg.rs:88-128): a[n*n] row-major in place, p[n-1]. */
void dzo_lu_factorize(double *a, int32_t n, int32_t *p);
/* lu_solve (linalg.rs:8-10): a is consumed, b[n] overwritten by the solution. */
void dzo_lu_solve(double *a, int32_t n, double *b, int variant);
/* CscMatrix::neg_t_dot (linalg.rs:199-207). */
void dzo_neg_t_dot(int32_t nrows, int32_t ncols, const int64_t *col_ptr, const int32_t *row_idx,
                   const double *val, const double *v, double *out);
/* Matrix::to_sparse (linalg.rs:254-270): dense row-major -> CSC, drops exact zeros.
 * Returns nnz; arrays must hold nrows*ncols entries. */
int64_t dzo_dense_to_csc(const double *dense, int32_t nrows, int32_t ncols, int64_t *col_ptr,
                         int32_t *row_idx, double *val);

/* Floating point operation counters of the last dzo_solve on this thread
 * (multiplications+subtractions+divisions actually executed in LU + solves +
 * pricing + updates).  Used by bench.py to state executed work. */
void dzo_last_flops(double *lu_flops, double *solve_flops, double *other_flops);

#ifdef __cplusplus
}
#endif
#endif
