/*
 * dantzig_b200.h -- C ABI of the B200-native simplex hot path.
 *
 * This is the drop-in boundary for dantzig's solver path: the entry points
 * below are what the reference's FFI layer would bind in place of
 *   Simplex::new(objective, constraints)          /root/reference/src/simplex.rs:123-224
 *   Simplex::solve(self) -> Result<Simplex,Error>  /root/reference/src/simplex.rs:332-343
 *   Simplex::objective_value / Simplex::solution   /root/reference/src/simplex.rs:345-371
 * as called from #[pyfunction] solve               /root/reference/src/lib.rs:16-27
 * plus the NEW batched entry point BASELINE.json's north_star asks for.
 *
 * Plain C, caller-owned buffers, no torch/pybind types.  Every function
 * returns 0 on success and a negative DZ_ERR_* on API misuse or CUDA failure
 * (text via dz_last_error()).  The SOLVER outcome (optimal / unbounded /
 * infeasible / breakdown) is data, reported in `status` outputs.
 *
 * There is no CPU fallback: every dz_solve_* call runs the CUDA kernels and
 * fails with DZ_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef DANTZIG_B200_H
#define DANTZIG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DZ_VERSION 100

/* ---- error codes (function return values) -------------------------------- */
enum {
    DZ_OK = 0,
    DZ_ERR_ARG = -1,      /* malformed argument / index out of range          */
    DZ_ERR_CUDA = -2,     /* CUDA runtime failure, or no usable device         */
    DZ_ERR_LIMIT = -3,    /* problem exceeds an implementation limit           */
    DZ_ERR_ALLOC = -4     /* host or device allocation failed                  */
};

/* ---- solver outcome (per LP) ---------------------------------------------
 * Mirrors error.rs:4-7 plus the two outcomes the reference can only express
 * as a Rust panic or never-ending recursion. */
enum {
    DZ_OPTIMAL = 0,    /* Ok(simplex)                      simplex.rs:334       */
    DZ_UNBOUNDED = 1,  /* Err(Error::Unbounded)            simplex.rs:313       */
    DZ_INFEASIBLE = 2, /* Err(Error::Infeasible)           simplex.rs:325       */
    DZ_BREAKDOWN = 3,  /* reference panics: safe_divide assert simplex.rs:466,
                          "unexpected code path" :304, 0x0 basis linalg.rs:95  */
    DZ_PIVOT_CAP = 4   /* watchdog: max_pivots reached (no reference analogue) */
};

/* ---- model description -----------------------------------------------------
 * Exactly the information rust.solve(objective, constraints) receives
 * (lib.rs:16-27, pyobjs.rs:12-19,114-152):
 *   MAXIMISE obj_const + sum_t obj_coef[t] * var(obj_var[t])
 *   s.t. for each row r: sum_{t in [row_ptr[r],row_ptr[r+1])} row_coef[t]*var(row_var[t]) <= rhs[r]
 * Variables are rows of a table with optional bounds.  TERM ORDER MATTERS: it
 * fixes the order of the lowered columns (simplex.rs:126-176) and therefore
 * every tie-break of the pivot rules.
 */
typedef struct dz_model {
    int32_t n_vars;
    const uint8_t *has_lb;  /* [n_vars] 1 = finite lower bound present */
    const uint8_t *has_ub;  /* [n_vars]                                */
    const double *lb;       /* [n_vars] ignored where has_lb == 0      */
    const double *ub;       /* [n_vars]                                */
    int32_t n_obj;
    const int32_t *obj_var; /* [n_obj] */
    const double *obj_coef; /* [n_obj] */
    double obj_const;
    int32_t n_rows;
    const int64_t *row_ptr; /* [n_rows+1] */
    const int32_t *row_var; /* [row_ptr[n_rows]] */
    const double *row_coef; /* [row_ptr[n_rows]] */
    const double *rhs;      /* [n_rows] */
} dz_model;

/* ---- lowered template -------------------------------------------------------
 * Result of the host-side restatement of Simplex::new: the computational form
 * (CSC pattern, column order, initial basis) with every numeric entry held as
 * a signed REFERENCE into a parameter vector `theta` instead of a value, so
 * that one template serves any number of LPs sharing the model's structure.
 *
 * theta layout for a model with n_obj objective terms, T = row_ptr[n_rows] row
 * terms, n_rows rows, n_vars variables (all doubles):
 *   [0]                       1.0
 *   [1]                       obj_const
 *   [2, 2+n_obj)              obj_coef
 *   [.., +T)                  row_coef
 *   [.., +n_rows)             rhs
 *   [.., +n_vars)             lb   (0.0 where absent)
 *   [.., +n_vars)             ub   (0.0 where absent)
 * A reference is (index << 1) | negate; -1 means the constant 0.0.
 */
typedef struct dz_template dz_template;

int dz_template_create(const dz_model *structure, dz_template **out);
void dz_template_destroy(dz_template *t);

typedef struct dz_template_info {
    int32_t m;       /* lowered rows  (m_int)                       */
    int32_t n_int;   /* lowered columns                              */
    int32_t n_orig;  /* distinct original variables, first-seen order */
    int64_t nnz;     /* structural nonzeros of the lowered matrix     */
    int64_t n_theta; /* length of one parameter vector                */
} dz_template_info;
int dz_template_get_info(const dz_template *t, dz_template_info *info);

/* Copy out the lowered arrays (any pointer may be NULL):
 * col_ptr[n_int+1], row_idx[nnz], val_ref[nnz], c_ref[n_int], b_ref[m],
 * basis0[m], nonbasis0[n_int-m], orig_var[n_orig], pos_index[n_orig],
 * neg_index[n_orig]. */
int dz_template_get_arrays(const dz_template *t, int64_t *col_ptr, int32_t *row_idx,
                           int32_t *val_ref, int32_t *c_ref, int32_t *b_ref, int32_t *basis0,
                           int32_t *nonbasis0, int32_t *orig_var, int32_t *pos_index,
                           int32_t *neg_index);

/* Pack the numeric content of `model` (same structure as the template's) into
 * theta[n_theta] following the layout above. */
int dz_template_pack_theta(const dz_template *t, const dz_model *model, double *theta);

/* ---- solve options ---------------------------------------------------------- */
enum { DZ_NUMERICS_EXACT = 0, DZ_NUMERICS_FAST = 1 };
typedef struct dz_options {
    int32_t device;       /* CUDA device ordinal                                       */
    int64_t max_pivots;   /* <=0: default watchdog 100*(m+n_int)+1000                  */
    int32_t trace_cap;    /* per-LP pivot trace entries to record (0 = none)           */
    int32_t worker_warps; /* 0 = auto; >0 = CTA per LP with that many worker warps;
                             -1 = one warp per LP (no CTA barriers).  Tuning knob,
                             never changes results of the exact path.  With
                             DZ_NUMERICS_FAST: 2 = blocked elimination with the rank-4
                             update on the FP64 tensor cores (same pivots; rounding
                             differs in the last bits), anything else = step by step  */
    int32_t ctas_per_sm;  /* 0 = auto                                                  */
    void *stream;         /* cudaStream_t to launch on (NULL = the library's stream)   */
    int32_t profile;      /* 1 = record per-LP phase cycle counts (dz_batch_result.prof) */
    int32_t basis_home;   /* where the dense working basis lives: 0 = auto, 1 = shared
                             memory (lowest latency, few LPs per SM), 2 = HBM/L2
                             workspace (more LPs in flight per SM), 3 = basis AND
                             per-LP vectors in HBM (chosen automatically when they
                             exceed shared memory: m in the thousands), 4 = the on-chip
                             coupled-core kernel (the automatic choice for m_int <= 256) */
    int32_t numerics;     /* DZ_NUMERICS_EXACT (0, the default): the reference's floating-point
                             operations in the reference's order -- bit-identical pivot
                             sequences.  DZ_NUMERICS_FAST (1): OPT-IN, never a default: one
                             factorisation per pivot reused for BTRAN (the reference factorises
                             B and B^T separately, simplex.rs:226-236), fused multiply-add,
                             identity blocks of the basis eliminated symbolically.  Same pivot
                             rules; results agree to rounding, not to the bit.  m_int <= 512
                             (the batched classes and config 1); larger LPs: exact path only */
} dz_options;
void dz_options_default(dz_options *o);

/* ---- per-LP results ---------------------------------------------------------
 * All arrays are [B] or [B][k] row-major and caller-owned; NULL = not wanted. */
typedef struct dz_batch_result {
    int32_t *status;      /* [B] DZ_OPTIMAL...                                         */
    int32_t *pivots;      /* [B] pivot count                                            */
    int32_t *n_primal;    /* [B] primal-step count                                      */
    uint64_t *trace_hash; /* [B] FNV-1a over (kind, leaving, entering)                  */
    double *objective;    /* [B] c0 + sum_p c[basis[p]]*x[p], p ascending               */
    double *values;       /* [B][n_orig] pos-neg per original variable (simplex.rs:354) */
    double *x_basic;      /* [B][m]      final x, position ordered                      */
    int32_t *basis;       /* [B][m]      final basic column per position                */
    int32_t *trace;       /* [B][trace_cap][3] (kind 0=primal 1=dual, leaving, entering)*/
    double *work;         /* [B][8] executed flops: LU, solves, pricing, updates; then, from
                             the single-LP kernel: doubles of working core cleared and
                             filled over all solves, elimination steps that did arithmetic,
                             of those the ones run grid-wide; one spare                  */
    int64_t *prof;        /* [B][16] SM cycles per phase + step counters (opt.profile)  */
} dz_batch_result;

/* ---- NEW batched entry point: host buffers in, host buffers out -------------
 * Solves B LPs that share `t`'s structure; LP i's numbers are theta[i*n_theta..].
 * Copies theta to the device, runs the CTA-per-LP kernel, copies results back. */
int dz_solve_batch(const dz_template *t, int64_t B, const double *theta, const dz_options *opt,
                   dz_batch_result *out);

/* Same, sharded over the first n_gpus devices: device d takes the contiguous
 * LP range [d*B/n_gpus, (d+1)*B/n_gpus), one host thread and one stream per device, no
 * collective -- every shard's results land in its slice of `out`.  opt->device is ignored;
 * n_gpus <= 0 means every visible device.  A single LP does not shard (its pivots are
 * serially dependent): B < n_gpus simply leaves devices idle. */
int dz_solve_batch_multi(const dz_template *t, int64_t B, const double *theta, int32_t n_gpus,
                         const dz_options *opt, dz_batch_result *out);

/* ---- device-resident batch (inputs stay in HBM across solves) -------------- */
typedef struct dz_batch dz_batch;
int dz_batch_create(const dz_template *t, int64_t B, const dz_options *opt, dz_batch **out);
void dz_batch_destroy(dz_batch *b);
int dz_batch_upload(dz_batch *b, const double *theta);        /* H2D, async on the stream */
/* Array-native front door with the lowering of the NUMBERS on the device:
 * B dense user-level LPs of one shape,  min/max c.x  s.t.  A x (senses) rhs,  lb <= x <= ub,  in
 * plain arrays A[B][m][n], rhs[B][m], c[B][n] (HOST memory when on_device == 0, DEVICE memory on
 * the batch's device otherwise), senses[m] in {0: <=, 1: >=, 2: ==} and lb/ub[n] on the host
 * (shared by the batch; ignored where the template's variable has no such bound).  A kernel
 * writes the batch's parameter vectors straight into HBM the way the reference frontend lowers
 * the same model (>= rows negated, == rows as the <= row followed by the negated row,
 * model.py:351-375; Minimize negates the objective, optimize.py:115): no host-side theta.
 * The template must come from the matching dense structure (every row and the objective mention
 * every variable in index order). */
int dz_batch_pack_dense(dz_batch *b, const double *A, const double *rhs, const double *c,
                        const int32_t *senses, const double *lb, const double *ub, int32_t m,
                        int32_t n, int32_t minimize, int32_t on_device);
int dz_batch_solve(dz_batch *b);                               /* launch, async            */
int dz_batch_download(dz_batch *b, dz_batch_result *out);      /* D2H + synchronize        */
int dz_batch_sync(dz_batch *b);
/* Device time of the last dz_batch_solve in ms (cudaEvents on the launch
 * stream), number of kernel launches it made, and the launch configuration. */
int dz_batch_last_timing(dz_batch *b, float *kernel_ms, int32_t *launches);
int dz_batch_launch_info(dz_batch *b, int32_t *grid, int32_t *block, int32_t *smem_bytes,
                         int32_t *ctas_per_sm, int32_t *w_in_smem);
/* Bytes moved by upload / download for a batch of this shape. */
int dz_batch_io_bytes(dz_batch *b, int64_t *h2d, int64_t *d2h);

/* ---- single LP: the replacement for Simplex::new(..).solve() ---------------
 * Lowers `model` on the host, solves it on the device (batch of one) and
 * returns status, objective and one value per variable of `model` (0.0 for a
 * variable the model never mentions, pyobjs.rs:163-165). */
typedef struct dz_solution {
    int32_t status;
    int32_t pivots;
    int32_t n_primal;
    uint64_t trace_hash;
    double objective;
} dz_solution;
int dz_solve_model(const dz_model *model, const dz_options *opt, dz_solution *sol,
                   double *values /* [model->n_vars] */);

/* ---- misc ------------------------------------------------------------------- */
const char *dz_last_error(void);
int dz_device_count(void);
int dz_device_info(int device, char *name, int name_len, int *sm_count, int *cc_major,
                   int *cc_minor, int64_t *smem_per_sm);
int dz_version(void);

/* FP64 pipe micro-benchmark used as the roofline denominator for the batched
 * kernel: sustained un-fused multiply+subtract rate (the arithmetic the exact
 * path is restricted to) and fused DFMA rate, in GFLOP/s. */
int dz_measure_fp64_peak(int device, double *mul_sub_gflops, double *fma_gflops);

#ifdef __cplusplus
}
#endif
#endif
