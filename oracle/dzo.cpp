// dzo.cpp -- CPU ORACLE: a restatement of dantzig's simplex hot path.
//
// TEST INFRASTRUCTURE ONLY (see dzo.h).  Parity status: PINNED against the
// reference's own known-answer tests (tests/test_oracle_kat.py).
//
// Every function names the reference lines it restates.  The code is written
// for fidelity, not speed: the LITERAL variant performs exactly the floating
// point operations of the reference in exactly its order; the SKIP variant
// elides only operations that are provably no-ops (an exact-zero factor whose
// co-factor is finite).  Build with -ffp-contract=off (oracle/Makefile).

#include "dzo.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

thread_local double g_lu_flops = 0.0, g_solve_flops = 0.0, g_other_flops = 0.0;

// ---------------------------------------------------------------------------
// linalg.rs
// ---------------------------------------------------------------------------

// Matrix::factorize, linalg.rs:88-128.  Row-major n x n, in place.
// p receives the n-1 pivot rows.  `skip` selects the SKIP variant.
void factorize(double *a, int n, int *p, bool skip) {
    if (n <= 0) return;
    std::vector<int> nzcols;
    double flops = 0.0;
    for (int k = 0; k + 1 < n; ++k) {
        // linalg.rs:98-105: first row with the largest |a_ik|, strict '>'
        int mu = k;
        double magnitude = std::fabs(a[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double cand = std::fabs(a[(size_t)i * n + k]);
            if (cand > magnitude) {
                mu = i;
                magnitude = cand;
            }
        }
        // linalg.rs:107-113: swap columns k..n of the two rows only
        if (mu != k) {
            for (int j = k; j < n; ++j) std::swap(a[(size_t)mu * n + j], a[(size_t)k * n + j]);
        }
        p[k] = mu;
        // linalg.rs:116-125
        const double pivot = a[(size_t)k * n + k];
        if (pivot != 0.0) {
            const double *rk = a + (size_t)k * n;
            if (!skip) {
                for (int i = k + 1; i < n; ++i) {
                    double *ri = a + (size_t)i * n;
                    ri[k] /= pivot;
                    const double l = ri[k];
                    for (int j = k + 1; j < n; ++j) {
                        const double adjustment = l * rk[j];
                        ri[j] -= adjustment;
                    }
                }
                flops += (double)(n - k - 1) * (1.0 + 2.0 * (n - k - 1));
            } else {
                // columns of the pivot row that can change anything
                nzcols.clear();
                bool row_finite = true;
                for (int j = k + 1; j < n; ++j) {
                    if (rk[j] != 0.0) nzcols.push_back(j);
                    if (!std::isfinite(rk[j])) row_finite = false;
                }
                for (int i = k + 1; i < n; ++i) {
                    double *ri = a + (size_t)i * n;
                    if (ri[k] == 0.0 && row_finite) continue; // l = +-0, finite row: no-op
                    ri[k] /= pivot;
                    const double l = ri[k];
                    flops += 1.0;
                    if (std::isfinite(l) && row_finite) {
                        for (int j : nzcols) {
                            const double adjustment = l * rk[j];
                            ri[j] -= adjustment;
                        }
                        flops += 2.0 * (double)nzcols.size();
                    } else {
                        for (int j = k + 1; j < n; ++j) {
                            const double adjustment = l * rk[j];
                            ri[j] -= adjustment;
                        }
                        flops += 2.0 * (n - k - 1);
                    }
                }
            }
        }
    }
    g_lu_flops += flops;
}

// LU::solve, linalg.rs:282-299.
void lu_apply(const double *a, const int *p, int n, double *b, bool skip) {
    double flops = 0.0;
    for (int k = 0; k + 1 < n; ++k) {
        std::swap(b[k], b[p[k]]);
        const double bk = b[k];
        if (!skip || !std::isfinite(bk)) {
            for (int i = k + 1; i < n; ++i) b[i] -= bk * a[(size_t)i * n + k];
            flops += 2.0 * (n - k - 1);
        } else if (bk != 0.0) {
            for (int i = k + 1; i < n; ++i) {
                const double l = a[(size_t)i * n + k];
                if (l != 0.0) {
                    b[i] -= bk * l;
                    flops += 2.0;
                }
            }
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        const double *ri = a + (size_t)i * n;
        double s = b[i];
        if (!skip) {
            for (int j = i + 1; j < n; ++j) s -= ri[j] * b[j];
            flops += 2.0 * (n - i - 1);
        } else {
            for (int j = i + 1; j < n; ++j) {
                const double u = ri[j], bj = b[j];
                if ((u == 0.0 || bj == 0.0) && std::isfinite(bj) && std::isfinite(u)) continue;
                s -= u * bj;
                flops += 2.0;
            }
        }
        b[i] = s / ri[i];
        flops += 1.0;
    }
    g_solve_flops += flops;
}

// lu_solve, linalg.rs:8-10.  n == 0 mirrors the `n - 1` underflow panic
// (linalg.rs:95): reported to the caller as false.
bool lu_solve(std::vector<double> &a, int n, std::vector<double> &b, bool skip) {
    if (n <= 0) return false;
    std::vector<int> p((size_t)(n > 1 ? n - 1 : 0));
    factorize(a.data(), n, p.data(), skip);
    lu_apply(a.data(), p.data(), n, b.data(), skip);
    return true;
}

// SPARSE variant of lu_solve (linalg.rs:8-10, 88-128, 282-299): the same floating
// point operations in the same order as the SKIP variant, on rows stored as ordered
// maps instead of a dense n x n array, so that a 70 000-row basis with two million
// nonzeros (BASELINE configs[3] at full size) can be followed on a CPU at all.
//   * row OBJECTS keep their identity; rowAt/posOf record the interchanges
//     (linalg.rs:107-114 swaps columns k..n of two rows: the L part stays, which
//     is the same as never moving it);
//   * the right-hand side rides along with its row object, which performs the
//     forward half of LU::solve (linalg.rs:286-291) inside the elimination with
//     the same products b_k * l_ik in the same order;
//   * absent entries are exact zeros; every skip is one the SKIP variant makes.
// Returns false when a non-finite value shows up in a place where SKIP falls back
// to dense rows (the caller then uses the dense SKIP variant, or gives up).
struct SparseRows {
    int n = 0;
    std::vector<std::map<int, double>> rows; // row object -> (column -> value)
    std::vector<double> rhs;                 // per row object
    std::vector<std::vector<int>> col_rows;  // column -> row objects that ever held an entry there
};

bool lu_solve_sparse(SparseRows &S, std::vector<double> &out) {
    const int n = S.n;
    if (n <= 0) return false;
    std::vector<int> rowAt((size_t)n), posOf((size_t)n);
    for (int i = 0; i < n; ++i) rowAt[i] = posOf[i] = i;
    std::vector<std::pair<int, int>> cand; // (position, row object)
    double flops = 0.0, sflops = 0.0;
    for (int k = 0; k + 1 < n; ++k) {
        cand.clear();
        for (int r : S.col_rows[k])
            if (posOf[r] >= k && S.rows[r].count(k)) cand.emplace_back(posOf[r], r);
        std::sort(cand.begin(), cand.end());
        // linalg.rs:98-105: first row with the largest |a_ik|, strict '>'
        const int inc = rowAt[k];
        int mu = k;
        {
            auto it = S.rows[inc].find(k);
            double magnitude = std::fabs(it == S.rows[inc].end() ? 0.0 : it->second);
            for (auto &pc : cand) {
                if (pc.second == inc) continue; // the incumbent at position k (it may have to be eliminated below)
                const double c = std::fabs(S.rows[pc.second][k]);
                if (c > magnitude) {
                    mu = pc.first;
                    magnitude = c;
                }
            }
        }
        if (mu != k) { // linalg.rs:107-114
            const int rk = rowAt[k], rm = rowAt[mu];
            rowAt[k] = rm;
            rowAt[mu] = rk;
            posOf[rm] = k;
            posOf[rk] = mu;
        }
        const int pr = rowAt[k];
        auto &prow = S.rows[pr];
        auto pit = prow.find(k);
        const double pivot = pit == prow.end() ? 0.0 : pit->second;
        if (pivot == 0.0) continue; // linalg.rs:117
        if (!std::isfinite(pivot) || !std::isfinite(S.rhs[pr])) return false;
        for (auto it = prow.upper_bound(k); it != prow.end(); ++it)
            if (!std::isfinite(it->second)) return false;
        const double bk = S.rhs[pr];
        for (auto &pc : cand) {
            const int r = pc.second;
            if (r == pr) continue;
            auto &row = S.rows[r];
            auto vit = row.find(k);
            const double v = vit->second;
            row.erase(vit); // the L part is never read again (the rhs rides along)
            if (v == 0.0) continue;
            const double l = v / pivot;
            flops += 1.0;
            if (!std::isfinite(l)) return false;
            for (auto it = prow.upper_bound(k); it != prow.end(); ++it) {
                if (it->second == 0.0) continue;
                const double adjustment = l * it->second;
                auto e = row.find(it->first);
                if (e == row.end()) {
                    row.emplace(it->first, 0.0 - adjustment);
                    S.col_rows[it->first].push_back(r);
                } else {
                    e->second -= adjustment;
                }
                flops += 2.0;
            }
            if (bk != 0.0) { // linalg.rs:288-290
                S.rhs[r] -= bk * l;
                sflops += 2.0;
            }
        }
    }
    // linalg.rs:292-297
    out.assign((size_t)n, 0.0);
    bool poisoned = false; // a non-finite b_j turns every later 0 * b_j into NaN
    for (int i = n - 1; i >= 0; --i) {
        const int r = rowAt[i];
        auto &row = S.rows[r];
        double s = S.rhs[r];
        double diag = 0.0;
        if (poisoned) return false;
        for (auto it = row.lower_bound(i); it != row.end(); ++it) {
            if (it->first == i) {
                diag = it->second;
                continue;
            }
            const double u = it->second, bj = out[(size_t)it->first];
            if ((u == 0.0 || bj == 0.0) && std::isfinite(bj) && std::isfinite(u)) continue;
            s -= u * bj;
            sflops += 2.0;
        }
        out[(size_t)i] = s / diag;
        sflops += 1.0;
        if (!std::isfinite(out[(size_t)i])) poisoned = i > 0;
    }
    g_lu_flops += flops;
    g_solve_flops += sflops;
    return true;
}

// CscMatrix, linalg.rs:160-252.
struct Csc {
    int nrows = 0, ncols = 0;
    std::vector<int64_t> col_ptr;
    std::vector<int32_t> row_idx;
    std::vector<double> val;
};

} // namespace

// ---------------------------------------------------------------------------
// lowered problem: the state Simplex::new leaves behind (simplex.rs:84-112)
// ---------------------------------------------------------------------------
struct dzo_lowered {
    Csc A;
    std::vector<double> c; // objective.coefficients
    double c0 = 0.0;       // objective.constant
    std::vector<double> b; // rhs per row == initial x
    std::vector<int32_t> basis0, nonbasis0;
    // original variables in first-appearance order and their split columns
    std::vector<int32_t> orig_var, pos_index, neg_index;
};

namespace {

// Simplex::new, simplex.rs:123-224 (+ Equality::from :19-31, Objective::new
// :38-49, chain_variable_ids :51-60, sparsify :62-81, split_variables
// model.rs:11-22).  Internal ids are drawn from one counter like the global
// atomic in pyobjs.rs:8,27; only their relative order matters.
dzo_lowered *lower(const dzo_model *m) {
    const int64_t n_row_terms = m->n_rows > 0 ? m->row_ptr[m->n_rows] : 0;
    for (int t = 0; t < m->n_obj; ++t)
        if (m->obj_var[t] < 0 || m->obj_var[t] >= m->n_vars) return nullptr;
    for (int64_t t = 0; t < n_row_terms; ++t)
        if (m->row_var[t] < 0 || m->row_var[t] >= m->n_vars) return nullptr;

    int64_t next_id = m->n_vars; // ids 0..n_vars-1 are the caller's variables
    struct Row {
        std::vector<double> coefs;
        std::vector<int64_t> vars;
        double b;
        int64_t slack;
    };
    std::vector<Row> extra;
    std::unordered_map<int64_t, std::pair<int64_t, int64_t>> key;
    std::vector<int32_t> orig_order;

    auto see = [&](int32_t v) { // simplex.rs:133-151
        if (key.find(v) != key.end()) return;
        const int64_t pos = next_id++;
        const int64_t neg = next_id++;
        if (m->has_ub[v]) extra.push_back(Row{{1.0, -1.0}, {pos, neg}, m->ub[v], -1});
        if (m->has_lb[v]) extra.push_back(Row{{-1.0, 1.0}, {pos, neg}, -m->lb[v], -1});
        key.emplace(v, std::make_pair(pos, neg));
        orig_order.push_back(v);
    };
    for (int t = 0; t < m->n_obj; ++t) see(m->obj_var[t]);
    for (int64_t t = 0; t < n_row_terms; ++t) see(m->row_var[t]);

    // simplex.rs:153 objective.split_variables
    std::vector<double> obj_coefs;
    std::vector<int64_t> obj_vars;
    for (int t = 0; t < m->n_obj; ++t) {
        const auto &pn = key[m->obj_var[t]];
        obj_coefs.push_back(m->obj_coef[t]);
        obj_vars.push_back(pn.first);
        obj_coefs.push_back(-m->obj_coef[t]);
        obj_vars.push_back(pn.second);
    }
    // simplex.rs:154-166: user rows (split) chained with the bound rows, each
    // given a slack (Equality::from)
    std::vector<Row> rows;
    for (int r = 0; r < m->n_rows; ++r) {
        Row row;
        row.b = m->rhs[r];
        for (int64_t t = m->row_ptr[r]; t < m->row_ptr[r + 1]; ++t) {
            const auto &pn = key[m->row_var[t]];
            row.coefs.push_back(m->row_coef[t]);
            row.vars.push_back(pn.first);
            row.coefs.push_back(-m->row_coef[t]);
            row.vars.push_back(pn.second);
        }
        rows.push_back(std::move(row));
    }
    for (auto &e : extra) rows.push_back(e);
    std::unordered_map<int64_t, double> slack_ids;
    for (auto &row : rows) {
        row.slack = next_id++;
        row.coefs.push_back(1.0);
        row.vars.push_back(row.slack);
        slack_ids[row.slack] = row.b;
    }
    // simplex.rs:168-176 index assignment in first-seen order
    std::vector<int64_t> index_to_id;
    std::unordered_map<int64_t, int32_t> id_to_index;
    auto assign = [&](int64_t id) {
        if (id_to_index.find(id) == id_to_index.end()) {
            id_to_index.emplace(id, (int32_t)index_to_id.size());
            index_to_id.push_back(id);
        }
    };
    for (int64_t id : obj_vars) assign(id);
    for (auto &row : rows)
        for (int64_t id : row.vars) assign(id);

    auto *lp = new dzo_lowered();
    const int n_int = (int)index_to_id.size();
    const int mm = (int)rows.size();
    // Objective::new, simplex.rs:38-49 (later duplicate overwrites)
    lp->c.assign((size_t)n_int, 0.0);
    for (size_t t = 0; t < obj_vars.size(); ++t) lp->c[id_to_index[obj_vars[t]]] = obj_coefs[t];
    lp->c0 = m->obj_const;
    // simplex.rs:190-201
    lp->b.reserve(mm);
    for (int i = 0; i < n_int; ++i) {
        auto it = slack_ids.find(index_to_id[i]);
        if (it != slack_ids.end()) {
            lp->basis0.push_back(i);
            lp->b.push_back(it->second);
        } else {
            lp->nonbasis0.push_back(i);
        }
    }
    // sparsify, simplex.rs:62-81: Matrix::coords (later duplicate overwrites,
    // linalg.rs:32-38) then to_sparse (drops exact zeros, rows ascending
    // within a column, linalg.rs:254-270).  Built per column, never dense.
    std::vector<std::map<int32_t, double>> cols((size_t)n_int);
    for (int i = 0; i < mm; ++i)
        for (size_t t = 0; t < rows[i].vars.size(); ++t)
            cols[id_to_index[rows[i].vars[t]]][i] = rows[i].coefs[t];
    lp->A.nrows = mm;
    lp->A.ncols = n_int;
    lp->A.col_ptr.push_back(0);
    for (int j = 0; j < n_int; ++j) {
        for (auto &kv : cols[j]) {
            if (kv.second != 0.0) {
                lp->A.row_idx.push_back(kv.first);
                lp->A.val.push_back(kv.second);
            }
        }
        lp->A.col_ptr.push_back((int64_t)lp->A.val.size());
    }
    for (int32_t v : orig_order) {
        lp->orig_var.push_back(v);
        lp->pos_index.push_back(id_to_index[key[v].first]);
        lp->neg_index.push_back(id_to_index[key[v].second]);
    }
    return lp;
}

// safe_divide, simplex.rs:464-468.  Returns false where the reference panics.
bool safe_divide(double x, double y, double *out) {
    const double div = (x == 0.0 && y == 0.0) ? 0.0 : x / y;
    *out = div;
    return !(std::isinf(div) || std::isnan(div));
}

// fn pivot, simplex.rs:410-421
void pivot_update(std::vector<double> &data, const std::vector<double> &delta, int index,
                  double step_length) {
    for (size_t i = 0; i < data.size(); ++i) {
        if ((int)i == index)
            data[i] = step_length;
        else
            data[i] -= step_length * delta[i];
    }
    g_other_flops += 2.0 * (double)data.size();
}

// find_first_pivot, simplex.rs:423-437.  Returns the POSITION (the reference
// returns index_lookup[position]); -1 for None.
int find_first_pivot(const std::vector<double> &y, const std::vector<double> &y_bar) {
    int best = -1;
    double best_ratio = 0.0;
    for (size_t k = 0; k < y.size(); ++k) {
        if (!(y_bar[k] > 0.0)) continue;
        const double ratio = -y[k] / y_bar[k];
        if (best < 0) {
            best = (int)k;
            best_ratio = ratio;
        } else if (ratio > best_ratio) {
            best = (int)k;
            best_ratio = ratio;
        }
    }
    g_other_flops += (double)y.size();
    return best;
}

// find_second_pivot, simplex.rs:439-461.  Returns the position; -1 for None.
int find_second_pivot(double mu, const std::vector<double> &y, const std::vector<double> &y_bar,
                      const std::vector<double> &dy) {
    int best = -1;
    double best_ratio = 0.0;
    for (size_t k = 0; k < y.size(); ++k) {
        const double scaled = mu * y_bar[k];
        const double denominator = y[k] + scaled;
        const double ratio = dy[k] / denominator;
        if (!(ratio > 0.0)) continue;
        if (best < 0) {
            best = (int)k;
            best_ratio = ratio;
        } else if (ratio > best_ratio) {
            best = (int)k;
            best_ratio = ratio;
        }
    }
    g_other_flops += 3.0 * (double)y.size();
    return best;
}

struct Solver {
    const dzo_lowered *lp;
    bool skip;
    bool sparse = false; // DZO_SPARSE: sparse rows, falling back to dense SKIP on non-finite values
    int M, Nn;
    std::vector<int32_t> b, n;
    std::vector<double> x, x_bar, z, z_bar;
    std::vector<double> dense, rhs, dx, dz;

    explicit Solver(const dzo_lowered *lp_, bool skip_) : lp(lp_), skip(skip_) {
        M = lp->A.nrows;
        b = lp->basis0;
        n = lp->nonbasis0;
        Nn = (int)n.size();
        x = lp->b; // simplex.rs:195
        z.resize((size_t)Nn);
        for (int k = 0; k < Nn; ++k) z[k] = -lp->c[n[k]]; // simplex.rs:199
        x_bar.assign((size_t)M, 1.0);                     // simplex.rs:204-205
        z_bar.assign((size_t)Nn, 1.0);
    }

    // basis_matrix + to_dense (simplex.rs:270-272, linalg.rs:188-192,131-140),
    // optionally transposed (linalg.rs:40-48).
    void build_basis(bool transposed) {
        dense.assign((size_t)M * M, 0.0);
        for (int p = 0; p < M; ++p) {
            const int col = b[p];
            for (int64_t e = lp->A.col_ptr[col]; e < lp->A.col_ptr[col + 1]; ++e) {
                const int r = lp->A.row_idx[e];
                if (transposed)
                    dense[(size_t)p * M + r] = lp->A.val[e];
                else
                    dense[(size_t)r * M + p] = lp->A.val[e];
            }
        }
    }
    // The same system as build_basis + rhs, as sparse row objects (DZO_SPARSE).
    bool sparse_solve(bool transposed, int arg) {
        SparseRows S;
        S.n = M;
        S.rows.resize((size_t)M);
        S.rhs.assign((size_t)M, 0.0);
        S.col_rows.resize((size_t)M);
        for (int p = 0; p < M; ++p) {
            const int col = b[p];
            for (int64_t e = lp->A.col_ptr[col]; e < lp->A.col_ptr[col + 1]; ++e) {
                const int r = lp->A.row_idx[e];
                const int wr = transposed ? p : r, wc = transposed ? r : p;
                S.rows[(size_t)wr][wc] = lp->A.val[e];
                S.col_rows[(size_t)wc].push_back(wr);
            }
        }
        if (transposed) {
            S.rhs[(size_t)arg] = 1.0;
        } else {
            for (int64_t e = lp->A.col_ptr[arg]; e < lp->A.col_ptr[arg + 1]; ++e)
                S.rhs[(size_t)lp->A.row_idx[e]] = lp->A.val[e];
        }
        return lu_solve_sparse(S, rhs);
    }
    // solve_for_dx, simplex.rs:226-229
    bool solve_for_dx(int j) {
        if (sparse && M > 0) {
            const double f0 = g_lu_flops, f1 = g_solve_flops;
            if (sparse_solve(false, j)) {
                dx = rhs;
                return true;
            }
            g_lu_flops = f0, g_solve_flops = f1;
            if ((int64_t)M * M > (int64_t)1 << 28) return false; // no dense fallback at this size
        }
        rhs.assign((size_t)M, 0.0);
        for (int64_t e = lp->A.col_ptr[j]; e < lp->A.col_ptr[j + 1]; ++e)
            rhs[lp->A.row_idx[e]] = lp->A.val[e];
        build_basis(false);
        if (!lu_solve(dense, M, rhs, skip)) return false;
        dx = rhs;
        return true;
    }
    // solve_for_dz, simplex.rs:231-236 + neg_t_dot linalg.rs:199-207
    bool solve_for_dz(int pos_i) {
        bool done = false;
        if (sparse && M > 0) {
            const double f0 = g_lu_flops, f1 = g_solve_flops;
            done = sparse_solve(true, pos_i);
            if (!done) {
                g_lu_flops = f0, g_solve_flops = f1;
                if ((int64_t)M * M > (int64_t)1 << 28) return false;
            }
        }
        if (!done) {
            rhs.assign((size_t)M, 0.0);
            rhs[pos_i] = 1.0;
            build_basis(true);
            if (!lu_solve(dense, M, rhs, skip)) return false;
        }
        dz.resize((size_t)Nn);
        double flops = 0.0;
        for (int k = 0; k < Nn; ++k) {
            const int col = n[k];
            double s = 0.0;
            for (int64_t e = lp->A.col_ptr[col]; e < lp->A.col_ptr[col + 1]; ++e)
                s += lp->A.val[e] * -rhs[lp->A.row_idx[e]];
            flops += 2.0 * (double)(lp->A.col_ptr[col + 1] - lp->A.col_ptr[col]);
            dz[k] = s;
        }
        g_other_flops += flops;
        return true;
    }
    // Simplex::pivot + swap, simplex.rs:239-268.  p, q are positions.
    bool do_pivot(int p, int q) {
        double t, s, t_bar, s_bar;
        if (!safe_divide(x[p], dx[p], &t)) return false;
        if (!safe_divide(z[q], dz[q], &s)) return false;
        if (!safe_divide(x_bar[p], dx[p], &t_bar)) return false;
        if (!safe_divide(z_bar[q], dz[q], &s_bar)) return false;
        pivot_update(x, dx, p, t);
        pivot_update(x_bar, dx, p, t_bar);
        pivot_update(z, dz, q, s);
        pivot_update(z_bar, dz, q, s_bar);
        std::swap(b[p], n[q]);
        return true;
    }
};

inline uint64_t trace_mix(uint64_t h, int kind, int i, int j) {
    const uint64_t w = (uint64_t)(uint32_t)kind | ((uint64_t)(uint32_t)i << 1) |
                       ((uint64_t)(uint32_t)j << 32);
    return (h ^ w) * 0x100000001b3ULL;
}

} // namespace

extern "C" {

dzo_lowered *dzo_lower(const dzo_model *model) { return lower(model); }
void dzo_lowered_free(dzo_lowered *lp) { delete lp; }

dzo_lowered *dzo_lowered_from_arrays(int32_t m, int32_t n_int, const int64_t *col_ptr,
                                     const int32_t *row_idx, const double *val, const double *c,
                                     double c0, const double *b, const int32_t *basis,
                                     const int32_t *nonbasis) {
    auto *lp = new dzo_lowered();
    lp->A.nrows = m;
    lp->A.ncols = n_int;
    lp->A.col_ptr.assign(col_ptr, col_ptr + n_int + 1);
    const int64_t nnz = col_ptr[n_int];
    lp->A.row_idx.assign(row_idx, row_idx + nnz);
    lp->A.val.assign(val, val + nnz);
    lp->c.assign(c, c + n_int);
    lp->c0 = c0;
    lp->b.assign(b, b + m);
    lp->basis0.assign(basis, basis + m);
    lp->nonbasis0.assign(nonbasis, nonbasis + (n_int - m));
    return lp;
}

void dzo_lowered_dims(const dzo_lowered *lp, int32_t *m, int32_t *n_int, int64_t *nnz,
                      int32_t *n_orig) {
    if (m) *m = lp->A.nrows;
    if (n_int) *n_int = lp->A.ncols;
    if (nnz) *nnz = (int64_t)lp->A.val.size();
    if (n_orig) *n_orig = (int32_t)lp->orig_var.size();
}

void dzo_lowered_get(const dzo_lowered *lp, int64_t *col_ptr, int32_t *row_idx, double *val,
                     double *c, double *c0, double *b, int32_t *basis0, int32_t *nonbasis0,
                     int32_t *orig_var, int32_t *pos_index, int32_t *neg_index) {
    auto cp = [](auto *dst, const auto &src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(src[0]));
    };
    cp(col_ptr, lp->A.col_ptr);
    cp(row_idx, lp->A.row_idx);
    cp(val, lp->A.val);
    cp(c, lp->c);
    if (c0) *c0 = lp->c0;
    cp(b, lp->b);
    cp(basis0, lp->basis0);
    cp(nonbasis0, lp->nonbasis0);
    cp(orig_var, lp->orig_var);
    cp(pos_index, lp->pos_index);
    cp(neg_index, lp->neg_index);
}

// Simplex::solve (simplex.rs:332-343) with status (:274-306), primal_step
// (:308-318) and dual_step (:320-330) unrolled into a loop; the recursion in
// the reference carries no state besides `self`.
int dzo_solve(const dzo_lowered *lp, int variant, int64_t max_pivots, dzo_result *res,
              double *x_basic, int32_t *basis, double *values, int32_t *trace,
              int64_t trace_cap) {
    g_lu_flops = g_solve_flops = g_other_flops = 0.0;
    Solver s(lp, variant == DZO_SKIP || variant == DZO_SPARSE);
    s.sparse = variant == DZO_SPARSE;
    int status = DZO_OPTIMAL;
    int64_t pivots = 0, n_primal = 0, n_dual = 0;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (;;) {
        // status(), simplex.rs:274-306
        const int q0 = find_first_pivot(s.z, s.z_bar);
        const int p0 = find_first_pivot(s.x, s.x_bar);
        bool primal_step;
        double mu;
        if (q0 >= 0 && p0 >= 0) {
            const double primal = -s.x[p0] / s.x_bar[p0];
            const double dual = -s.z[q0] / s.z_bar[q0];
            if (primal <= 1e-12 && dual <= 1e-12) break; // Optimal
            if (primal < dual) {
                primal_step = true;
                mu = dual;
            } else {
                primal_step = false;
                mu = primal;
            }
        } else if (q0 >= 0) {
            primal_step = true;
            mu = -s.z[q0] / s.z_bar[q0];
        } else if (p0 >= 0) {
            primal_step = false;
            mu = -s.x[p0] / s.x_bar[p0];
        } else {
            status = DZO_PANIC; // simplex.rs:304
            break;
        }
        if (max_pivots > 0 && pivots >= max_pivots) {
            status = DZO_PIVOT_CAP;
            break;
        }
        int p, q;
        if (primal_step) { // simplex.rs:308-318
            q = q0;
            if (!s.solve_for_dx(s.n[q])) {
                status = DZO_PANIC;
                break;
            }
            p = find_second_pivot(mu, s.x, s.x_bar, s.dx);
            if (p < 0) {
                status = DZO_UNBOUNDED;
                break;
            }
            if (!s.solve_for_dz(p)) {
                status = DZO_PANIC;
                break;
            }
        } else { // simplex.rs:320-330
            p = p0;
            if (!s.solve_for_dz(p)) {
                status = DZO_PANIC;
                break;
            }
            q = find_second_pivot(mu, s.z, s.z_bar, s.dz);
            if (q < 0) {
                status = DZO_INFEASIBLE;
                break;
            }
            if (!s.solve_for_dx(s.n[q])) {
                status = DZO_PANIC;
                break;
            }
        }
        const int leaving = s.b[p], entering = s.n[q];
        if (!s.do_pivot(p, q)) {
            status = DZO_PANIC; // safe_divide assert, simplex.rs:466
            break;
        }
        if (trace && pivots < trace_cap) {
            trace[3 * pivots + 0] = primal_step ? 0 : 1;
            trace[3 * pivots + 1] = leaving;
            trace[3 * pivots + 2] = entering;
        }
        h = trace_mix(h, primal_step ? 0 : 1, leaving, entering);
        ++pivots;
        if (primal_step)
            ++n_primal;
        else
            ++n_dual;
    }
    // objective_value, simplex.rs:345-352 (position order; the reference's
    // HashMap order is unspecified)
    double obj = 0.0;
    for (int p = 0; p < s.M; ++p) obj += lp->c[s.b[p]] * s.x[p];
    obj = lp->c0 + obj;
    if (res) {
        res->status = status;
        res->pivots = pivots;
        res->n_primal = n_primal;
        res->n_dual = n_dual;
        res->trace_hash = h;
        res->objective = obj;
    }
    if (x_basic) std::memcpy(x_basic, s.x.data(), sizeof(double) * (size_t)s.M);
    if (basis) std::memcpy(basis, s.b.data(), sizeof(int32_t) * (size_t)s.M);
    if (values) { // solution(), simplex.rs:354-371
        std::vector<int> where((size_t)lp->A.ncols, -1);
        for (int p = 0; p < s.M; ++p) where[s.b[p]] = p;
        for (size_t v = 0; v < lp->orig_var.size(); ++v) {
            const int wp = where[lp->pos_index[v]], wn = where[lp->neg_index[v]];
            const double pos = wp >= 0 ? s.x[wp] : 0.0;
            const double neg = wn >= 0 ? s.x[wn] : 0.0;
            values[v] = pos - neg;
        }
    }
    return 0;
}

void dzo_lu_factorize(double *a, int32_t n, int32_t *p) { factorize(a, n, p, false); }

void dzo_lu_solve(double *a, int32_t n, double *b, int variant) {
    std::vector<int> p((size_t)(n > 1 ? n - 1 : 0));
    factorize(a, n, p.data(), variant == DZO_SKIP);
    lu_apply(a, p.data(), n, b, variant == DZO_SKIP);
}

void dzo_neg_t_dot(int32_t nrows, int32_t ncols, const int64_t *col_ptr, const int32_t *row_idx,
                   const double *val, const double *v, double *out) {
    (void)nrows;
    for (int j = 0; j < ncols; ++j) {
        double s = 0.0;
        for (int64_t e = col_ptr[j]; e < col_ptr[j + 1]; ++e) s += val[e] * -v[row_idx[e]];
        out[j] = s;
    }
}

int64_t dzo_dense_to_csc(const double *dense, int32_t nrows, int32_t ncols, int64_t *col_ptr,
                         int32_t *row_idx, double *val) {
    int64_t nnz = 0;
    col_ptr[0] = 0;
    for (int j = 0; j < ncols; ++j) {
        for (int i = 0; i < nrows; ++i) {
            const double v = dense[(size_t)i * ncols + j];
            if (v != 0.0) {
                row_idx[nnz] = i;
                val[nnz] = v;
                ++nnz;
            }
        }
        col_ptr[j + 1] = nnz;
    }
    return nnz;
}

void dzo_last_flops(double *lu_flops, double *solve_flops, double *other_flops) {
    if (lu_flops) *lu_flops = g_lu_flops;
    if (solve_flops) *solve_flops = g_solve_flops;
    if (other_flops) *other_flops = g_other_flops;
}

} // extern "C"
