// dz_lower.cpp -- host-side lowering to computational form (product code).
//
// Replaces Simplex::new (/root/reference/src/simplex.rs:123-224) and its
// helpers (Equality::from :19-31, Objective::new :38-49, chain_variable_ids
// :51-60, sparsify :62-81, LinExpr::split_variables model.rs:11-22).
//
// The lowering is purely STRUCTURAL: which lowered column a term lands in and
// in what order is decided by first-appearance order alone, never by a value.
// So instead of numbers the template stores signed references into a parameter
// vector theta (layout in include/dantzig_b200.h); one template then serves a
// whole batch of LPs that share the structure, and the device reads
// val = +-theta[ref].  Exact zeros, which the reference drops when it builds
// its CSC (linalg.rs:254-270), are skipped by value on the device.
//
// Unlike the reference, nothing is ever densified (simplex.rs:80 allocates
// m_int x n_int doubles): the CSC pattern is assembled column by column.

#include "dz_internal.h"

#include <algorithm>
#include <cstring>
#include <string>

namespace dz {

static inline int32_t mkref(int64_t index, bool negate) {
    return (int32_t)((index << 1) | (negate ? 1 : 0));
}

// FNV-1a over the arrays that decide the lowered structure (pack_theta refuses a model
// whose numbers would land on another structure's columns).
static uint64_t structure_hash(const dz_model *m) {
    uint64_t h = 0xcbf29ce484222325ULL;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 0x100000001b3ULL;
    };
    const int64_t T = m->n_rows > 0 ? m->row_ptr[m->n_rows] : 0;
    for (int32_t v = 0; v < m->n_vars; ++v) {
        const unsigned char f[2] = {(unsigned char)(m->has_lb[v] != 0), (unsigned char)(m->has_ub[v] != 0)};
        mix(f, 2);
    }
    if (m->n_obj) mix(m->obj_var, sizeof(int32_t) * (size_t)m->n_obj);
    if (m->n_rows) mix(m->row_ptr, sizeof(int64_t) * ((size_t)m->n_rows + 1));
    if (T) mix(m->row_var, sizeof(int32_t) * (size_t)T);
    return h;
}

int build_template(const dz_model *m, Template *t, std::string *err) {
    if (!m || m->n_vars < 0 || m->n_obj < 0 || m->n_rows < 0) {
        *err = "dz_template_create: negative count in model";
        return DZ_ERR_ARG;
    }
    if ((m->n_vars > 0 && (!m->has_lb || !m->has_ub)) || (m->n_obj > 0 && !m->obj_var) ||
        (m->n_rows > 0 && !m->row_ptr)) {
        *err = "dz_template_create: missing structural array";
        return DZ_ERR_ARG;
    }
    const int64_t T = m->n_rows > 0 ? m->row_ptr[m->n_rows] : 0;
    if (T < 0 || (T > 0 && !m->row_var)) {
        *err = "dz_template_create: bad row_ptr";
        return DZ_ERR_ARG;
    }
    for (int32_t r = 0; r < m->n_rows; ++r)
        if (m->row_ptr[r] > m->row_ptr[r + 1] || m->row_ptr[r] < 0) {
            *err = "dz_template_create: row_ptr not monotone";
            return DZ_ERR_ARG;
        }
    for (int32_t k = 0; k < m->n_obj; ++k)
        if (m->obj_var[k] < 0 || m->obj_var[k] >= m->n_vars) {
            *err = "dz_template_create: objective variable out of range";
            return DZ_ERR_ARG;
        }
    for (int64_t k = 0; k < T; ++k)
        if (m->row_var[k] < 0 || m->row_var[k] >= m->n_vars) {
            *err = "dz_template_create: row variable out of range";
            return DZ_ERR_ARG;
        }

    // theta layout
    t->n_vars = m->n_vars;
    t->n_obj = m->n_obj;
    t->n_rows_user = m->n_rows;
    t->n_row_terms = T;
    t->off_obj = 2;
    t->off_rowcoef = t->off_obj + m->n_obj;
    t->off_rhs = t->off_rowcoef + T;
    t->off_lb = t->off_rhs + m->n_rows;
    t->off_ub = t->off_lb + m->n_vars;
    t->n_theta = t->off_ub + m->n_vars;
    if (t->n_theta >= (int64_t(1) << 30)) {
        *err = "dz_template_create: model too large for 32-bit parameter references";
        return DZ_ERR_LIMIT;
    }

    // (1) original variables in first-appearance order: objective terms, then
    // the rows' terms (simplex.rs:126-151).
    std::vector<int32_t> seen_rank((size_t)m->n_vars, -1);
    t->orig_var.clear();
    auto see = [&](int32_t v) {
        if (seen_rank[v] < 0) {
            seen_rank[v] = (int32_t)t->orig_var.size();
            t->orig_var.push_back(v);
        }
    };
    for (int32_t k = 0; k < m->n_obj; ++k) see(m->obj_var[k]);
    for (int64_t k = 0; k < T; ++k) see(m->row_var[k]);
    const int32_t n_orig = (int32_t)t->orig_var.size();

    // (2) bound rows, in that same order, upper before lower (simplex.rs:141-148)
    struct BoundRow {
        int32_t rank;
        bool upper;
    };
    std::vector<BoundRow> bound_rows;
    for (int32_t k = 0; k < n_orig; ++k) {
        const int32_t v = t->orig_var[k];
        if (m->has_ub[v]) bound_rows.push_back({k, true});
        if (m->has_lb[v]) bound_rows.push_back({k, false});
    }
    const int64_t M64 = (int64_t)m->n_rows + (int64_t)bound_rows.size();
    const int64_t N64 = 2 * (int64_t)n_orig + M64;
    if (N64 > 0x7fffff00LL) {
        *err = "dz_template_create: lowered dimensions exceed 32-bit indices";
        return DZ_ERR_LIMIT;
    }
    const int32_t M = (int32_t)M64, Nint = (int32_t)N64;

    // (3) lowered column numbering: first-seen order over the split objective
    // terms, then row by row the split terms followed by that row's slack
    // (simplex.rs:168-176).  pos/neg of original variable rank k are the
    // "virtual" columns 2k, 2k+1 until numbered.
    std::vector<int32_t> col_of_virtual((size_t)2 * n_orig, -1);
    int32_t next_col = 0;
    auto number = [&](int32_t virt) {
        if (col_of_virtual[virt] < 0) col_of_virtual[virt] = next_col++;
        return col_of_virtual[virt];
    };
    std::vector<std::vector<std::pair<int32_t, int32_t>>> cols((size_t)Nint); // (row, ref)
    auto put = [&](int32_t col, int32_t row, int32_t ref) {
        auto &c = cols[col];
        if (!c.empty() && c.back().first == row)
            c.back().second = ref; // Matrix::coords: later duplicate wins (linalg.rs:34-36)
        else
            c.emplace_back(row, ref);
    };
    t->c_ref.assign((size_t)Nint, -1);
    for (int32_t k = 0; k < m->n_obj; ++k) {
        const int32_t rank = seen_rank[m->obj_var[k]];
        const int32_t cp = number(2 * rank), cn = number(2 * rank + 1);
        t->c_ref[cp] = mkref(t->off_obj + k, false); // Objective::new overwrites (simplex.rs:41-43)
        t->c_ref[cn] = mkref(t->off_obj + k, true);
    }
    t->b_ref.assign((size_t)M, -1);
    t->basis0.assign((size_t)M, -1);
    for (int32_t r = 0; r < m->n_rows; ++r) {
        for (int64_t k = m->row_ptr[r]; k < m->row_ptr[r + 1]; ++k) {
            const int32_t rank = seen_rank[m->row_var[k]];
            const int32_t cp = number(2 * rank), cn = number(2 * rank + 1);
            put(cp, r, mkref(t->off_rowcoef + k, false));
            put(cn, r, mkref(t->off_rowcoef + k, true));
        }
        const int32_t slack = next_col++;
        put(slack, r, mkref(0, false));
        t->b_ref[r] = mkref(t->off_rhs + r, false);
        t->basis0[r] = slack;
    }
    for (size_t e = 0; e < bound_rows.size(); ++e) {
        const int32_t r = m->n_rows + (int32_t)e;
        const int32_t rank = bound_rows[e].rank;
        const int32_t v = t->orig_var[rank];
        const int32_t cp = number(2 * rank), cn = number(2 * rank + 1);
        if (bound_rows[e].upper) { //  pos - neg <= ub
            put(cp, r, mkref(0, false));
            put(cn, r, mkref(0, true));
            t->b_ref[r] = mkref(t->off_ub + v, false);
        } else { // -pos + neg <= -lb
            put(cp, r, mkref(0, true));
            put(cn, r, mkref(0, false));
            t->b_ref[r] = mkref(t->off_lb + v, true);
        }
        const int32_t slack = next_col++;
        put(slack, r, mkref(0, false));
        t->basis0[r] = slack;
    }
    if (next_col != Nint) {
        *err = "dz_template_create: internal column count mismatch";
        return DZ_ERR_ARG;
    }

    // (4) CSC pattern + references; rows ascend within a column by construction.
    t->m = M;
    t->n_int = Nint;
    t->col_ptr.assign((size_t)Nint + 1, 0);
    int64_t nnz = 0;
    for (int32_t j = 0; j < Nint; ++j) nnz += (int64_t)cols[j].size();
    if (nnz > 0x7fffff00LL) {
        *err = "dz_template_create: too many nonzeros for 32-bit offsets";
        return DZ_ERR_LIMIT;
    }
    t->row_idx.resize((size_t)nnz);
    t->val_ref.resize((size_t)nnz);
    int64_t at = 0;
    for (int32_t j = 0; j < Nint; ++j) {
        for (auto &e : cols[j]) {
            t->row_idx[at] = e.first;
            t->val_ref[at] = e.second;
            ++at;
        }
        t->col_ptr[j + 1] = at;
    }
    // (5) initial nonbasis: every non-slack column, ascending (simplex.rs:190-201)
    std::vector<uint8_t> is_slack((size_t)Nint, 0);
    for (int32_t r = 0; r < M; ++r) is_slack[t->basis0[r]] = 1;
    t->nonbasis0.clear();
    for (int32_t j = 0; j < Nint; ++j)
        if (!is_slack[j]) t->nonbasis0.push_back(j);
    t->pos_index.resize((size_t)n_orig);
    t->neg_index.resize((size_t)n_orig);
    for (int32_t k = 0; k < n_orig; ++k) {
        t->pos_index[k] = col_of_virtual[2 * k];
        t->neg_index[k] = col_of_virtual[2 * k + 1];
    }
    t->c0_ref = mkref(1, false);
    // (6) structure the kernels exploit without changing any value: slack columns (the
    // only columns that are unit vectors by construction) and exact-negative twins.
    t->slack_row.assign((size_t)Nint, -1);
    for (int32_t j = 0; j < Nint; ++j)
        if (t->col_ptr[j + 1] - t->col_ptr[j] == 1 && t->val_ref[t->col_ptr[j]] == mkref(0, false))
            t->slack_row[j] = t->row_idx[t->col_ptr[j]];
    t->twin.assign((size_t)Nint, -1);
    for (int32_t k = 0; k < n_orig; ++k) {
        const int32_t cp = t->pos_index[k], cn = t->neg_index[k];
        const int64_t a0 = t->col_ptr[cp], a1 = t->col_ptr[cp + 1], b0 = t->col_ptr[cn];
        bool same = (a1 - a0) == (t->col_ptr[cn + 1] - b0) && a1 > a0;
        for (int64_t e = 0; same && e < a1 - a0; ++e)
            same = t->row_idx[a0 + e] == t->row_idx[b0 + e] && (t->val_ref[a0 + e] ^ 1) == t->val_ref[b0 + e] &&
                   t->val_ref[a0 + e] >= 0;
        if (same && t->slack_row[cp] < 0 && t->slack_row[cn] < 0) {
            t->twin[cp] = cn;
            t->twin[cn] = cp;
        }
    }
    t->structure_hash = structure_hash(m);
    return DZ_OK;
}

int pack_theta(const Template *t, const dz_model *m, double *theta, std::string *err) {
    const int64_t T = m->n_rows > 0 ? m->row_ptr[m->n_rows] : 0;
    if (m->n_vars != t->n_vars || m->n_obj != t->n_obj || m->n_rows != t->n_rows_user ||
        T != t->n_row_terms) {
        *err = "dz_template_pack_theta: model shape differs from the template's";
        return DZ_ERR_ARG;
    }
    if ((m->n_vars > 0 && (!m->has_lb || !m->has_ub || !m->lb || !m->ub)) || (m->n_obj > 0 && (!m->obj_var || !m->obj_coef)) ||
        (m->n_rows > 0 && (!m->row_ptr || !m->rhs)) || (T > 0 && (!m->row_var || !m->row_coef))) {
        *err = "dz_template_pack_theta: missing array";
        return DZ_ERR_ARG;
    }
    if (structure_hash(m) != t->structure_hash) {
        *err = "dz_template_pack_theta: model structure (bound flags, variable indices) differs from the template's";
        return DZ_ERR_ARG;
    }
    theta[0] = 1.0;
    theta[1] = m->obj_const;
    if (m->n_obj) std::memcpy(theta + t->off_obj, m->obj_coef, sizeof(double) * (size_t)m->n_obj);
    if (T) std::memcpy(theta + t->off_rowcoef, m->row_coef, sizeof(double) * (size_t)T);
    if (m->n_rows) std::memcpy(theta + t->off_rhs, m->rhs, sizeof(double) * (size_t)m->n_rows);
    for (int32_t v = 0; v < m->n_vars; ++v) {
        theta[t->off_lb + v] = m->has_lb[v] ? m->lb[v] : 0.0;
        theta[t->off_ub + v] = m->has_ub[v] ? m->ub[v] : 0.0;
    }
    return DZ_OK;
}

} // namespace dz
