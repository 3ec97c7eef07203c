// dz_rustmod.cpp -- the drop-in replacement for the reference's compiled
// extension module `dantzig.rust` (/root/reference/src/lib.rs:29-38), written
// as host C++ above the C ABI (include/dantzig_b200.h).
//
// It exports exactly the Python-visible surface of the pyo3 module:
//   Variable(*, lb, ub)            id/lb/ub read-only; ids from a process-global
//                                   counter                      pyobjs.rs:8-31
//   PyLinExpr(coefs, vars)         map_ids_to_coefs, __neg__, __add__ (merge by
//                                   id, first-appearance order), __mul__
//                                                                pyobjs.rs:39-112
//   PyAffExpr(*, linexpr, constant) getters pylinexpr, constant   pyobjs.rs:114-133
//   PyInequality(*, linexpr, b)                                   pyobjs.rs:135-152
//   PySolution                      objective_value, __getitem__ (0.0 for an
//                                   unknown variable)             pyobjs.rs:154-175
//   solve_batch(objectives, constraints)  new: one launch per group of equal structure
//   solve(objective, constraints)   lib.rs:16-27; raises
//                                   dantzig.exceptions.UnboundedError /
//                                   InfeasibleError with the reference's text.
// The solve itself is dz_solve_model: host lowering + the CUDA kernels.  There
// is no CPU path; without a GPU solve() raises RuntimeError.  A Rust panic of
// the reference (safe_divide assert etc.) surfaces as dantzig_b200 PanicException
// (a BaseException subclass, like pyo3_runtime.PanicException).

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <cstdint>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dantzig_b200.h"

namespace py = pybind11;

namespace {

std::atomic<std::size_t> g_counter{0}; // pyobjs.rs:8

struct Variable {
    std::size_t id;
    std::optional<double> lb, ub;
    Variable(std::optional<double> lb_, std::optional<double> ub_)
        : id(g_counter.fetch_add(1, std::memory_order_relaxed)), lb(lb_), ub(ub_) {}
};

struct LinExpr {
    std::vector<double> coefs;
    std::vector<Variable> vars;
    std::unordered_map<std::size_t, std::size_t> id_to_index;

    LinExpr(std::vector<double> c, std::vector<Variable> v) : coefs(std::move(c)), vars(std::move(v)) {
        for (std::size_t i = 0; i < vars.size(); ++i) id_to_index[vars[i].id] = i; // later wins
    }
    py::dict map_ids_to_coefs() const { // pyobjs.rs:62-69 (later duplicate wins)
        py::dict d;
        for (std::size_t i = 0; i < coefs.size() && i < vars.size(); ++i)
            d[py::int_(vars[i].id)] = coefs[i];
        return d;
    }
    LinExpr neg() const { // pyobjs.rs:71-76
        LinExpr r(*this);
        for (auto &c : r.coefs) c = -c;
        return r;
    }
    LinExpr add(const LinExpr &o) const { // pyobjs.rs:78-104
        // assert_eq!(coefs.len(), vars.len()); assert_eq!(coefs.len(), id_to_index.len()) (pyobjs.rs:83-84):
        // mismatched lengths or duplicate variable ids on the left panic in the reference
        if (coefs.size() != vars.size() || coefs.size() != id_to_index.size()) {
            py::object exc = py::module_::import("dantzig.rust").attr("PanicException");
            PyErr_SetString(exc.ptr(), "assertion failed: `(left == right)` (PyLinExpr.__add__, pyobjs.rs:83-84)");
            throw py::error_already_set();
        }
        LinExpr r(*this);
        for (std::size_t i = 0; i < o.coefs.size() && i < o.vars.size(); ++i) {
            auto it = r.id_to_index.find(o.vars[i].id);
            if (it != r.id_to_index.end()) {
                r.coefs[it->second] += o.coefs[i];
            } else {
                r.id_to_index.emplace(o.vars[i].id, r.vars.size());
                r.vars.push_back(o.vars[i]);
                r.coefs.push_back(o.coefs[i]);
            }
        }
        return r;
    }
    LinExpr mul(double k) const { // pyobjs.rs:106-111, model.rs:31-36
        LinExpr r(*this);
        for (auto &c : r.coefs) c = k * c;
        return r;
    }
};

struct AffExpr {
    LinExpr linexpr;
    double constant;
};
struct Inequality {
    LinExpr linexpr;
    double b;
};
struct Solution {
    double objective_value;
    std::unordered_map<std::size_t, double> values;
    int32_t pivots;
    uint64_t trace_hash;
};

// Flatten (objective, constraints) into the dz_model arrays: a variable table
// keyed by Variable.id in first-appearance order, terms in their given order.
struct Flat {
    std::vector<std::size_t> ids;
    std::unordered_map<std::size_t, int32_t> index;
    std::vector<uint8_t> has_lb, has_ub;
    std::vector<double> lb, ub;
    std::vector<int32_t> obj_var, row_var;
    std::vector<double> obj_coef, row_coef, rhs;
    std::vector<int64_t> row_ptr{0};

    int32_t var(const Variable &v) {
        auto it = index.find(v.id);
        if (it != index.end()) return it->second; // bounds of the first occurrence (simplex.rs:134)
        const int32_t k = (int32_t)ids.size();
        index.emplace(v.id, k);
        ids.push_back(v.id);
        has_lb.push_back(v.lb.has_value());
        has_ub.push_back(v.ub.has_value());
        lb.push_back(v.lb.value_or(0.0));
        ub.push_back(v.ub.value_or(0.0));
        return k;
    }
};

Flat flatten(const AffExpr &objective, const std::vector<Inequality> &constraints) {
    // A PyLinExpr whose coefs and vars differ in length is malformed.  The reference zips
    // them (model.rs:14) after registering every listed variable (simplex.rs:126-151), which
    // leaves bound rows of variables that have no term; rather than guess at that, the
    // drop-in refuses the model.
    auto well_formed = [](const LinExpr &e) { return e.coefs.size() == e.vars.size(); };
    bool ok = well_formed(objective.linexpr);
    for (const auto &c : constraints) ok = ok && well_formed(c.linexpr);
    if (!ok) throw py::value_error("PyLinExpr: coefs and vars differ in length");
    Flat f;
    const std::size_t n_obj = std::min(objective.linexpr.coefs.size(), objective.linexpr.vars.size());
    for (std::size_t i = 0; i < n_obj; ++i) {
        f.obj_var.push_back(f.var(objective.linexpr.vars[i]));
        f.obj_coef.push_back(objective.linexpr.coefs[i]);
    }
    for (const auto &c : constraints) {
        const std::size_t n = std::min(c.linexpr.coefs.size(), c.linexpr.vars.size());
        for (std::size_t i = 0; i < n; ++i) {
            f.row_var.push_back(f.var(c.linexpr.vars[i]));
            f.row_coef.push_back(c.linexpr.coefs[i]);
        }
        f.row_ptr.push_back((int64_t)f.row_var.size());
        f.rhs.push_back(c.b);
    }
    return f;
}

dz_model model_of(const Flat &f, double obj_const) {
    dz_model m{};
    m.n_vars = (int32_t)f.ids.size();
    m.has_lb = f.has_lb.data();
    m.has_ub = f.has_ub.data();
    m.lb = f.lb.data();
    m.ub = f.ub.data();
    m.n_obj = (int32_t)f.obj_var.size();
    m.obj_var = f.obj_var.data();
    m.obj_coef = f.obj_coef.data();
    m.obj_const = obj_const;
    m.n_rows = (int32_t)f.rhs.size();
    m.row_ptr = f.row_ptr.data();
    m.row_var = f.row_var.data();
    m.row_coef = f.row_coef.data();
    m.rhs = f.rhs.data();
    return m;
}

// What the lowering depends on besides the numbers: the variable table's bound
// flags and the index pattern of the objective and of every row.  LPs with equal
// keys share one dz_template.
std::string structure_key(const Flat &f) {
    std::string k;
    auto put = [&k](const void *p, std::size_t n) { k.append(static_cast<const char *>(p), n); };
    const int64_t head[3] = {(int64_t)f.ids.size(), (int64_t)f.obj_var.size(), (int64_t)f.rhs.size()};
    put(head, sizeof(head));
    put(f.has_lb.data(), f.has_lb.size());
    put(f.has_ub.data(), f.has_ub.size());
    put(f.obj_var.data(), f.obj_var.size() * sizeof(int32_t));
    put(f.row_ptr.data(), f.row_ptr.size() * sizeof(int64_t));
    put(f.row_var.data(), f.row_var.size() * sizeof(int32_t));
    return k;
}

Solution solve(const AffExpr &objective, const std::vector<Inequality> &constraints) {
    const Flat f = flatten(objective, constraints);
    dz_model m = model_of(f, objective.constant);

    dz_options opt;
    dz_options_default(&opt);
    dz_solution sol{};
    std::vector<double> values(f.ids.size() + 1, 0.0);
    int rc;
    {
        py::gil_scoped_release release; // the reference holds the GIL; we need not
        rc = dz_solve_model(&m, &opt, &sol, values.data());
    }
    if (rc != DZ_OK)
        throw std::runtime_error(std::string("dantzig_b200: ") + dz_last_error());
    switch (sol.status) {
    case DZ_OPTIMAL: break;
    case DZ_UNBOUNDED: { // lib.rs:24
        py::object exc = py::module_::import("dantzig.exceptions").attr("UnboundedError");
        PyErr_SetString(exc.ptr(), "The objective is unbounded");
        throw py::error_already_set();
    }
    case DZ_INFEASIBLE: { // lib.rs:25
        py::object exc = py::module_::import("dantzig.exceptions").attr("InfeasibleError");
        PyErr_SetString(exc.ptr(), "The model is infeasible");
        throw py::error_already_set();
    }
    case DZ_BREAKDOWN: {
        py::object exc = py::module_::import("dantzig.rust").attr("PanicException");
        PyErr_SetString(exc.ptr(), "numerical breakdown: the reference solver panics on this model "
                                   "(safe_divide / unexpected code path)");
        throw py::error_already_set();
    }
    default:
        throw std::runtime_error("dantzig_b200: pivot watchdog reached (the reference would not terminate)");
    }
    Solution s;
    s.objective_value = sol.objective;
    s.pivots = sol.pivots;
    s.trace_hash = sol.trace_hash;
    for (std::size_t k = 0; k < f.ids.size(); ++k) s.values[f.ids[k]] = values[k];
    return s;
}

// Batched entry (no reference analogue): LP i is
// (objectives[i], constraints[i]) with solve()'s meaning.  LPs are grouped by
// structure, each group goes through dz_solve_batch as one launch, and the
// result list holds, per LP and in input order, a PySolution or the exception
// instance solve() would have raised (returned, not raised).
py::list solve_batch(const std::vector<AffExpr> &objectives,
                     const std::vector<std::vector<Inequality>> &constraints, int n_gpus,
                     const std::string &numerics) {
    // numerics: "exact" (default, the reference's floating-point order) or "fast" (opt-in,
    // DZ_NUMERICS_FAST: agrees with the reference to rounding only)
    if (numerics != "exact" && numerics != "fast")
        throw py::value_error("solve_batch: numerics must be \"exact\" or \"fast\"");
    if (objectives.size() != constraints.size())
        throw py::value_error("solve_batch: objectives and constraints differ in length");
    const std::size_t n = objectives.size();
    std::vector<Flat> flats;
    flats.reserve(n);
    std::vector<std::string> keys;                                   // first-seen order
    std::unordered_map<std::string, std::vector<std::size_t>> groups;
    for (std::size_t i = 0; i < n; ++i) {
        flats.push_back(flatten(objectives[i], constraints[i]));
        std::string k = structure_key(flats.back());
        auto it = groups.find(k);
        if (it == groups.end()) {
            keys.push_back(k);
            groups.emplace(std::move(k), std::vector<std::size_t>{i});
        } else {
            it->second.push_back(i);
        }
    }
    std::vector<int32_t> status(n, 0), pivots(n, 0);
    std::vector<uint64_t> hash(n, 0);
    std::vector<double> objective(n, 0.0);
    std::vector<std::vector<double>> values(n);
    std::string failure;
    {
        py::gil_scoped_release release;
        dz_options opt;
        dz_options_default(&opt);
        opt.numerics = numerics == "fast" ? DZ_NUMERICS_FAST : DZ_NUMERICS_EXACT;
        for (const auto &k : keys) {
            const auto &members = groups[k];
            const std::size_t B = members.size();
            dz_template *t = nullptr;
            dz_model m0 = model_of(flats[members[0]], objectives[members[0]].constant);
            if (dz_template_create(&m0, &t) != DZ_OK) {
                failure = dz_last_error();
                break;
            }
            dz_template_info info{};
            dz_template_get_info(t, &info);
            std::vector<int32_t> orig((std::size_t)std::max(info.n_orig, 1));
            dz_template_get_arrays(t, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   orig.data(), nullptr, nullptr);
            std::vector<double> theta(B * (std::size_t)info.n_theta);
            int rc = DZ_OK;
            for (std::size_t j = 0; j < B && rc == DZ_OK; ++j) {
                dz_model mj = model_of(flats[members[j]], objectives[members[j]].constant);
                rc = dz_template_pack_theta(t, &mj, theta.data() + j * (std::size_t)info.n_theta);
            }
            std::vector<int32_t> st(B), pv(B), np(B);
            std::vector<uint64_t> th(B);
            std::vector<double> ob(B), vals(std::max<std::size_t>(B * (std::size_t)info.n_orig, 1));
            if (rc == DZ_OK) {
                dz_batch_result r{};
                r.status = st.data();
                r.pivots = pv.data();
                r.n_primal = np.data();
                r.trace_hash = th.data();
                r.objective = ob.data();
                r.values = vals.data();
                // n_gpus == 1: this device; otherwise the group is sharded over n_gpus devices
                // (0 = all visible), contiguous ranges, no collective (dz_solve_batch_multi)
                rc = n_gpus == 1 ? dz_solve_batch(t, (int64_t)B, theta.data(), &opt, &r)
                                 : dz_solve_batch_multi(t, (int64_t)B, theta.data(), n_gpus, &opt, &r);
            }
            if (rc != DZ_OK) failure = dz_last_error();
            dz_template_destroy(t);
            if (rc != DZ_OK) break;
            for (std::size_t j = 0; j < B; ++j) {
                const std::size_t i = members[j];
                status[i] = st[j];
                pivots[i] = pv[j];
                hash[i] = th[j];
                objective[i] = ob[j];
                values[i].assign(flats[i].ids.size(), 0.0);
                for (int32_t o = 0; o < info.n_orig; ++o)
                    values[i][(std::size_t)orig[(std::size_t)o]] = vals[j * (std::size_t)info.n_orig + (std::size_t)o];
            }
        }
    }
    if (!failure.empty()) throw std::runtime_error("dantzig_b200: " + failure);
    py::object exceptions = py::module_::import("dantzig.exceptions");
    py::object panic = py::module_::import("dantzig.rust").attr("PanicException");
    py::list out;
    for (std::size_t i = 0; i < n; ++i) {
        switch (status[i]) {
        case DZ_OPTIMAL: {
            Solution s;
            s.objective_value = objective[i];
            s.pivots = pivots[i];
            s.trace_hash = hash[i];
            for (std::size_t k = 0; k < flats[i].ids.size(); ++k) s.values[flats[i].ids[k]] = values[i][k];
            out.append(py::cast(std::move(s)));
            break;
        }
        case DZ_UNBOUNDED:
            out.append(exceptions.attr("UnboundedError")("The objective is unbounded"));
            break;
        case DZ_INFEASIBLE:
            out.append(exceptions.attr("InfeasibleError")("The model is infeasible"));
            break;
        case DZ_BREAKDOWN:
            out.append(panic("numerical breakdown: the reference solver panics on this model "
                             "(safe_divide / unexpected code path)"));
            break;
        default:
            out.append(py::module_::import("builtins").attr("RuntimeError")(
                "dantzig_b200: pivot watchdog reached (the reference would not terminate)"));
        }
    }
    return out;
}

} // namespace

PYBIND11_MODULE(rust, m) {
    m.attr("__name__") = "dantzig.rust"; // pyclass(module = "dantzig.rust")
    m.doc() = "B200-native drop-in for dantzig's compiled solver module";

    // PanicException derives from BaseException like pyo3_runtime.PanicException
    m.attr("PanicException") = py::reinterpret_steal<py::object>(
        PyErr_NewException("dantzig.rust.PanicException", PyExc_BaseException, nullptr));

    py::class_<Variable>(m, "Variable")
        .def(py::init<std::optional<double>, std::optional<double>>(), py::kw_only(), py::arg("lb"),
             py::arg("ub"))
        .def_readonly("id", &Variable::id)
        .def_readonly("lb", &Variable::lb)
        .def_readonly("ub", &Variable::ub);

    py::class_<LinExpr>(m, "PyLinExpr")
        .def(py::init<std::vector<double>, std::vector<Variable>>(), py::arg("coefs"), py::arg("vars"))
        .def("map_ids_to_coefs", &LinExpr::map_ids_to_coefs)
        .def("__neg__", &LinExpr::neg)
        .def("__add__", &LinExpr::add)
        .def("__mul__", &LinExpr::mul);

    py::class_<AffExpr>(m, "PyAffExpr")
        .def(py::init([](const LinExpr &l, double c) { return AffExpr{l, c}; }), py::kw_only(),
             py::arg("linexpr"), py::arg("constant"))
        .def_property_readonly("pylinexpr", [](const AffExpr &a) { return a.linexpr; })
        .def_readonly("constant", &AffExpr::constant);

    py::class_<Inequality>(m, "PyInequality")
        .def(py::init([](const LinExpr &l, double b) { return Inequality{l, b}; }), py::kw_only(),
             py::arg("linexpr"), py::arg("b"));

    py::class_<Solution>(m, "PySolution")
        .def_readonly("objective_value", &Solution::objective_value)
        .def_readonly("pivots", &Solution::pivots)         // extension: not in the reference
        .def_readonly("trace_hash", &Solution::trace_hash) // extension
        .def("__getitem__", [](const Solution &s, const Variable &v) {
            auto it = s.values.find(v.id);
            return it == s.values.end() ? 0.0 : it->second; // pyobjs.rs:163-165
        });

    m.def("solve", &solve, py::arg("objective"), py::arg("constraints"));
    m.def("solve_batch", &solve_batch, py::arg("objectives"), py::arg("constraints"), py::arg("n_gpus") = 1,
          py::arg("numerics") = "exact");
}
