// harness_common.cpp -- C entry points of oracle/_ref/libbxref.so (test
// infrastructure; see bxref.h).  Solver-independent part: problem loading via
// the reference's own parser / merged constraints, dispatch on
// problem_solver_type like lib/src/itm.hpp:94-203, and the libstdc++
// primitives the path depends on.
#include "harness.hpp"

#include "branch-and-bound-solver.hpp"
#include "exhaustive-solver.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>
#include <random>

using namespace baryonyx;

namespace {

solver_parameters
to_reference(const bxref_params& p)
{
    solver_parameters r;
    r.time_limit = p.time_limit;
    r.theta = p.theta;
    r.delta = p.delta;
    r.kappa_min = p.kappa_min;
    r.kappa_step = p.kappa_step;
    r.kappa_max = p.kappa_max;
    r.alpha = p.alpha;
    r.pushing_k_factor = p.pushing_k_factor;
    r.pushing_objective_amplifier = p.pushing_objective_amplifier;
    r.init_policy_random = p.init_policy_random;
    r.init_crossover_bastert_insertion = p.init_crossover_bastert_insertion;
    r.init_crossover_solution_selection_mean =
      p.init_crossover_solution_selection_mean;
    r.init_crossover_solution_selection_stddev =
      p.init_crossover_solution_selection_stddev;
    r.init_mutation_variable_mean = p.init_mutation_variable_mean;
    r.init_mutation_variable_stddev = p.init_mutation_variable_stddev;
    r.init_mutation_value_mean = p.init_mutation_value_mean;
    r.init_mutation_value_stddev = p.init_mutation_value_stddev;
    r.init_kappa_improve_start = p.init_kappa_improve_start;
    r.init_kappa_improve_increase = p.init_kappa_improve_increase;
    r.init_kappa_improve_stop = p.init_kappa_improve_stop;
    r.w = p.w;
    r.seed = p.seed;
    r.thread = p.thread;
    r.limit = static_cast<long int>(p.limit);
    r.print_level = p.print_level;
    r.pushes_limit = p.pushes_limit;
    r.pushing_iteration_limit = p.pushing_iteration_limit;
    r.init_population_size = p.init_population_size;
    r.pre_order =
      static_cast<solver_parameters::pre_constraint_order>(p.pre_order);
    r.order = static_cast<solver_parameters::constraint_order>(p.order);
    r.float_type =
      static_cast<solver_parameters::floating_point_type>(p.float_type);
    r.init_policy =
      static_cast<solver_parameters::init_policy_type>(p.init_policy);
    r.cost_norm = static_cast<solver_parameters::cost_norm_type>(p.cost_norm);
    r.mode = static_cast<solver_parameters::mode_type>(p.mode);
    r.preprocessor =
      static_cast<solver_parameters::preprocessor_options>(p.preprocessor);
    r.observer = static_cast<solver_parameters::observer_type>(p.observer);
    r.storage = static_cast<solver_parameters::storage_type>(p.storage);
    r.solver = static_cast<solver_parameters::solver_type>(p.solver);
    r.debug = p.debug != 0;
    return r;
}

void
from_reference(const solver_parameters& r, bxref_params& p)
{
    p.time_limit = r.time_limit;
    p.theta = r.theta;
    p.delta = r.delta;
    p.kappa_min = r.kappa_min;
    p.kappa_step = r.kappa_step;
    p.kappa_max = r.kappa_max;
    p.alpha = r.alpha;
    p.pushing_k_factor = r.pushing_k_factor;
    p.pushing_objective_amplifier = r.pushing_objective_amplifier;
    p.init_policy_random = r.init_policy_random;
    p.init_crossover_bastert_insertion = r.init_crossover_bastert_insertion;
    p.init_crossover_solution_selection_mean =
      r.init_crossover_solution_selection_mean;
    p.init_crossover_solution_selection_stddev =
      r.init_crossover_solution_selection_stddev;
    p.init_mutation_variable_mean = r.init_mutation_variable_mean;
    p.init_mutation_variable_stddev = r.init_mutation_variable_stddev;
    p.init_mutation_value_mean = r.init_mutation_value_mean;
    p.init_mutation_value_stddev = r.init_mutation_value_stddev;
    p.init_kappa_improve_start = r.init_kappa_improve_start;
    p.init_kappa_improve_increase = r.init_kappa_improve_increase;
    p.init_kappa_improve_stop = r.init_kappa_improve_stop;
    p.w = r.w;
    p.seed = r.seed;
    p.thread = r.thread;
    p.limit = r.limit;
    p.print_level = r.print_level;
    p.pushes_limit = r.pushes_limit;
    p.pushing_iteration_limit = r.pushing_iteration_limit;
    p.init_population_size = r.init_population_size;
    p.pre_order = static_cast<int>(r.pre_order);
    p.order = static_cast<int>(r.order);
    p.float_type = static_cast<int>(r.float_type);
    p.init_policy = static_cast<int>(r.init_policy);
    p.cost_norm = static_cast<int>(r.cost_norm);
    p.mode = static_cast<int>(r.mode);
    p.preprocessor = static_cast<int>(r.preprocessor);
    p.observer = static_cast<int>(r.observer);
    p.storage = static_cast<int>(r.storage);
    p.solver = static_cast<int>(r.solver);
    p.debug = r.debug ? 1 : 0;
}

struct rc_item
{
    double value;
    int id;
    int f;
};

struct rc_view
{
    rc_item* p;
    rc_item& operator[](int i) const
    {
        return p[i];
    }
};

struct fe
{
    int factor;
};

template<typename Mode>
int
bb_solve(double* value, int* factor, int* id, int n, int bmin, int bmax)
{
    std::vector<rc_item> R(n);
    for (int i = 0; i != n; ++i)
        R[i] = { value[i], id[i], factor[i] };

    itm::branch_and_bound_solver<Mode, double> bb;
    bb.reserve(n);
    rc_view v{ R.data() };
    int ret = bb.solve(v, n, bmin, bmax);

    for (int i = 0; i != n; ++i) {
        value[i] = R[i].value;
        id[i] = R[i].id;
        factor[i] = R[i].f;
    }
    return ret;
}

template<typename Mode>
int
ex_solve(const int* cst_factor,
         double* value,
         int* factor,
         int* id,
         int n,
         int bmin,
         int bmax)
{
    std::vector<fe> elems(n);
    for (int i = 0; i != n; ++i)
        elems[i].factor = cst_factor[i];

    std::vector<rc_item> R(n);
    for (int i = 0; i != n; ++i)
        R[i] = { value[i], id[i], factor[i] };

    itm::exhaustive_solver<Mode, double> ex;
    ex.reserve(n, 1);
    ex.build_constraints(0, elems, bmin, bmax);
    rc_view v{ R.data() };
    int ret = ex.solve(0, v, n);

    for (int i = 0; i != n; ++i) {
        value[i] = R[i].value;
        id[i] = R[i].id;
        factor[i] = R[i].f;
    }
    return ret;
}

} // namespace

extern "C" {

void
bxref_default_params(bxref_params* p)
{
    from_reference(solver_parameters{}, *p);
}

void*
bxref_open(const char* lp_path, const bxref_params* p)
{
    auto h = std::make_unique<bxref::handle>();

    {
        // context_set_solver_parameters (lib/src/private.cpp:51-172) takes a
        // context_ptr; wrap a temporary one to apply the reference's clamps.
        auto ptr = make_context(3);
        h->orig = to_reference(*p);
        context_set_solver_parameters(ptr, h->orig);
        h->ctx.parameters = context_get_solver_parameters(ptr);
    }

    std::ifstream ifs(lp_path);
    if (!ifs.is_open())
        return nullptr;

    h->raw = make_problem(h->ctx, ifs);
    if (!h->raw)
        return nullptr;

    h->pb = unpreprocess(h->ctx, h->raw);
    h->csts = itm::make_merged_constraints(h->ctx, h->pb);
    h->m = length(h->csts);
    h->n = length(h->pb.vars.values);
    h->nnz = 0;
    for (const auto& c : h->csts)
        h->nnz += length(c.elements);

    return h.release();
}

// The handle behind the reference's preprocessor: baryonyx::preprocess
// (lib/src/preprocessor.cpp:893-903) on the parse, then baryonyx::affect
// (:876-891) for every (variable, value) given, one after the other, then
// make_merged_constraints.  A contract failure of the reference (bx_reach,
// bx_assert: std::logic_error in this build) comes back as a null handle with
// the message in err.  The reference writes "preprocessed.lp" into the working
// directory as a side effect: callers chdir to a scratch directory first.
void*
bxref_open_pp(const char* lp_path,
              const bxref_params* p,
              int n_affect,
              const int* variable,
              const int* value,
              char* err,
              int err_len)
{
    auto note = [&](const char* what) {
        if (err && err_len > 0) {
            std::strncpy(err, what, static_cast<size_t>(err_len) - 1);
            err[err_len - 1] = 0;
        }
    };
    note("");
    try {
        auto h = std::make_unique<bxref::handle>();
        {
            auto ptr = make_context(3);
            h->orig = to_reference(*p);
            context_set_solver_parameters(ptr, h->orig);
            h->ctx.parameters = context_get_solver_parameters(ptr);
        }
        std::ifstream ifs(lp_path);
        if (!ifs.is_open()) {
            note("file_not_found");
            return nullptr;
        }
        h->raw = make_problem(h->ctx, ifs);
        if (!h->raw) {
            note("parse");
            return nullptr;
        }
        h->pb = preprocess(h->ctx, h->raw);
        for (int i = 0; i != n_affect; ++i) {
            if (variable[i] < 0 || variable[i] >= length(h->pb.vars.values)) {
                note("variable index out of range");
                return nullptr;
            }
            h->pb = affect(h->ctx, h->pb, variable[i], value[i] != 0);
        }
        h->csts = itm::make_merged_constraints(h->ctx, h->pb);
        h->m = length(h->csts);
        h->n = length(h->pb.vars.values);
        h->nnz = 0;
        for (const auto& c : h->csts)
            h->nnz += length(c.elements);
        return h.release();
    } catch (const std::exception& e) {
        note(e.what());
        return nullptr;
    }
}

int
bxref_affected_count(void* hv)
{
    return length(static_cast<bxref::handle*>(hv)->pb.affected_vars.names);
}

const char*
bxref_affected(void* hv, int i, int* value)
{
    auto* h = static_cast<bxref::handle*>(hv);
    if (i < 0 || i >= length(h->pb.affected_vars.names))
        return nullptr;
    thread_local std::string name;
    name.assign(h->pb.affected_vars.names[i].data(), h->pb.affected_vars.names[i].size());
    *value = h->pb.affected_vars.values[i] ? 1 : 0;
    return name.c_str();
}

// The same handle from arrays: baryonyx::raw_problem filled directly, for
// instances whose LP text the reference's parser would take minutes to read
// (its objective look-up is quadratic: parser.cpp:1085-1097; 519 s for the
// rail4284-shaped file).  Rows: cmin == cmax -> "=", a finite side -> ">=" /
// "<=", both -> two constraints over the same elements, which
// make_merged_constraints merges again.  Variables are binary, named x0..x{n-1}.
void*
bxref_open_arrays(int m,
                  int n,
                  const int* row_ptr,
                  const int* col,
                  const int* factor,
                  const int* cmin,
                  const int* cmax,
                  int maximize,
                  int n_obj,
                  const int* obj_var,
                  const double* obj_factor,
                  double cost_constant,
                  const bxref_params* p)
{
    auto h = std::make_unique<bxref::handle>();
    {
        auto ptr = make_context(3);
        h->orig = to_reference(*p);
        context_set_solver_parameters(ptr, h->orig);
        h->ctx.parameters = context_get_solver_parameters(ptr);
    }

    raw_problem& raw = h->raw;
    raw.strings = std::make_shared<string_buffer>();
    raw.type = maximize ? objective_function_type::maximize : objective_function_type::minimize;
    raw.vars.names.reserve(n);
    raw.vars.values.reserve(n);
    for (int i = 0; i != n; ++i) {
        raw.vars.names.emplace_back(raw.strings->append("x" + std::to_string(i)));
        raw.vars.values.emplace_back(0, 1, variable_type::binary);
    }
    raw.objective.value = cost_constant;
    for (int i = 0; i != n_obj; ++i)
        raw.objective.elements.emplace_back(obj_factor[i], obj_var[i]);

    int id = 0;
    for (int k = 0; k != m; ++k) {
        auto elements = [&] {
            std::vector<function_element> e;
            e.reserve(row_ptr[k + 1] - row_ptr[k]);
            for (int t = row_ptr[k]; t != row_ptr[k + 1]; ++t)
                e.emplace_back(factor[t], col[t]);
            return e;
        };
        const auto label = raw.strings->append("c" + std::to_string(k));
        const bool lo = cmin[k] != std::numeric_limits<int>::min();
        const bool hi = cmax[k] != std::numeric_limits<int>::max();
        if (lo && hi && cmin[k] == cmax[k])
            raw.equal_constraints.emplace_back(label, elements(), cmin[k], id++);
        else {
            if (lo)
                raw.greater_constraints.emplace_back(label, elements(), cmin[k], id++);
            if (hi)
                raw.less_constraints.emplace_back(label, elements(), cmax[k], id++);
        }
    }

    h->pb = unpreprocess(h->ctx, h->raw);
    h->csts = itm::make_merged_constraints(h->ctx, h->pb);
    h->m = length(h->csts);
    h->n = length(h->pb.vars.values);
    h->nnz = 0;
    for (const auto& c : h->csts)
        h->nnz += length(c.elements);
    return h.release();
}

void
bxref_close(void* h)
{
    delete static_cast<bxref::handle*>(h);
}

int
bxref_info(void* hv,
           int* m,
           int* n,
           int* nnz,
           int* problem_type,
           int* maximize,
           double* cost_constant,
           int* n_obj)
{
    auto* h = static_cast<bxref::handle*>(hv);
    *m = h->m;
    *n = h->n;
    *nnz = h->nnz;
    *problem_type = static_cast<int>(h->pb.problem_type);
    *maximize = h->pb.type == objective_function_type::maximize ? 1 : 0;
    *cost_constant = h->pb.objective.value;
    *n_obj = length(h->pb.objective.elements);
    return 0;
}

int
bxref_get_params(void* hv, bxref_params* out)
{
    auto* h = static_cast<bxref::handle*>(hv);
    from_reference(h->ctx.parameters, *out);
    return 0;
}

int
bxref_export_rows(void* hv,
                  int* row_ptr,
                  int* col,
                  int* factor,
                  int* cmin,
                  int* cmax)
{
    auto* h = static_cast<bxref::handle*>(hv);
    int e = 0;
    row_ptr[0] = 0;
    for (int k = 0; k != h->m; ++k) {
        for (const auto& el : h->csts[k].elements) {
            col[e] = el.variable_index;
            factor[e] = el.factor;
            ++e;
        }
        row_ptr[k + 1] = e;
        cmin[k] = h->csts[k].min;
        cmax[k] = h->csts[k].max;
    }
    return 0;
}

int
bxref_export_objective(void* hv, int* var, double* factor)
{
    auto* h = static_cast<bxref::handle*>(hv);
    int i = 0;
    for (const auto& el : h->pb.objective.elements) {
        var[i] = el.variable_index;
        factor[i] = el.factor;
        ++i;
    }
    return 0;
}

const char*
bxref_variable(void* hv, int i, int* min, int* max, int* type)
{
    auto* h = static_cast<bxref::handle*>(hv);
    if (i < 0 || i >= length(h->pb.vars.values))
        return nullptr;
    thread_local std::string name;
    name.assign(h->pb.vars.names[i].data(), h->pb.vars.names[i].size());
    *min = h->pb.vars.values[i].min;
    *max = h->pb.vars.values[i].max;
    *type = static_cast<int>(h->pb.vars.values[i].type);
    return name.c_str();
}

#define BXREF_DISPATCH(h, CALL)                                               \
    switch ((h)->pb.problem_type) {                                           \
    case problem_solver_type::equalities_01:                                  \
        CALL(eq01);                                                           \
        break;                                                                \
    case problem_solver_type::equalities_101:                                 \
        CALL(eq101);                                                          \
        break;                                                                \
    case problem_solver_type::inequalities_01:                                \
        CALL(ineq01);                                                         \
        break;                                                                \
    case problem_solver_type::inequalities_101:                               \
        CALL(ineq101);                                                        \
        break;                                                                \
    case problem_solver_type::equalities_Z:                                   \
    case problem_solver_type::inequalities_Z:                                 \
        CALL(ineqZ);                                                          \
        break;                                                                \
    }

void*
bxref_run_begin(void* hv, uint32_t seed)
{
    auto* h = static_cast<bxref::handle*>(hv);
    bxref::run_base* run = nullptr;
#define CALL(tag) run = bxref::make_run_##tag(*h, seed)
    BXREF_DISPATCH(h, CALL)
#undef CALL
    return run;
}

void
bxref_run_end(void* run)
{
    delete static_cast<bxref::run_base*>(run);
}

int
bxref_run_info(void* run, double* delta, double* theta)
{
    static_cast<bxref::run_base*>(run)->info(delta, theta);
    return 0;
}

double
bxref_run_tail_value(void* run)
{
    return static_cast<bxref::run_base*>(run)->tail_value();
}

int
bxref_run_sweep(void* run, int push, double kappa, double obj_amp)
{
    return static_cast<bxref::run_base*>(run)->sweep(push, kappa, obj_amp);
}

int
bxref_run_state(void* run,
                uint64_t* x_words,
                double* pi,
                double* P,
                uint32_t* rng_state,
                int* order_list,
                int* order_len,
                double* norm_cost)
{
    static_cast<bxref::run_base*>(run)->state(
      x_words, pi, P, rng_state, order_list, order_len, norm_cost);
    return 0;
}

int
bxref_run_set_x(void* run, const uint64_t* x_words)
{
    static_cast<bxref::run_base*>(run)->set_x(x_words);
    return 0;
}

int
bxref_run_set_rng(void* run, uint32_t state)
{
    static_cast<bxref::run_base*>(run)->set_rng(state);
    return 0;
}

double
bxref_run_objective(void* run)
{
    return static_cast<bxref::run_base*>(run)->objective();
}

int
bxref_solve(void* hv,
            uint32_t seed,
            int8_t* x,
            double* value,
            int* remaining,
            int64_t* updates,
            int64_t* sweeps)
{
    auto* h = static_cast<bxref::handle*>(hv);
    bxref::solve_out out;
#define CALL(tag) bxref::solve_##tag(*h, seed, out)
    BXREF_DISPATCH(h, CALL)
#undef CALL
    if (x)
        std::copy(out.x.begin(), out.x.end(), x);
    if (value)
        *value = out.value;
    if (remaining)
        *remaining = out.remaining;
    if (updates)
        *updates = out.updates;
    if (sweeps)
        *sweeps = out.sweeps;
    return out.status;
}

int
bxref_solve_public(void* hv, uint32_t seed, int8_t* x, double* value)
{
    auto* h = static_cast<bxref::handle*>(hv);

    auto ptr = make_context(3);
    auto params = h->orig;
    params.seed = seed;
    params.preprocessor = solver_parameters::preprocessor_options::none;
    context_set_solver_parameters(ptr, params);
    result r = solve(ptr, h->raw);

    if (!r.solutions.empty()) {
        const auto& s = r.solutions.front();
        if (value)
            *value = s.value;
        if (x)
            for (int i = 0, e = length(s.variables); i != e; ++i)
                x[i] = static_cast<int8_t>(s.variables[i]);
    }
    return static_cast<int>(r.status);
}

int
bxref_optimize(void* hv,
               uint32_t seed,
               int threads,
               double time_limit,
               int8_t* x,
               double* value,
               int* remaining,
               int64_t* updates,
               int64_t* sweeps,
               double* seconds)
{
    auto* h = static_cast<bxref::handle*>(hv);
    bxref::solve_out out;
#define CALL(tag) bxref::optimize_##tag(*h, seed, threads, time_limit, out)
    BXREF_DISPATCH(h, CALL)
#undef CALL
    if (x)
        std::copy(out.x.begin(), out.x.end(), x);
    if (value)
        *value = out.value;
    if (remaining)
        *remaining = out.remaining;
    if (updates)
        *updates = out.updates;
    if (sweeps)
        *sweeps = out.sweeps;
    if (seconds)
        *seconds = out.seconds;
    return out.status;
}

void
bxref_optimize_report(double* first_feasible_seconds, int64_t* restarts)
{
    // set by the per-solver translation units; read through any of them
    bxref::optimize_report(first_feasible_seconds, restarts);
}

void*
bxref_opt_trace(void* hv, uint32_t seed, int rounds)
{
    auto* h = static_cast<bxref::handle*>(hv);
    auto t = std::make_unique<bxref::trace_data>();
#define CALL(tag) bxref::trace_##tag(*h, seed, rounds, *t)
    BXREF_DISPATCH(h, CALL)
#undef CALL
    t->call_number = nullptr;
    t->stop = nullptr;
    t->population = nullptr;
    return t.release();
}

void
bxref_opt_trace_free(void* t)
{
    delete static_cast<bxref::trace_data*>(t);
}

int
bxref_opt_trace_rounds(void* tv)
{
    return static_cast<int>(static_cast<bxref::trace_data*>(tv)->rounds.size());
}

int
bxref_opt_trace_round(void* tv, int r, uint64_t* x_words, uint32_t* rng_state, double* kappa,
                      double* pi, double* P)
{
    auto* t = static_cast<bxref::trace_data*>(tv);
    if (r < 0 || r >= static_cast<int>(t->rounds.size()))
        return -1;
    const auto& s = t->rounds[r];
    if (x_words)
        std::copy(s.x.begin(), s.x.end(), x_words);
    if (rng_state)
        *rng_state = s.rng;
    if (kappa)
        *kappa = s.kappa;
    if (pi)
        std::copy(s.pi.begin(), s.pi.end(), pi);
    if (P)
        std::copy(s.P.begin(), s.P.end(), P);
    return 0;
}

int
bxref_opt_trace_member(void* tv, int r, int i, uint64_t* x_words, double* value, int* remaining,
                       uint64_t* hash, int64_t* loop)
{
    auto* t = static_cast<bxref::trace_data*>(tv);
    if (r < 0 || r >= static_cast<int>(t->rounds.size()))
        return -1;
    const auto& ms = t->rounds[r].members;
    if (i < 0 || i >= static_cast<int>(ms.size()))
        return -1;
    const auto& m = ms[i];
    if (x_words)
        std::copy(m.x.begin(), m.x.end(), x_words);
    if (value)
        *value = m.value;
    if (remaining)
        *remaining = m.remaining;
    if (hash)
        *hash = m.hash;
    if (loop)
        *loop = m.loop;
    return 0;
}

int
bxref_opt_trace_sweeps(void* tv)
{
    return static_cast<int>(static_cast<bxref::trace_data*>(tv)->sweep_len.size());
}

int
bxref_opt_trace_sweep_log(void* tv, int* round, int* push, double* kappa, int* len)
{
    auto* t = static_cast<bxref::trace_data*>(tv);
    std::copy(t->sweep_round.begin(), t->sweep_round.end(), round);
    std::copy(t->sweep_push.begin(), t->sweep_push.end(), push);
    std::copy(t->sweep_kappa.begin(), t->sweep_kappa.end(), kappa);
    std::copy(t->sweep_len.begin(), t->sweep_len.end(), len);
    return 0;
}

// ---------------------------------------------------------------------------
// libstdc++ primitives
// ---------------------------------------------------------------------------

void
bxref_std_sort(double* value, int* id, int n, int descending)
{
    struct rc
    {
        double value;
        int id;
    };
    std::vector<rc> v(n);
    for (int i = 0; i != n; ++i)
        v[i] = { value[i], id[i] };

    if (descending)
        std::sort(v.begin(), v.end(), [](const auto& l, const auto& r) {
            return r.value < l.value;
        });
    else
        std::sort(v.begin(), v.end(), [](const auto& l, const auto& r) {
            return l.value < r.value;
        });

    for (int i = 0; i != n; ++i) {
        value[i] = v[i].value;
        id[i] = v[i].id;
    }
}

void
bxref_std_sort_pairs(int* first, int* second, int n, int descending)
{
    std::vector<std::pair<int, int>> v(n);
    for (int i = 0; i != n; ++i)
        v[i] = { first[i], second[i] };

    if (descending)
        std::sort(v.begin(), v.end(), [](const auto& l, const auto& r) {
            return r.second < l.second;
        });
    else
        std::sort(v.begin(), v.end(), [](const auto& l, const auto& r) {
            return l.second < r.second;
        });

    for (int i = 0; i != n; ++i) {
        first[i] = v[i].first;
        second[i] = v[i].second;
    }
}

void
bxref_std_shuffle(int* a, int n, uint32_t* rng_state)
{
    random_engine rng;
    bxref::rng_set_state(rng, *rng_state);
    std::shuffle(a, a + n, rng);
    *rng_state = bxref::rng_state_of(rng);
}

int
bxref_std_bernoulli(double p, uint32_t* rng_state)
{
    random_engine rng;
    bxref::rng_set_state(rng, *rng_state);
    std::bernoulli_distribution d(p);
    int r = d(rng) ? 1 : 0;
    *rng_state = bxref::rng_state_of(rng);
    return r;
}

uint64_t
bxref_std_uniform_u64(uint64_t a, uint64_t b, uint32_t* rng_state)
{
    random_engine rng;
    bxref::rng_set_state(rng, *rng_state);
    std::uniform_int_distribution<std::size_t> d(a, b);
    uint64_t r = d(rng);
    *rng_state = bxref::rng_state_of(rng);
    return r;
}

void
bxref_std_normal(double mean,
                 double stddev,
                 int count,
                 double* out,
                 uint32_t* rng_state)
{
    random_engine rng;
    bxref::rng_set_state(rng, *rng_state);
    std::normal_distribution<double> d(mean, stddev);
    for (int i = 0; i != count; ++i)
        out[i] = d(rng);
    *rng_state = bxref::rng_state_of(rng);
}

void
bxref_generate_seed(uint64_t master_seed, unsigned count, uint64_t* out)
{
    random_engine rng(static_cast<random_engine::result_type>(master_seed));
    auto seeds = itm::generate_seed(rng, count);
    for (unsigned i = 0; i != count; ++i)
        out[i] = seeds[i];
}

double
bxref_std_pow(double a, double b)
{
    return std::pow(a, b);
}

int
bxref_bb_solve(int maximize,
               double* value,
               int* factor,
               int* id,
               int n,
               int bmin,
               int bmax)
{
    return maximize
             ? bb_solve<itm::maximize_tag>(value, factor, id, n, bmin, bmax)
             : bb_solve<itm::minimize_tag>(value, factor, id, n, bmin, bmax);
}

int
bxref_exhaustive_solve(int maximize,
                       const int* cst_factor,
                       double* value,
                       int* factor,
                       int* id,
                       int n,
                       int bmin,
                       int bmax)
{
    return maximize ? ex_solve<itm::maximize_tag>(
                        cst_factor, value, factor, id, n, bmin, bmax)
                    : ex_solve<itm::minimize_tag>(
                        cst_factor, value, factor, id, n, bmin, bmax);
}

} // extern "C"
