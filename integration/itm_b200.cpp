// itm_b200.cpp -- the file a maintainer adds to the reference (lib/src/) to
// run baryonyx::solve and baryonyx::optimize on the B200 engine: it sees the reference's own types
// (context, problem, merged_constraint, result) on one side and only
// include/bxg.h on the other.  Compiled in this repository by oracle/Makefile
// (as part of the reference harness) so that tests/test_gpu_integration.py can
// check on the GPU box that the reference's API over the engine returns what
// the reference itself returns.  See INTEGRATION.md.
#include <baryonyx/core>

#include "itm-common.hpp"
#include "private.hpp"
#include "problem.hpp"

#include <bxg.h>

#include <chrono>
#include <vector>

namespace baryonyx {
namespace itm {

static bxg_params
to_pod(const solver_parameters& p)
{
    bxg_params q;
    bxg_default_params(&q);
    q.time_limit = p.time_limit;
    q.theta = p.theta;
    q.delta = p.delta;
    q.kappa_min = p.kappa_min;
    q.kappa_step = p.kappa_step;
    q.kappa_max = p.kappa_max;
    q.alpha = p.alpha;
    q.pushing_k_factor = p.pushing_k_factor;
    q.pushing_objective_amplifier = p.pushing_objective_amplifier;
    q.init_policy_random = p.init_policy_random;
    q.init_crossover_bastert_insertion = p.init_crossover_bastert_insertion;
    q.init_crossover_solution_selection_mean = p.init_crossover_solution_selection_mean;
    q.init_crossover_solution_selection_stddev = p.init_crossover_solution_selection_stddev;
    q.init_mutation_variable_mean = p.init_mutation_variable_mean;
    q.init_mutation_variable_stddev = p.init_mutation_variable_stddev;
    q.init_mutation_value_mean = p.init_mutation_value_mean;
    q.init_mutation_value_stddev = p.init_mutation_value_stddev;
    q.init_kappa_improve_start = p.init_kappa_improve_start;
    q.init_kappa_improve_increase = p.init_kappa_improve_increase;
    q.init_kappa_improve_stop = p.init_kappa_improve_stop;
    q.w = p.w;
    q.seed = p.seed;
    q.thread = p.thread;
    q.limit = p.limit;
    q.print_level = p.print_level;
    q.pushes_limit = p.pushes_limit;
    q.pushing_iteration_limit = p.pushing_iteration_limit;
    q.init_population_size = p.init_population_size;
    q.pre_order = static_cast<int>(p.pre_order);
    q.order = static_cast<int>(p.order);
    q.float_type = static_cast<int>(p.float_type);
    q.init_policy = static_cast<int>(p.init_policy);
    q.cost_norm = static_cast<int>(p.cost_norm);
    q.mode = static_cast<int>(p.mode);
    q.preprocessor = static_cast<int>(p.preprocessor);
    q.observer = static_cast<int>(p.observer);
    q.storage = static_cast<int>(p.storage);
    q.solver = static_cast<int>(p.solver);
    q.debug = p.debug ? 1 : 0;
    return q;
}

// The merged constraints and the objective as the arrays bxg_problem_create
// takes; failures of the engine become the reference's solver_failure.
static bxg_problem*
make_engine_problem(const problem& pb, const std::vector<merged_constraint>& csts, int device)
{
    std::vector<int32_t> row_ptr{ 0 }, col, fac, cmin, cmax;
    for (const auto& c : csts) {
        for (const auto& e : c.elements) {
            col.push_back(e.variable_index);
            fac.push_back(e.factor);
        }
        row_ptr.push_back(static_cast<int32_t>(col.size()));
        cmin.push_back(c.min);
        cmax.push_back(c.max);
    }
    std::vector<int32_t> ov;
    std::vector<double> of;
    for (const auto& e : pb.objective.elements) {
        ov.push_back(e.variable_index);
        of.push_back(e.factor);
    }

    bxg_problem* P = nullptr;
    const int rc = bxg_problem_create(length(csts), length(pb.vars.values), row_ptr.data(), col.data(),
                                      fac.data(), cmin.data(), cmax.data(),
                                      static_cast<int>(pb.problem_type),
                                      pb.type == objective_function_type::maximize,
                                      static_cast<int>(ov.size()), ov.data(), of.data(),
                                      pb.objective.value, device, &P);
    if (rc == BXG_E_UNREALISABLE)
        throw solver_failure(solver_error_tag::unrealisable_constraint);
    if (rc == BXG_E_MEMORY)
        throw solver_failure(solver_error_tag::not_enough_memory);
    if (rc != BXG_OK)
        throw solver_failure(solver_error_tag::no_solver_available);
    return P;
}

// What itm::solve (itm.hpp:94-145) calls instead of solve_equalities_01 & co.
// when a device is present: `restarts` lanes, lane 0 on the context's seed (so
// one lane returns what solve_problem<...> returns), the best lane wins.
// ctx.parameters are already clamped (context_set_solver_parameters), so w is
// a sweep count: the POD goes to the engine as it is.
result
solve_b200(const context& ctx, const problem& pb, int restarts, int device)
{
    result r;
    const auto variables = length(pb.vars.values);
    auto csts = make_merged_constraints(ctx, pb); // merged-constraint.cpp:297-343

    if (csts.empty() || pb.vars.values.empty()) {
        r.status = result_status::success;
        r.solutions.resize(1);
        r.solutions.back().value = pb.objective.value;
    } else {
        bxg_problem* P = make_engine_problem(pb, csts, device);
        int rc = BXG_OK;

        const bxg_params p = to_pod(ctx.parameters);
        bxg_batch* B = nullptr;
        if (bxg_batch_create(P, restarts, &p, &B) != BXG_OK) {
            bxg_problem_destroy(P);
            throw solver_failure(solver_error_tag::not_enough_memory);
        }

        std::vector<uint32_t> seeds(restarts); // init_random_generator_seed, itm-common.hpp:1442
        for (int l = 0; l != restarts; ++l)
            seeds[l] = static_cast<uint32_t>(init_random_generator_seed(ctx)) + static_cast<uint32_t>(l);
        std::vector<int32_t> status(restarts), remaining(restarts);
        std::vector<double> value(restarts);
        std::vector<int64_t> loop(restarts);
        rc = bxg_batch_begin(B, seeds.data(), 1);
        if (rc == BXG_OK)
            rc = bxg_batch_solve(B, status.data(), value.data(), remaining.data(), loop.data());
        int lane = 0, rem = 0;
        double best = 0;
        std::vector<int8_t> x(variables);
        if (rc == BXG_OK)
            rc = bxg_batch_argbest(B, &lane, &best, &rem);
        if (rc == BXG_OK)
            rc = bxg_batch_best(B, lane, x.data());
        bxg_batch_destroy(B);
        bxg_problem_destroy(P);
        if (rc != BXG_OK)
            throw solver_failure(solver_error_tag::no_solver_available);

        // solver_functor::operator()'s epilogue, itm-solver-common.hpp:216-224
        // (like the reference, which leaves result::loop and
        // result::remaining_constraints at their defaults there)
        r.status = static_cast<result_status>(status[lane]);
        if (rem != std::numeric_limits<int>::max()) { // !m_best.x.empty()
            r.solutions.resize(1);
            r.solutions[0].value = best;
            r.solutions[0].variables.assign(x.begin(), x.end());
        }
    }

    r.strings = pb.strings;
    r.variable_name = pb.vars.names;
    r.affected_vars = pb.affected_vars;
    r.variables = variables;
    r.constraints = length(csts);
    return r;
}


// What itm::optimize (itm.hpp:156-203) calls instead of optimize_equalities_01 &
// co.: optimize_problem<...> (itm-optimizer-common.hpp:778-908) over the engine.
// `workers` lanes play the reference's threads (0: as many as make the device
// busy); the population, the workers' seeds (generate_seed on the master engine),
// the once-a-second ctx.update report, the time limit and the shape of the result
// (storage_type one / bound / five) are the reference's.  max_rounds > 0 bounds
// the run by restarts per worker instead of by time (tests).
static void
member_to_solution(bxg_population* pop, int i, int variables, solution& s, int* remaining,
                   long int* loop)
{
    std::vector<uint64_t> words(static_cast<size_t>(variables) / 64 + 1);
    double value = 0;
    int rem = 0;
    int64_t lp = 0;
    if (bxg_population_member(pop, i, words.data(), &value, &rem, nullptr, &lp) != BXG_OK)
        throw solver_failure(solver_error_tag::no_solver_available);
    s.value = value;
    s.variables.resize(variables);
    for (int v = 0; v != variables; ++v) // convert(), itm-common.hpp (raw_result -> solution)
        s.variables[v] = static_cast<var_value>((words[v >> 6] >> (v & 63)) & 1u);
    if (remaining)
        *remaining = rem;
    if (loop)
        *loop = static_cast<long int>(lp);
}

result
optimize_b200(const context& ctx, const problem& pb, int workers, int device, long int max_rounds = 0)
{
    result r;

    if (ctx.start)
        ctx.start(ctx.parameters);

    const auto variables = length(pb.vars.values);
    auto csts = make_merged_constraints(ctx, pb);
    if (csts.empty() || pb.vars.values.empty()) {
        r.status = result_status::success;
        r.solutions.resize(1);
        r.solutions.back().value = pb.objective.value;
        r.strings = pb.strings;
        r.affected_vars = pb.affected_vars;
        r.variable_name = pb.vars.names;
        return r;
    }

    if (workers <= 0)
        workers = 1024;
    bxg_problem* P = make_engine_problem(pb, csts, device);
    const bxg_params p = to_pod(ctx.parameters);
    bxg_batch* B = nullptr;
    bxg_population* pop = nullptr;
    auto release = [&] {
        bxg_population_destroy(pop);
        bxg_batch_destroy(B);
        bxg_problem_destroy(P);
    };
    if (bxg_batch_create(P, workers, &p, &B) != BXG_OK ||
        bxg_population_create(P, &p, init_random_generator_seed(ctx), workers, &pop) != BXG_OK) {
        release();
        throw solver_failure(solver_error_tag::not_enough_memory);
    }
    std::vector<uint64_t> seeds(workers); // generate_seed(rng, thread), :817
    int rc = bxg_population_generate_seeds(pop, workers, seeds.data());
    for (int l = 0; rc == BXG_OK && l != workers; ++l)
        rc = bxg_population_worker_seed(pop, l, seeds[l]);

    const auto start = std::chrono::steady_clock::now();
    struct report
    {
        const context* ctx;
    } rep{ &ctx };
    auto update = [](int remaining, double value, int64_t loop, double duration, int64_t restarts,
                     void* user) {
        const context& c = *static_cast<report*>(user)->ctx;
        if (c.update)
            c.update(remaining, value, static_cast<long int>(loop), duration, static_cast<long int>(restarts));
    };
    if (rc == BXG_OK)
        rc = bxg_optimize(B, pop, nullptr, max_rounds > 0 ? 0.0 : ctx.parameters.time_limit, max_rounds,
                          0, update, &rep, nullptr);
    if (rc != BXG_OK) {
        release();
        throw solver_failure(solver_error_tag::no_solver_available);
    }

    r.strings = pb.strings;
    r.affected_vars = pb.affected_vars;
    r.variable_name = pb.vars.names;
    r.variables = variables;
    r.constraints = length(csts);

    // :881-902
    const int size = bxg_population_size(pop);
    auto fill = [&](int i, solution& s, int* rem, long int* loop) {
        try {
            member_to_solution(pop, i, variables, s, rem, loop);
        } catch (...) {
            release();
            throw;
        }
    };
    int first_remaining = 0;
    long int first_loop = 0;
    switch (ctx.parameters.storage) {
    case solver_parameters::storage_type::one:
        r.solutions.resize(1);
        fill(0, r.solutions[0], &first_remaining, &first_loop);
        break;
    case solver_parameters::storage_type::bound: {
        r.solutions.resize(2);
        fill(0, r.solutions[0], &first_remaining, &first_loop);
        // get_worst, :328-337: the first member after the front that is not a solution
        int last = 1, rem = 0;
        solution probe;
        while (last != size) {
            fill(last, probe, &rem, nullptr);
            if (rem != 0)
                break;
            ++last;
        }
        fill(last == size ? last - 1 : last, r.solutions[1], nullptr, nullptr);
    } break;
    case solver_parameters::storage_type::five:
        r.solutions.resize(5);
        fill(0, r.solutions[0], &first_remaining, &first_loop);
        for (int i = 1; i != 5 && i < size; ++i)
            fill(i, r.solutions[i], nullptr, nullptr);
        break;
    }
    r.status = first_remaining == 0 ? result_status::success : result_status::time_limit_reached;
    r.loop = first_loop;
    r.remaining_constraints = first_remaining;
    r.duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    release();

    if (ctx.finish)
        ctx.finish(r);
    return r;
}

} // namespace itm
} // namespace baryonyx
